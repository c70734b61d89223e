"""Run under torchrun (one process per GPU): the NCCL transport reproduces the undecomposed run bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py [mesh] [cycles]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "medium"
    cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    transport = sys.argv[3] if len(sys.argv) > 3 else "nccl"          # nccl | ipc
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = ge.load_package()
    mesh = pkg.meshgen.make_multigrid(name)
    parts = pkg.partition_levels(mesh["levels"], mesh["base_array_index"], world)
    lm = pkg.LocalMesh(mesh["levels"], mesh["base_array_index"], parts, rank, world)
    gpu = pkg.MGCFD(local_mesh=lm, device=local, exact_arith=True)
    if transport == "ipc":
        mine = torch.frombuffer(bytearray(gpu.ipc_export()), dtype=torch.uint8).cuda()
        blobs = [torch.zeros(4096, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(blobs, mine)
        gpu.comm_init_ipc(b"".join(b.cpu().numpy().tobytes() for b in blobs))
        dist.barrier()
    else:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(pkg.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        gpu.comm_init_nccl(uid.cpu().numpy().tobytes())
    gpu.run_cycles(cycles)           # pairs of cycles replay as a CUDA graph with the NCCL calls captured
    gpu.run_cycles(2)
    ok = True
    if rank == 0:
        with pkg.MGCFD(mesh["levels"], device=local, exact_arith=True) as single:
            single.run_cycles(cycles + 2)
            refs = [single.fetch(l, "variables") for l in range(len(mesh["levels"]))]
    for l, lev in enumerate(mesh["levels"]):
        n = lev["node_coordinates"].shape[0]
        full = torch.zeros((n, 5), dtype=torch.float64, device="cuda")
        gn = torch.from_numpy(lm.query(l, "global_node")[:gpu.n_owned[l]].astype(np.int64)).cuda()
        full[gn] = torch.from_numpy(gpu.fetch(l, "variables")[:gpu.n_owned[l]]).cuda()
        dist.all_reduce(full)                       # every node is owned exactly once, the rest are zeros
        if rank == 0:
            same = np.array_equal(full.cpu().numpy(), refs[l])
            print(f"level {l}: {transport} x{world} == single GPU bit for bit: {same}")
            ok = ok and same
    halo = gpu.halo_bytes_sent()
    gpu.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", "halo bytes sent by rank 0:", halo)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

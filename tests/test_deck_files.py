"""Deck files: the container that stands in for the reference's HDF5 level files (same dataset names / shapes /
dtypes, euler3d.cpp:248-312) and the input.dat deck (io.h:28-205)."""
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "mg-cfd-app-op2_b200", "euler3d_b200")


def test_container_round_trip(tmp_path, meshgen):
    mesh = meshgen.make_multigrid("tiny")
    deck = meshgen.write_deck(str(tmp_path), mesh)
    text = open(deck).read()
    assert "num_levels = 3" in text and "mesh_name = m6wing" in text and "base_array_index = 1" in text and "[levels]" in text
    for l, lev in enumerate(mesh["levels"]):
        back = meshgen.read_container(os.path.join(str(tmp_path), f"mesh.L{l}.mgb"))
        assert set(back) == set(lev)
        for k in lev:
            assert back[k].dtype == lev[k].dtype and np.array_equal(back[k], lev[k])
    p = meshgen.write_solution(str(tmp_path), 1, 10, np.arange(20.0).reshape(4, 5))
    assert os.path.basename(p) == "solution.variables.L1.cycles=10.mgb"              # Q14 naming
    assert np.array_equal(meshgen.read_container(p)["p_variables_result_L1"], np.arange(20.0).reshape(4, 5))


def test_hdf5_deck_round_trip(tmp_path, meshgen):
    """the same deck as HDF5 level files (the reference's format, euler3d.cpp:248-327) through the library's own
    HDF5 subset: datasets come back bit for bit, and the independent Python restatement reads them too"""
    import h5_oracle
    mesh = meshgen.make_multigrid("tiny")
    meshgen.write_deck(str(tmp_path), mesh, fmt="h5")
    for l, lev in enumerate(mesh["levels"]):
        path = os.path.join(str(tmp_path), f"mesh.L{l}.h5")
        back, indep = meshgen.read_h5(path), h5_oracle.read_h5(path)
        assert set(back) == set(lev) == set(indep)
        for k in lev:
            assert back[k].dtype == lev[k].dtype and np.array_equal(back[k], lev[k])
            assert np.array_equal(indep[k]["data"], lev[k])
    p = meshgen.write_solution(str(tmp_path), 0, 25, np.arange(35.0).reshape(7, 5), fmt="h5")
    assert os.path.basename(p) == "solution.variables.L0.cycles=25.h5"               # the reference's own file name
    assert np.array_equal(meshgen.read_h5(p)["p_variables_result_L0"], np.arange(35.0).reshape(7, 5))


def test_native_driver_loads_hdf5_and_container_decks_alike(tmp_path, meshgen):
    """euler3d_b200 --check-deck (no GPU needed): the loader gets identical sizes and order-sensitive checksums out of
    an HDF5 deck, a container deck, and an HDF5 deck written by the independent Python restatement with chunked,
    shuffled, deflated datasets under a version-2 superblock"""
    import h5_oracle
    mesh = meshgen.make_multigrid("tiny")
    outs = {}
    for fmt in ("mgb", "h5", "h5-chunked"):
        d = tmp_path / fmt
        meshgen.write_deck(str(d), mesh, fmt="h5" if fmt != "mgb" else "mgb")
        if fmt == "h5-chunked":
            for l, lev in enumerate(mesh["levels"]):
                h5_oracle.write_h5(str(d / f"mesh.L{l}.h5"), lev, superblock=2, layout="chunked", filters=("shuffle", "deflate"))
        p = subprocess.run([EXE, "-i", "input.dat", "-d", str(d), "--check-deck"], capture_output=True, text=True, timeout=120)
        assert p.returncode == 0, p.stdout + p.stderr
        outs[fmt] = re.sub(r"format=\w+", "format=*", p.stdout)
        assert ("format=hdf5" in p.stdout) == (fmt != "mgb")
    assert "nodes=378 edges=1000" in outs["mgb"] and "checksum=" in outs["mgb"]
    assert outs["mgb"] == outs["h5"] == outs["h5-chunked"]


def test_perf_csv_files_have_the_reference_layout(tmp_path, meshgen):
    """<prefix>P=<rank>.PerfData.csv and <prefix>P=<rank>.FileIoTimes.csv (io.h:296-417), the files the reference's
    aggregate-output-data.py collects: header once, rows appended on every run; written here by --check-deck (no GPU)"""
    mesh = meshgen.make_multigrid("tiny")
    meshgen.write_deck(str(tmp_path), mesh)
    prefix = str(tmp_path / "out")
    for _ in range(2):
        p = subprocess.run([EXE, "-i", "input.dat", "-d", str(tmp_path), "-o", prefix, "-r", "kway", "--check-deck"],
                           capture_output=True, text=True, timeout=120)
        assert p.returncode == 0, p.stdout + p.stderr
    perf = open(prefix + ".P=0.PerfData.csv").read().splitlines()
    assert perf[0] == "rank,partitioner,kernel,level,computeTime,syncTime,iters"
    assert len(perf) == 1 + 2 * 3 and perf[1].startswith("0,kway,compute_flux_edge_kernel,0,") and perf[3].split(",")[3] == "2"
    io = open(prefix + ".P=0.FileIoTimes.csv").read().splitlines()
    assert io[0] == "rank,partitioner,level,writeInterval,numberOfWrites,fileIoTime,wallTime"
    assert len(io) == 3 and io[1].split(",")[:5] == ["0", "kway", "0", "0", "0"]


def test_native_driver_rejects_malformed_decks(tmp_path, meshgen):
    """op_decl_map_hdf5 / op_decl_dat_hdf5 refuse datasets whose size, dim or type does not match their set
    (euler3d.cpp:262-312); the driver does the same before it takes any pointer, and the container reader bounds every
    length field by the file size (no GPU needed: --check-deck)"""
    mesh = meshgen.make_multigrid("tiny")

    def check(mutate, expect):
        d = tmp_path / expect.replace(" ", "_")[:24]
        meshgen.write_deck(str(d), mesh)
        mutate(str(d))
        p = subprocess.run([EXE, "-i", "input.dat", "-d", str(d), "--check-deck"], capture_output=True, text=True, timeout=120)
        assert p.returncode == 1 and expect in p.stderr, (p.returncode, p.stderr[-400:])

    def rewrite(path, edit):
        lev = meshgen.read_container(path)
        edit(lev)
        meshgen.write_container(path, lev)

    # a weights array one row short of its edge set
    check(lambda d: rewrite(os.path.join(d, "mesh.L0.mgb"), lambda lev: lev.__setitem__("edge_weights", lev["edge_weights"][:-1])),
          "edge_weights does not have the expected type / shape")
    # a map stored as floating point
    check(lambda d: rewrite(os.path.join(d, "mesh.L1.mgb"), lambda lev: lev.__setitem__("bnd_node-->group", lev["bnd_node-->group"].astype(np.float64))),
          "bnd_node-->group does not have the expected type / shape")
    # a map with the wrong second dimension
    check(lambda d: rewrite(os.path.join(d, "mesh.L0.mgb"), lambda lev: lev.__setitem__("edge-->node", np.ascontiguousarray(lev["edge-->node"][:, :1]))),
          "edge-->node does not have the expected type / shape")

    # a truncated file and a header whose byte count runs past the end of the file
    def truncate(d):
        p = os.path.join(d, "mesh.L0.mgb")
        data = open(p, "rb").read()
        open(p, "wb").write(data[:len(data) // 2])
    check(truncate, "mesh.L0.mgb")

    def huge_len(d):
        p = os.path.join(d, "mesh.L0.mgb")
        data = bytearray(open(p, "rb").read())
        data[16:20] = (0x7fffffff).to_bytes(4, "little")           # name length of the first dataset
        open(p, "wb").write(bytes(data))
    check(huge_len, "corrupt dataset name length")

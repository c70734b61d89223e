// flux_fast.cu -- fast-math build of the flux kernels (default nvcc contraction).
#include "flux_kernels.cuh"

namespace mgcfd {
int fast_flux_atomic(cudaStream_t s, const FluxArgs &a, const AtomicPlanDev &p) { return fast::launch_atomic(s, a, p); }
int fast_flux_colour(cudaStream_t s, const FluxArgs &a, const ColourPlanDev &p, const ColourPlanHost &h)
{
    return fast::launch_colour(s, a, p, h);
}
int fast_flux_owner(cudaStream_t s, const FluxArgs &a, const OwnerPlanDev &p, const OwnerPlanHost &h)
{
    return fast::launch_owner(s, a, p, h);
}
int fast_bnd_flux(cudaStream_t s, int n_unique, const int *bu_node, const int *bu_ptr, const int *b_group,
                  const double *b_wt, const double *var, double *flux, const DevConsts &c)
{
    return fast::launch_bnd(s, n_unique, bu_node, bu_ptr, b_group, b_wt, var, flux, c);
}
int fast_flux_gather(cudaStream_t s, const FluxArgs &a, const GatherPlanDev &p, int n_chunks, int max_loc)
{
    return fast::launch_gather(s, a, p, n_chunks, max_loc);
}
size_t fast_gather_smem(int max_loc) { return fast::gather_smem(max_loc, false); }
int flux_emit(cudaStream_t s, const FluxArgs &a, const EmitPlanDev &p) { return fast::launch_emit(s, a, p); }
size_t flux_emit_smem_bytes(int max_loc, int max_ent, int max_blob, int max_own) { return fast::emit_smem(max_loc, max_ent, max_blob, max_own); }
std::string fast_configure() { return fast::configure(); }
bool fast_owner_uses_stage2(const OwnerPlanDev &p, const OwnerPlanHost &h) { return fast::stage2_applies(p, h) && h.max_own <= 64; }
size_t fast_owner_smem(int max_loc, int max_edges, int max_blob) { return fast::owner_smem(max_loc, max_edges, max_blob, false); }
size_t fast_colour_smem(int max_nodes) { return fast::colour_smem(max_nodes, false); }
}  // namespace mgcfd

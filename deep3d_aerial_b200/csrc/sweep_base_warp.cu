// Baseline sweep kernel instantiations for D3D_AGG_WARP (exactly one source view).
#include "sweep_base.cuh"

namespace d3d {
int sweep_base_warp(int cpt, int nv, const SweepParams& p, dim3 grid, cudaStream_t stream) {
    if (nv != 1) return fail(D3D_ERR_BAD_ARGUMENT, "sweep: D3D_AGG_WARP takes one source view");
    if (cpt == 8) return launch_sweep_base<8, 1, D3D_AGG_WARP>(p, grid, stream);
    return launch_sweep_base<4, 1, D3D_AGG_WARP>(p, grid, stream);
}
}  // namespace d3d

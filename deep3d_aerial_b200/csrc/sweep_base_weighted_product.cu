// Baseline sweep kernel instantiations for D3D_AGG_WEIGHTED_PRODUCT (one translation unit per mode so the
// modes compile in parallel).
#include "sweep_base.cuh"

namespace d3d {
int sweep_base_weighted_product(int cpt, int nv, const SweepParams& p, dim3 grid, cudaStream_t stream) {
    if (cpt == 8) return dispatch_sweep_base<8, D3D_AGG_WEIGHTED_PRODUCT>(nv, p, grid, stream);
    return dispatch_sweep_base<4, D3D_AGG_WEIGHTED_PRODUCT>(nv, p, grid, stream);
}
}  // namespace d3d

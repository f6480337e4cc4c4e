// Production sweep kernel for 32-, 16- and 8-channel features (four planes per projection pass): AdaMVS weighted product.
#include "sweep_quad.cuh"

namespace d3d {

int sweep_quad_weighted_product(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee) {
    return sweep_quad_dispatch<D3D_AGG_WEIGHTED_PRODUCT>(nv, p, grid, stream, ieee);
}

}  // namespace d3d

// Production sweep kernel (four planes per projection pass): AdaMVS pair volumes, mean_C(ref * warped_i).
#include "sweep_quad.cuh"

namespace d3d {

int sweep_quad_pair_mean(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee) {
    return sweep_quad_dispatch<D3D_AGG_PAIR_MEAN>(nv, p, grid, stream, ieee);
}

}  // namespace d3d

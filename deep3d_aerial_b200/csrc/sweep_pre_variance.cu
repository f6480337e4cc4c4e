// sweep_quad's formulation with TMA-prefetched footprints (sweep_pre.cuh): variance volume, 32-channel features.
#include "sweep_pre.cuh"

namespace d3d {

int sweep_pre_variance(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream) {
    return sweep_pre_dispatch(nv, p, grid, stream);
}

}  // namespace d3d

// Warp-specialised sweep kernel (producer warps project + prefetch footprints into shared-memory slots, consumer warps
// do the packed arithmetic): variance volume from 32-channel features.
#include "sweep_ws.cuh"

namespace d3d {

int sweep_ws_variance(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream) {
    return sweep_ws_dispatch<D3D_AGG_VARIANCE>(nv, p, grid, stream);
}

}  // namespace d3d

// Window-gather sweep kernel (short per-pixel-hypothesis stages, view-weighted product volume).
#include "sweep_win.cuh"

namespace d3d {

int sweep_win_weighted_product(int nv, const SweepParams& p, cudaStream_t stream, bool ieee_div) {
    return sweep_win_dispatch(nv, p, stream, ieee_div);
}

}  // namespace d3d

// Register-accumulated sweep kernel (short per-pixel-hypothesis stages, view-weighted product volume).
#include "sweep_acc.cuh"

namespace d3d {

int sweep_acc_weighted_product(int nv, const SweepParams& p, cudaStream_t stream, bool ieee_div) {
    return sweep_acc_dispatch(nv, p, stream, ieee_div);
}

}  // namespace d3d

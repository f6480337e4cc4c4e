// Direct-gather sweep kernel (8-channel features): variance and weighted-product volumes.
#include <algorithm>

#include "sweep_direct.cuh"

namespace d3d {

int sweep_direct_variance(int nv, const SweepParams& p, cudaStream_t stream) {
    return sweep_direct_dispatch<D3D_AGG_VARIANCE>(nv, p, stream);
}
int sweep_direct_weighted_product(int nv, const SweepParams& p, cudaStream_t stream) {
    return sweep_direct_dispatch<D3D_AGG_WEIGHTED_PRODUCT>(nv, p, stream);
}

}  // namespace d3d

// Production sweep kernel instantiations for D3D_AGG_VARIANCE: 8 channels per lane, 1..4 source views,
// C = 8, 16, 32, 64 (1, 2, 4, 8 lanes per pixel).
#include "sweep_fast.cuh"

namespace d3d {

template <int LPP>
static int by_views(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee) {
    switch (nv) {
        case 1: return launch_sweep_fast<1, LPP, D3D_AGG_VARIANCE>(p, grid, stream, ieee);
        case 2: return launch_sweep_fast<2, LPP, D3D_AGG_VARIANCE>(p, grid, stream, ieee);
        case 3: return launch_sweep_fast<3, LPP, D3D_AGG_VARIANCE>(p, grid, stream, ieee);
        case 4: return launch_sweep_fast<4, LPP, D3D_AGG_VARIANCE>(p, grid, stream, ieee);
        default: return -1;
    }
}

// returns -1 when the shape is not covered (the caller falls back to the baseline kernel)
int sweep_fast_variance(int cpt, int nv, const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee) {
    if (cpt != 8 || p.W > 16000 || p.H > 16000) return -1;   // 14-bit corner fields in the footprint key
    switch (p.lpp_log2) {
        case 0: return by_views<1>(nv, p, grid, stream, ieee);
        case 1: return by_views<2>(nv, p, grid, stream, ieee);
        case 2: return by_views<4>(nv, p, grid, stream, ieee);
        case 3: return by_views<8>(nv, p, grid, stream, ieee);
        default: return -1;
    }
}

}  // namespace d3d

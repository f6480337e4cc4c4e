// sweep_lean_kernel instantiations for D3D_AGG_VARIANCE, 1..4 source views:
//   4 channels per lane (16 resident warps per SM): C = 4, 8, 16, 32
//   8 channels per lane ( 8 resident warps per SM): C = 64 only (4 per lane would need 16 lanes per pixel)
#include "sweep_lean.cuh"

namespace d3d {

template <int CPT, int LPP>
static int by_views(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee) {
    switch (nv) {
        case 1: return launch_sweep_lean<CPT, 1, LPP, D3D_AGG_VARIANCE>(p, grid, stream, ieee);
        case 2: return launch_sweep_lean<CPT, 2, LPP, D3D_AGG_VARIANCE>(p, grid, stream, ieee);
        case 3: return launch_sweep_lean<CPT, 3, LPP, D3D_AGG_VARIANCE>(p, grid, stream, ieee);
        case 4: return launch_sweep_lean<CPT, 4, LPP, D3D_AGG_VARIANCE>(p, grid, stream, ieee);
        default: return -1;
    }
}

// returns -1 when the shape is not covered (the caller falls back)
int sweep_lean_variance(int cpt, int nv, const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee) {
    if (p.W > 16000 || p.H > 16000) return -1;   // 14-bit corner fields in the footprint key
    // footprint keys are 32-bit byte offsets into `feats`
    if ((unsigned long long)(nv + 1) * (unsigned long long)p.HW * (unsigned long long)p.C * 4ull >= (1ull << 32)) return -1;
    if (cpt == 4) {
        switch (p.lpp_log2) {
            case 0: return by_views<4, 1>(nv, p, grid, stream, ieee);
            case 1: return by_views<4, 2>(nv, p, grid, stream, ieee);
            case 2: return by_views<4, 4>(nv, p, grid, stream, ieee);
            case 3: return by_views<4, 8>(nv, p, grid, stream, ieee);
            default: return -1;
        }
    }
    return p.lpp_log2 == 3 ? by_views<8, 8>(nv, p, grid, stream, ieee) : -1;
}

}  // namespace d3d

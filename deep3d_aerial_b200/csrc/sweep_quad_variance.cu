// Production sweep kernel for 32-channel features (four planes per projection pass), D3D_AGG_VARIANCE,
// 1..4 source views.
#include "sweep_quad.cuh"

namespace d3d {

// returns -1 when the shape is not covered (the caller falls back to sweep_lean / sweep_base)
int sweep_quad_variance(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee) {
    if (p.C != 32 || p.W > 16000 || p.H > 16000) return -1;   // 16-bit corner fields in the footprint key
    // rows leave as 16-byte chunks, every CTA owns a whole 32-pixel row segment
    if ((p.HW & 31) != 0 || ((p.out_sc | p.out_sd) & 3) != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15) != 0) return -1;
    switch (nv) {
        case 1: return launch_sweep_quad<1>(p, grid, stream, ieee);
        case 2: return launch_sweep_quad<2>(p, grid, stream, ieee);
        case 3: return launch_sweep_quad<3>(p, grid, stream, ieee);
        case 4: return launch_sweep_quad<4>(p, grid, stream, ieee);
        default: return -1;
    }
}

}  // namespace d3d

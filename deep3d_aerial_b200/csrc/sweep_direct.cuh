// Fused plane sweep for 8-channel features (cascade stage 3: full resolution, a handful of planes, per-pixel
// hypotheses 0.3-0.4 px apart): direct gather, no footprint cache.
//
//   lane = one reference pixel with all 8 channels; a warp = 32 consecutive pixels, so every channel row of
//   the volume leaves as one coalesced 128-byte store and nothing is staged in shared memory; no state is
//   carried from plane to plane (a footprint would serve ~3 planes of an 8-plane sweep), which keeps the
//   kernel at ~64 registers -- 4x the resident warps of the cached formulation, and that is what hides the
//   gather latency (sweep_base_kernel, 128 cached footprint registers per lane: 2.96 ms on the stage-3 shape).
//   Each lane runs the projection chain of project_frac() (sweep_util.cuh) for its own pixel -- with one lane
//   per pixel nothing is computed twice -- and gathers the four 32-byte texels of every view with one
//   LDG.256; corners outside the image are zero (loads predicated off), as grid_sample's zeros padding.
#pragma once
#include "sweep_util.cuh"

namespace d3d {

// one 32-byte texel (8 channels) with a single 256-bit load (LDG.E.256, sm_100), or zeros
__device__ __forceinline__ void ldg8_or_zero(float4& lo, float4& hi, const float* ptr, bool valid) {
    lo = make_float4(0.f, 0.f, 0.f, 0.f);
    hi = lo;
    if (valid)
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                     : "l"(ptr));
}

template <int NV, int MODE, bool kPerPix>
__global__ void __launch_bounds__(256, 3) sweep_direct_kernel(const SweepParams p) {
    constexpr int C = 8;
    const long long pix_raw = (long long)blockIdx.x * 256 + threadIdx.x;
    const bool live = pix_raw < p.HW;
    const int pix = live ? (int)pix_raw : p.HW - 1;
    const int py = pix / p.W, px = pix - py * p.W;
    const int d0 = p.d_begin + blockIdx.y * p.d_chunk;
    const int d1 = min(d0 + p.d_chunk, p.d_end);

    float rx[NV], ry[NV], rz[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const float* m = p.pose + v * 16;
        rx[v] = fmaf(m[2], 1.f, fmaf(m[1], (float)py, m[0] * (float)px));
        ry[v] = fmaf(m[6], 1.f, fmaf(m[5], (float)py, m[4] * (float)px));
        rz[v] = fmaf(m[10], 1.f, fmaf(m[9], (float)py, m[8] * (float)px));
        if (p.rays) {   // the reference's own rot @ [x,y,1] (cuBLAS), whatever order it rounded in
            const float* rr = p.rays + (size_t)v * 3 * p.HW + pix;
            rx[v] = __ldg(rr); ry[v] = __ldg(rr + p.HW); rz[v] = __ldg(rr + 2 * (size_t)p.HW);
        }
    }
    float2 rf[4];
    {
        const float* rt = p.feats + ((size_t)p.view_tex[0] + pix) * C;
        const float4 a = ldg4(rt), b = ldg4(rt + 4);
        rf[0] = f2(a.x, a.y); rf[1] = f2(a.z, a.w); rf[2] = f2(b.x, b.y); rf[3] = f2(b.z, b.w);
    }
    float wt[NV];
    float winv = 0.f;
    if (MODE == D3D_AGG_WEIGHTED_PRODUCT) {
        float wsum = p.eps_num ? 0.f : 1e-5f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            wt[v] = __ldg(p.weights + (size_t)v * p.HW + pix);
            wsum += wt[v];                                 // adamvs.py:494,506 accumulation order
        }
        winv = __frcp_rn(wsum);
    }
    const float invV = 1.f / (float)(NV + 1);
    const size_t hyp_stride = kPerPix ? (size_t)p.HW : 1;
    const float* hp = p.hyps + (kPerPix ? (size_t)pix : 0) + (size_t)d0 * hyp_stride;
    float* optr = p.out + (size_t)(d0 - p.d_begin) * p.out_sd + pix;

    for (int dd = d0; dd < d1; ++dd, hp += hyp_stride, optr += p.out_sd) {
        const float depth = __ldg(hp);
        float2 s[4], sq[4];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const float* m = p.pose + v * 16;
            const float4 g = project_frac<false>(rx[v], ry[v], rz[v], m[3], m[7], m[11], depth, p);
            const unsigned key = __float_as_uint(g.w);     // [27:14] y0+4 | [13:0] x0+4
            const int x0 = (int)(key & 0x3fffu) - 4, y0 = (int)((key >> 14) & 0x3fffu) - 4;
            const bool xa = (unsigned)x0 < (unsigned)p.W, xb = (unsigned)(x0 + 1) < (unsigned)p.W;
            const bool ya = (unsigned)y0 < (unsigned)p.H, yb = (unsigned)(y0 + 1) < (unsigned)p.H;
            const float* t = p.feats + ((long long)p.view_tex[v + 1] + (long long)y0 * p.W + x0) * C;
            float4 c[4][2];
            ldg8_or_zero(c[0][0], c[0][1], t, xa && ya);
            ldg8_or_zero(c[1][0], c[1][1], t + C, xb && ya);
            ldg8_or_zero(c[2][0], c[2][1], t + (size_t)p.W * C, xa && yb);
            ldg8_or_zero(c[3][0], c[3][1], t + (size_t)p.W * C + C, xb && yb);
            const float2 fx = splat(g.x), fy = splat(g.y), fxy = splat(g.z);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 qa = c[0][j >> 1], qb = c[1][j >> 1], qc = c[2][j >> 1], qd = c[3][j >> 1];
                const float2 a = (j & 1) ? f2(qa.z, qa.w) : f2(qa.x, qa.y), b = (j & 1) ? f2(qb.z, qb.w) : f2(qb.x, qb.y);
                const float2 cc = (j & 1) ? f2(qc.z, qc.w) : f2(qc.x, qc.y), d = (j & 1) ? f2(qd.z, qd.w) : f2(qd.x, qd.y);
                // A + fx*B + fy*C + fxy*D with B = b-a, C = c-a, D = (d-c)-(b-a): the production kernels' form
                const float2 B = __fadd2_rn(b, f2(-a.x, -a.y)), Cc = __fadd2_rn(cc, f2(-a.x, -a.y));
                const float2 D = __fadd2_rn(__fadd2_rn(d, f2(-cc.x, -cc.y)), f2(-B.x, -B.y));
                float2 o = __ffma2_rn(fx, B, a);
                o = __ffma2_rn(fy, Cc, o);
                o = __ffma2_rn(fxy, D, o);
                if (MODE == D3D_AGG_VARIANCE) {
                    if (v == 0) {
                        s[j] = __fadd2_rn(rf[j], o);
                        sq[j] = __ffma2_rn(o, o, __fmul2_rn(rf[j], rf[j]));
                    } else {
                        s[j] = __fadd2_rn(s[j], o);
                        sq[j] = __ffma2_rn(o, o, sq[j]);
                    }
                } else {                                   // sum_v (warped_v * ref) * weight_v
                    s[j] = __ffma2_rn(__fmul2_rn(o, rf[j]), splat(wt[v]),
                                      v == 0 ? splat(p.eps_num ? 1e-5f : 0.f) : s[j]);
                }
            }
        }
        if (live) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 r;
                if (MODE == D3D_AGG_VARIANCE) {
                    const float2 tneg = __fmul2_rn(s[j], splat(-invV));
                    r = __fmul2_rn(__ffma2_rn(tneg, s[j], sq[j]), splat(invV));
                } else {
                    r = __fmul2_rn(s[j], splat(winv));
                }
                asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(optr + (size_t)(2 * j) * p.out_sc), "f"(r.x) : "memory");
                asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(optr + (size_t)(2 * j + 1) * p.out_sc), "f"(r.y) : "memory");
            }
        }
    }
}

template <int NV, int MODE>
int launch_sweep_direct(const SweepParams& p, cudaStream_t stream) {
    // grid: 256 pixels per CTA; depth is split only when the pixels alone leave SMs idle
    const long long tiles = ((long long)p.HW + 255) / 256;
    SweepParams q = p;
    const int d_count = p.d_end - p.d_begin;
    int chunks = 1;
    if (tiles < 4 * 148) chunks = (int)std::min<long long>((4 * 148 + tiles - 1) / tiles, d_count);
    q.d_chunk = (d_count + chunks - 1) / chunks;
    chunks = (d_count + q.d_chunk - 1) / q.d_chunk;
    const dim3 grid((unsigned)tiles, (unsigned)chunks);
    if (p.perpix) sweep_direct_kernel<NV, MODE, true><<<grid, 256, 0, stream>>>(q);
    else sweep_direct_kernel<NV, MODE, false><<<grid, 256, 0, stream>>>(q);
    count_launch();
    return check_launch("sweep_direct_kernel");
}

// returns -1 when the shape is not covered
template <int MODE>
int sweep_direct_dispatch(int nv, const SweepParams& p, cudaStream_t stream) {
    if (p.C != 8 || p.W > 16000 || p.H > 16000) return -1;
    if (reinterpret_cast<uintptr_t>(p.feats) & 31) return -1;      // 256-bit texel loads
    switch (nv) {
        case 1: return launch_sweep_direct<1, MODE>(p, stream);
        case 2: return launch_sweep_direct<2, MODE>(p, stream);
        case 3: return launch_sweep_direct<3, MODE>(p, stream);
        case 4: return launch_sweep_direct<4, MODE>(p, stream);
        default: return -1;
    }
}

}  // namespace d3d

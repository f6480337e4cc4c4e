// Fused plane sweep for the short per-pixel-hypothesis stages of the cascade (stage 3: 8 channels, 8 planes), view-weighted
// product volume: VIEW-OUTER, PLANES-INNER, the volume of a chunk of planes accumulated in registers.
//
//   lane = one reference pixel x CPL channels (blockIdx.z picks the channel group) x a chunk of KP planes (blockIdx.y).
//   The hypotheses of these stages are 0.3-0.4 source pixels apart, so one 2x2 footprint of a view serves ~3 consecutive
//   planes.  sweep_direct.cuh (planes outer, views inner) cannot use that without caching the footprints of every view
//   (128 registers); here the lane walks the KP planes of ONE view at a time, so a single footprint (32 registers at
//   CPL = 8) is alive, re-fetched only when the plane's corner key differs from the cached one (4 x LDG.256 off one
//   address; footprints that touch the image border go through clamped addresses and are zeroed corner by corner on a
//   rare side path, as grid_sample's zeros padding), and the KP x CPL partial sums  sum_v (warped_v * ref) * weight_v
//   stay in registers across the views -- the same operations in the same view order as the other kernels
//   (adamvs.py:494-506).  L1 traffic drops ~2.5x against the direct gather, which was bound by it (l1tex throughput
//   85 % -> 36 %), instructions by 40 % (519 M against 869 M warp instructions at the stage-3 shape); the bilinear
//   weights are applied as the reference's own four corner weights (4 packed operations per channel pair, no
//   differencing at fetch time).  Projection: project_pair() of sweep_util.cuh, two planes per packed chain.  Every
//   channel row of a plane leaves as one coalesced 128-byte store per warp; nothing is staged in shared memory.
//
//   What bounds it now is the latency of the dependent gather (project -> key -> LDG -> arithmetic on nearly every plane of
//   every warp, because SOME lane's footprint moves): 2.96 load-stall cycles per issued instruction with 8 planes per
//   lane at 12 warps per SM (profiles/ncu_r2_acc_stage3.txt).  Production is therefore KP = 4 at 20 warps per SM
//   (stage-3 shape on B200: 1.04 ms smooth depth map / 1.42 ms white-noise depth map, against 1.16 / 1.58 ms for
//   sweep_direct; KP = 8: 1.09 ms).  Measured and not adopted: `prefetch.global.L1` of the next pair's footprints
//   (projection run one pair ahead, across the view boundary too): 1.17 ms; 16 channels per lane for the stage-2 shape:
//   2.5-3.1 ms against sweep_quad's 1.8 ms (216 registers, or spills at 168); two lanes per pixel there: 2.6 ms; the
//   VARIANCE volume of an 8-channel stage in this form (sum and sum of squares double the accumulators: 4 planes per
//   lane at 12 warps per SM, or 2 planes at 20): 1.40 / 1.33 ms against sweep_direct's 1.15 ms -- not instantiated.
//   sweep_win.cuh is this kernel with the footprints read from a TMA-fetched shared-memory window.
#pragma once
#include <type_traits>

#include "sweep_util.cuh"

namespace d3d {

constexpr int kAccThreads = 128;

// CPL channels of a texel, 8 (32 bytes) per 256-bit load (LDG.E.256, sm_100)
template <int CPL>
__device__ __forceinline__ void ldg8(float2 (&c)[CPL / 2], const float* ptr) {
#pragma unroll
    for (int i = 0; i < CPL / 8; ++i)
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(c[4 * i].x), "=f"(c[4 * i].y), "=f"(c[4 * i + 1].x), "=f"(c[4 * i + 1].y), "=f"(c[4 * i + 2].x),
                       "=f"(c[4 * i + 2].y), "=f"(c[4 * i + 3].x), "=f"(c[4 * i + 3].y)
                     : "l"(ptr + 8 * i));
}

// vf + texel * C floats, as one IMAD.WIDE.U32
template <int C>
__device__ __forceinline__ const float* texel_ptr(const float* vf, unsigned texel) {
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(texel), "r"(C * 4u), "l"(vf));
    return reinterpret_cast<const float*>(a);
}

template <int C, bool kPerPix, bool kIeeeDiv, int KP, int MB = (KP == 8 ? 3 : 5), int CPL = 8>
__global__ void __launch_bounds__(kAccThreads, MB) sweep_acc_kernel(const SweepParams p, const int nv) {
    const long long pix_raw = (long long)blockIdx.x * kAccThreads + threadIdx.x;
    const bool live = pix_raw < p.HW;
    const int pix = live ? (int)pix_raw : p.HW - 1;
    const int py = pix / p.W, px = pix - py * p.W;
    const int d0 = p.d_begin + blockIdx.y * KP;
    const int d1 = min(d0 + KP, p.d_end);
    constexpr int NJ = CPL / 2;                            // packed channel pairs per lane
    const int c0 = blockIdx.z * CPL;

    float2 rf[NJ];
    ldg8<CPL>(rf, p.feats + ((size_t)p.view_tex[0] + pix) * C + c0);
    float depth[KP];                                       // planes past the chunk's end repeat the last one (not stored)
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const int dd = min(d0 + k, d1 - 1);
        depth[k] = kPerPix ? __ldg(p.hyps + (size_t)dd * p.HW + pix) : __ldg(p.hyps + dd);
    }
    float wsum = p.eps_num ? 0.f : 1e-5f;
    for (int v = 0; v < nv; ++v) wsum += __ldg(p.weights + (size_t)v * p.HW + pix);       // adamvs.py:494,506 order
    const float winv = __frcp_rn(wsum);

    float2 acc[KP][NJ];
#pragma unroll
    for (int k = 0; k < KP; ++k)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[k][j] = splat(p.eps_num ? 1e-5f : 0.f);

    const int wmax = p.W - 1, hmax = p.H - 1;
    const size_t row = (size_t)p.W * C;
    // rays and translation of view v (module.py:538-541)
    float rx, ry, rz, tx, ty, tz;
    auto load_view = [&](int v) {
        const float* m = p.pose + v * 16;
        rx = fmaf(m[2], 1.f, fmaf(m[1], (float)py, m[0] * (float)px));
        ry = fmaf(m[6], 1.f, fmaf(m[5], (float)py, m[4] * (float)px));
        rz = fmaf(m[10], 1.f, fmaf(m[9], (float)py, m[8] * (float)px));
        if (p.rays) {   // the reference's own rot @ [x,y,1] (cuBLAS), whatever order it rounded in
            const float* rr = p.rays + (size_t)v * 3 * p.HW + pix;
            rx = __ldg(rr); ry = __ldg(rr + p.HW); rz = __ldg(rr + 2 * (size_t)p.HW);
        }
        tx = m[3]; ty = m[7]; tz = m[11];
    };
#pragma unroll 1
    for (int v = 0; v < nv; ++v) {
        load_view(v);
        const float wt = __ldg(p.weights + (size_t)v * p.HW + pix);
        const float* vf = p.feats + (size_t)p.view_tex[v + 1] * C + c0;
        unsigned ckey = 0x7fff7fffu;                       // no footprint yet (x0 = y0 = 32767 cannot occur)
        float2 A[NJ], B[NJ], Cc[NJ], D[NJ];
#pragma unroll
        for (int k = 0; k < KP; k += 2) {
            float4 e[2];
            project_pair<kIeeeDiv>(rx, ry, rz, tx, ty, tz, f2(depth[k], depth[k + 1]), p, e[0], e[1]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned key = __float_as_uint(e[h].w);
                if (key != ckey) {                         // the footprint moved (or first plane): fetch it
                    ckey = key;
                    const int x0 = (int)(short)(key & 0xffffu), y0 = (int)(short)(key >> 16);
                    if ((unsigned)x0 < (unsigned)wmax && (unsigned)y0 < (unsigned)hmax) {       // all four corners inside
                        const float* t = texel_ptr<C>(vf, (unsigned)(y0 * p.W + x0));
                        ldg8<CPL>(A, t);
                        ldg8<CPL>(B, t + C);
                        t += row;
                        ldg8<CPL>(Cc, t);
                        ldg8<CPL>(D, t + C);
                    } else {                               // image border: clamped addresses, zeros padding
                        const int xa = min(max(x0, 0), wmax), xb = min(max(x0 + 1, 0), wmax);
                        const int ya = min(max(y0, 0), hmax), yb = min(max(y0 + 1, 0), hmax);
                        const int ra = ya * p.W, rb = yb * p.W;
                        ldg8<CPL>(A, texel_ptr<C>(vf, (unsigned)(ra + xa)));
                        ldg8<CPL>(B, texel_ptr<C>(vf, (unsigned)(ra + xb)));
                        ldg8<CPL>(Cc, texel_ptr<C>(vf, (unsigned)(rb + xa)));
                        ldg8<CPL>(D, texel_ptr<C>(vf, (unsigned)(rb + xb)));
                        const bool vxa = xa == x0, vxb = xb == x0 + 1, vya = ya == y0, vyb = yb == y0 + 1;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            if (!(vxa && vya)) A[j] = splat(0.f);
                            if (!(vxb && vya)) B[j] = splat(0.f);
                            if (!(vxa && vyb)) Cc[j] = splat(0.f);
                            if (!(vxb && vyb)) D[j] = splat(0.f);
                        }
                    }
                }
                // grid_sample's corner weights (GridSampler.cuh: nw, ne, sw, se) from the fractions
                const float fx = e[h].x, fy = e[h].y, w11 = e[h].z;
                const float w01 = fx - w11, w10 = fy - w11, w00 = (1.f - fx) - w10;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    float2 o = __fmul2_rn(splat(w00), A[j]);
                    o = __ffma2_rn(splat(w01), B[j], o);
                    o = __ffma2_rn(splat(w10), Cc[j], o);
                    o = __ffma2_rn(splat(w11), D[j], o);
                    acc[k + h][j] = __ffma2_rn(__fmul2_rn(o, rf[j]), splat(wt), acc[k + h][j]);   // (warped * ref) * weight
                }
            }
        }
    }
    if (!live) return;
    float* optr = p.out + (size_t)(d0 - p.d_begin) * p.out_sd + (size_t)c0 * p.out_sc + pix;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        if (d0 + k < d1) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float2 r = __fmul2_rn(acc[k][j], splat(winv));
                asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(optr + (size_t)(2 * j) * p.out_sc), "f"(r.x) : "memory");
                asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(optr + (size_t)(2 * j + 1) * p.out_sc), "f"(r.y) : "memory");
            }
        }
        optr += p.out_sd;
    }
}

template <int C, int KP, int MB, int CPL>
int launch_sweep_acc(int nv, const SweepParams& p, cudaStream_t stream, bool ieee_div) {
    const long long tiles = ((long long)p.HW + kAccThreads - 1) / kAccThreads;
    const int chunks = (p.d_end - p.d_begin + KP - 1) / KP;
    if (tiles > 0x7fffffffLL || chunks > 65535) return -1;
    const dim3 grid((unsigned)tiles, (unsigned)chunks, (unsigned)(C / CPL));
    if (p.perpix) {
        if (ieee_div) sweep_acc_kernel<C, true, true, KP, MB, CPL><<<grid, kAccThreads, 0, stream>>>(p, nv);
        else sweep_acc_kernel<C, true, false, KP, MB, CPL><<<grid, kAccThreads, 0, stream>>>(p, nv);
    } else {
        if (ieee_div) sweep_acc_kernel<C, false, true, KP, MB, CPL><<<grid, kAccThreads, 0, stream>>>(p, nv);
        else sweep_acc_kernel<C, false, false, KP, MB, CPL><<<grid, kAccThreads, 0, stream>>>(p, nv);
    }
    count_launch();
    return check_launch("sweep_acc_kernel");
}

// returns -1 when the shape is not covered.  SweepParams.flags (variant >= 16) pick the A/B configurations.
inline int sweep_acc_dispatch(int nv, const SweepParams& p, cudaStream_t stream, bool ieee_div) {
    if (p.W > 16000 || p.H > 16000 || nv < 1 || !p.weights) return -1;
    if (reinterpret_cast<uintptr_t>(p.feats) & 31) return -1;      // 256-bit texel loads
    const int f = p.flags & 7;
    if (p.C == 8) {
        if (f == 1) return launch_sweep_acc<8, 8, 3, 8>(nv, p, stream, ieee_div);
        return launch_sweep_acc<8, 4, 5, 8>(nv, p, stream, ieee_div);
    }
    if (p.C == 16) {
        if (f == 1) return launch_sweep_acc<16, 4, 5, 8>(nv, p, stream, ieee_div);      // two lanes per pixel
        if (f == 2) return launch_sweep_acc<16, 2, 3, 16>(nv, p, stream, ieee_div);
        if (f == 3) return launch_sweep_acc<16, 4, 3, 16>(nv, p, stream, ieee_div);
        return launch_sweep_acc<16, 4, 2, 16>(nv, p, stream, ieee_div);
    }
    if (p.C == 32) return launch_sweep_acc<32, 4, 5, 8>(nv, p, stream, ieee_div);
    return -1;
}

}  // namespace d3d

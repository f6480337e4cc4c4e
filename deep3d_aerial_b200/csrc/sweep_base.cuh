// Fused plane sweep, baseline kernel ("variant 1"): every lane computes its own projection.
//
// Work decomposition
//   lane  = (pixel within warp, channel group of CPT channels);  LPP = C/CPT lanes share a pixel
//   warp  = 32/LPP consecutive reference pixels (flat index y*W+x), all channels
//   CTA   = 8 warps;  blockIdx.y = chunk of depth planes
// Each thread walks its depth planes in order and keeps, per source view, the 2x2 texel footprint
// of its CPT channels in registers; the footprint is re-fetched only when floor(ix), floor(iy)
// move (adjacent planes move the sample by a fraction of a pixel along the epipolar line), so the
// bilinear gather mostly runs out of registers.  The warped V x C x D x H x W volume never exists.
//
// Arithmetic follows mvs/mvs_cas/models/module.py:528-546 and ATen's grid_sampler_2d
// (bilinear, zeros padding, align_corners=True) operation by operation; see project().
#pragma once
#include "common.cuh"

namespace d3d {

// Source-image position of one reference pixel on one depth plane, as ATen sees it.
//   module.py:539-541  X = rot_xyz * d  (rounded)  + trans (rounded)   -- no FMA contraction
//   module.py:542      u = X / Z, v = Y / Z
//   module.py:543-544  g = u / ((W-1)/2) - 1   (ATen-CUDA multiplies by the reciprocal scalar)
//   GridSampler.h:27-32 ix = ((g + 1) / 2) * (W-1)
// Non-finite or far-outside positions are clamped to just outside the image, where every corner is
// out of bounds and contributes zero (SURVEY.md §7 hard part 6).
struct Footprint {
    int x0, y0;   // floor(ix), floor(iy)
    float w[4];   // nw, ne, sw, se  (GridSampler.cu: (ix_se-ix)*(iy_se-iy) ...)
};

template <bool kIeeeDiv>
__device__ __forceinline__ Footprint project(float rx, float ry, float rz, float tx, float ty, float tz,
                                             float d, const SweepParams& p) {
    float X = __fadd_rn(__fmul_rn(rx, d), tx);
    float Y = __fadd_rn(__fmul_rn(ry, d), ty);
    float Z = __fadd_rn(__fmul_rn(rz, d), tz);
    float u, v;
    if (kIeeeDiv) {
        u = __fdiv_rn(X, Z);
        v = __fdiv_rn(Y, Z);
    } else {
        float r = __frcp_rn(Z);
        u = __fmul_rn(X, r);
        v = __fmul_rn(Y, r);
    }
    float ix = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(u, p.inv_half_w), 1.f), 1.f), 0.5f), p.wm1);
    float iy = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(v, p.inv_half_h), 1.f), 1.f), 0.5f), p.hm1);
    ix = fminf(fmaxf(ix, -2.f), p.wm1 + 2.f);   // NaN -> -2
    iy = fminf(fmaxf(iy, -2.f), p.hm1 + 2.f);
    float fx0 = floorf(ix), fy0 = floorf(iy);
    Footprint f;
    f.x0 = (int)fx0;
    f.y0 = (int)fy0;
    float ax = __fsub_rn(__fadd_rn(fx0, 1.f), ix), bx = __fsub_rn(ix, fx0);
    float ay = __fsub_rn(__fadd_rn(fy0, 1.f), iy), by = __fsub_rn(iy, fy0);
    f.w[0] = __fmul_rn(ax, ay);
    f.w[1] = __fmul_rn(bx, ay);
    f.w[2] = __fmul_rn(ax, by);
    f.w[3] = __fmul_rn(bx, by);
    return f;
}

template <int CPT>
__device__ __forceinline__ void load_texel(float (&dst)[CPT], const float* __restrict__ view, int x, int y,
                                           const SweepParams& p, int choff) {
    bool inb = (unsigned)x < (unsigned)p.W && (unsigned)y < (unsigned)p.H;
    if (inb) {
        const float* t = view + ((size_t)y * p.W + x) * p.C + choff;
#pragma unroll
        for (int k = 0; k < CPT; k += 4) {
            float4 q = ldg4(t + k);
            dst[k] = q.x; dst[k + 1] = q.y; dst[k + 2] = q.z; dst[k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < CPT; ++k) dst[k] = 0.f;
    }
}

template <int CPT, int NV, int MODE>
__global__ void __launch_bounds__(256) sweep_base_kernel(const SweepParams p) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int lpp = 1 << p.lpp_log2;
    const int ppw = 32 >> p.lpp_log2;                   // pixels per warp
    const int cg = lane & (lpp - 1);
    const int choff = cg * CPT;
    long long pix_raw = ((long long)blockIdx.x * 8 + warp) * ppw + (lane >> p.lpp_log2);
    const bool live = pix_raw < p.HW;
    const int pix = live ? (int)pix_raw : p.HW - 1;     // clamp: keeps the warp whole for shuffles
    const int py = pix / p.W, px = pix - py * p.W;

    const int d0 = p.d_begin + blockIdx.y * p.d_chunk;
    const int d1 = min(d0 + p.d_chunk, p.d_end);

    // rays: rot @ [x, y, 1]  (module.py:538; cuBLAS accumulates with FMAs)
    float rx[NV], ry[NV], rz[NV], tx[NV], ty[NV], tz[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const float* m = p.pose + v * 16;
        rx[v] = fmaf(m[2], 1.f, fmaf(m[1], (float)py, m[0] * (float)px));
        ry[v] = fmaf(m[6], 1.f, fmaf(m[5], (float)py, m[4] * (float)px));
        rz[v] = fmaf(m[10], 1.f, fmaf(m[9], (float)py, m[8] * (float)px));
        if (p.rays) {   // the reference's own rot @ [x,y,1] (cuBLAS), whatever order it rounded in
            const float* rr = p.rays + (size_t)(v) * 3 * p.HW + pix;
            rx[v] = __ldg(rr); ry[v] = __ldg(rr + p.HW); rz[v] = __ldg(rr + 2 * (size_t)p.HW);
        }
        tx[v] = m[3]; ty[v] = m[7]; tz[v] = m[11];
    }

    float rf[CPT];
    {
        const float* t = p.feats + ((size_t)p.view_tex[0] + pix) * p.C + choff;
#pragma unroll
        for (int k = 0; k < CPT; k += 4) {
            float4 q = ldg4(t + k);
            rf[k] = q.x; rf[k + 1] = q.y; rf[k + 2] = q.z; rf[k + 3] = q.w;
        }
    }
    float wt[NV];
    float wsum = 0.f;
    if (MODE == D3D_AGG_WEIGHTED_PRODUCT) {
        wsum = p.eps_num ? 0.f : 1e-5f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            wt[v] = __ldg(p.weights + (size_t)v * p.HW + pix);
            wsum += wt[v];                       // adamvs.py:494,506 accumulation order
        }
    }

    float tex[NV][4][CPT];
    int cx[NV], cy[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) { cx[v] = INT_MIN; cy[v] = INT_MIN; }

    const float invV = 1.f / (float)(NV + 1);
    const size_t hyp_stride = p.perpix ? (size_t)p.HW : 1;
    const float* hp = p.hyps + (p.perpix ? (size_t)pix : 0);
    float dnext = (d0 < d1) ? __ldg(hp + (size_t)d0 * hyp_stride) : 0.f;

    for (int dd = d0; dd < d1; ++dd) {
        const float depth = dnext;
        if (dd + 1 < d1) dnext = __ldg(hp + (size_t)(dd + 1) * hyp_stride);

        float acc[CPT], acc2[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            if (MODE == D3D_AGG_VARIANCE) { acc[c] = rf[c]; acc2[c] = rf[c] * rf[c]; }
            else if (MODE == D3D_AGG_WEIGHTED_PRODUCT) { acc[c] = p.eps_num ? 1e-5f : 0.f; acc2[c] = 0.f; }
            else { acc[c] = 0.f; acc2[c] = 0.f; }
        }
        float* oplane = p.out + (size_t)(dd - p.d_begin) * p.out_sd + pix;

#pragma unroll
        for (int v = 0; v < NV; ++v) {
            Footprint f = project<true>(rx[v], ry[v], rz[v], tx[v], ty[v], tz[v], depth, p);
            if (f.x0 != cx[v] || f.y0 != cy[v]) {
                const float* view = p.feats + (size_t)p.view_tex[v + 1] * p.C;
                load_texel<CPT>(tex[v][0], view, f.x0, f.y0, p, choff);
                load_texel<CPT>(tex[v][1], view, f.x0 + 1, f.y0, p, choff);
                load_texel<CPT>(tex[v][2], view, f.x0, f.y0 + 1, p, choff);
                load_texel<CPT>(tex[v][3], view, f.x0 + 1, f.y0 + 1, p, choff);
                cx[v] = f.x0; cy[v] = f.y0;
            }
            float wv[CPT];
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                float o = tex[v][0][c] * f.w[0];
                o = fmaf(tex[v][1][c], f.w[1], o);
                o = fmaf(tex[v][2][c], f.w[2], o);
                o = fmaf(tex[v][3][c], f.w[3], o);
                wv[c] = o;
            }
            if (MODE == D3D_AGG_VARIANCE) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) { acc[c] += wv[c]; acc2[c] = fmaf(wv[c], wv[c], acc2[c]); }
            } else if (MODE == D3D_AGG_WARP) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) acc[c] = wv[c];
            } else if (MODE == D3D_AGG_GROUP_CORR) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) acc[c] = fmaf(rf[c], wv[c], acc[c]);
            } else if (MODE == D3D_AGG_WEIGHTED_PRODUCT) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) acc[c] = fmaf(wv[c] * rf[c], wt[v], acc[c]);
            } else if (MODE == D3D_AGG_PAIR_MEAN) {
                float dot = 0.f;
#pragma unroll
                for (int c = 0; c < CPT; ++c) dot = fmaf(rf[c], wv[c], dot);
                for (int o = 1; o < lpp; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                if (live && cg == 0) oplane[(size_t)v * p.out_sc] = dot * (1.f / (float)p.C);
            }
        }

        if (MODE == D3D_AGG_VARIANCE) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                float m = acc[c] * invV;
                float r = fmaf(-m, m, acc2[c] * invV);
                if (live) oplane[(size_t)(choff + c) * p.out_sc] = r;
            }
        } else if (MODE == D3D_AGG_WARP) {
#pragma unroll
            for (int c = 0; c < CPT; ++c)
                if (live) oplane[(size_t)(choff + c) * p.out_sc] = acc[c];
        } else if (MODE == D3D_AGG_WEIGHTED_PRODUCT) {
            const float r = __frcp_rn(wsum);
#pragma unroll
            for (int c = 0; c < CPT; ++c)
                if (live) oplane[(size_t)(choff + c) * p.out_sc] = acc[c] * r;
        } else if (MODE == D3D_AGG_GROUP_CORR) {
            const int gs = p.C / p.groups;          // channels per group
            const float scale = 1.f / ((float)gs * (float)NV);
            if (gs <= CPT) {                        // whole groups inside the thread
                for (int g0 = 0; g0 < CPT; g0 += gs) {
                    float s = 0.f;
#pragma unroll
                    for (int c = 0; c < CPT; ++c)
                        if (c >= g0 && c < g0 + gs) s += acc[c];
                    if (live) oplane[(size_t)((choff + g0) / gs) * p.out_sc] = s * scale;
                }
            } else {                                // a group spans gs/CPT neighbouring lanes
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < CPT; ++c) s += acc[c];
                const int span = gs / CPT;
                for (int o = 1; o < span; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (live && (cg & (span - 1)) == 0) oplane[(size_t)(choff / gs) * p.out_sc] = s * scale;
            }
        }
    }
}

template <int CPT, int NV, int MODE>
int launch_sweep_base(const SweepParams& p, dim3 grid, cudaStream_t stream) {
    sweep_base_kernel<CPT, NV, MODE><<<grid, 256, 0, stream>>>(p);
    count_launch();
    return check_launch("sweep_base_kernel");
}

// Dispatch over the number of source views (1..8) for one aggregation mode.
template <int CPT, int MODE>
int dispatch_sweep_base(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream) {
    switch (nv) {
        case 1: return launch_sweep_base<CPT, 1, MODE>(p, grid, stream);
        case 2: return launch_sweep_base<CPT, 2, MODE>(p, grid, stream);
        case 3: return launch_sweep_base<CPT, 3, MODE>(p, grid, stream);
        case 4: return launch_sweep_base<CPT, 4, MODE>(p, grid, stream);
        default: break;
    }
    if (CPT == 4) {
        switch (nv) {
            case 5: return launch_sweep_base<4, 5, MODE>(p, grid, stream);
            case 6: return launch_sweep_base<4, 6, MODE>(p, grid, stream);
            case 7: return launch_sweep_base<4, 7, MODE>(p, grid, stream);
            case 8: return launch_sweep_base<4, 8, MODE>(p, grid, stream);
            default: break;
        }
    }
    return fail(D3D_ERR_UNSUPPORTED, "sweep: %d source views with %d channels per lane not instantiated", nv, CPT);
}

}  // namespace d3d

// Fused depth regression: softmax over the D planes of a logit volume, expected depth, photometric
// confidence, arg-max plane, optional UCS-Net spread and optional next-stage hypotheses, in ONE
// pass over the logits (4 B/voxel of HBM traffic; the reference makes 4-8 passes).
//
//   one thread = one pixel; consecutive lanes = consecutive x, so every plane access of a warp is a
//   single 128-byte line; planes are consumed 8 at a time (8 independent loads in flight per
//   thread) with one online-softmax rescale per group.
//
// Reference lines: cas_mvsnet.py:69-76 (softmax, regression, 4-plane window confidence),
// msrednet.py:234-238 / adamvs.py:306-310,478-483 (max-prob confidence), adamvs.py:514-529 /
// msrednet.py:418-437 (streaming un-normalised exp), module.py:605-613 (depth_regression incl. the
// bilinear resize of 4-D hypotheses), ucsnet.py:148-149 (exp_variance), module.py:616-630
// (next-stage samples).
#include <type_traits>

#include "common.cuh"

namespace d3d {

struct RegressParams {
    const float* __restrict__ logits;
    const float* __restrict__ hyps;
    float* __restrict__ depth;
    float* __restrict__ conf;
    int* __restrict__ index;
    float* __restrict__ state;
    float* __restrict__ expvar;
    float* __restrict__ next_hyps;
    long long stride_d;
    int D, H, W, HW;
    int d_begin, d_count;
    int conf_mode, hyps_mode, hh, hw, finalize;
    int identity;           // D3D_SOFTMAX_NONE: the input is a probability volume
    int next_nd;
    float next_half_span;   // (float)(next_nd/2 * interval), module.py:619-620
    float lamb;
    float scale_h, scale_w; // hh/H, hw/W for the align_corners=False resize
    const float* planes[D3D_REGRESS_MAX_PLANES];   // scattered [H,W] planes of the slice (logits == nullptr)
};

constexpr int kGroup = 8;

// Bilinear tap of ATen's upsample_bilinear2d (align_corners=False):
//   src = scale*(dst+0.5)-0.5 clamped at 0;  i1 = (int)src;  lambda1 = src-i1;  lambda0 = 1-lambda1
struct ResizeTap {
    int o00, o01, o10, o11;
    float h0, h1, w0, w1;
};

__device__ __forceinline__ ResizeTap make_tap(int y, int x, const RegressParams& p) {
    float sy = fmaxf(p.scale_h * ((float)y + 0.5f) - 0.5f, 0.f);
    float sx = fmaxf(p.scale_w * ((float)x + 0.5f) - 0.5f, 0.f);
    int y1 = (int)sy, x1 = (int)sx;
    int yp = (y1 < p.hh - 1) ? 1 : 0, xp = (x1 < p.hw - 1) ? 1 : 0;
    ResizeTap t;
    t.h1 = sy - (float)y1; t.h0 = 1.f - t.h1;
    t.w1 = sx - (float)x1; t.w0 = 1.f - t.w1;
    t.o00 = y1 * p.hw + x1;
    t.o01 = t.o00 + xp;
    t.o10 = t.o00 + yp * p.hw;
    t.o11 = t.o10 + xp;
    return t;
}

template <int HYPS>
__device__ __forceinline__ float hyp_at(const RegressParams& p, int k, int pix, const ResizeTap& t) {
    if (HYPS == D3D_HYPS_UNIFORM) return __ldg(p.hyps + k);
    if (HYPS == D3D_HYPS_PER_PIXEL) return __ldg(p.hyps + (size_t)k * p.HW + pix);
    const float* q = p.hyps + (size_t)k * p.hh * p.hw;
    return t.h0 * (t.w0 * __ldg(q + t.o00) + t.w1 * __ldg(q + t.o01)) +
           t.h1 * (t.w0 * __ldg(q + t.o10) + t.w1 * __ldg(q + t.o11));
}

__device__ __forceinline__ void emit_next(const RegressParams& p, int pix, float depth) {
    if (p.next_nd <= 0) return;
    // module.py:619-627: lo = cur - nd/2*itv; hi = cur + nd/2*itv; step = (hi-lo)/(nd-1); lo + k*step
    float lo = __fsub_rn(depth, p.next_half_span);
    float hi = __fadd_rn(depth, p.next_half_span);
    float step = __fdiv_rn(__fsub_rn(hi, lo), (float)(p.next_nd - 1));
    for (int k = 0; k < p.next_nd; ++k)
        p.next_hyps[(size_t)k * p.HW + pix] = __fadd_rn(lo, __fmul_rn((float)k, step));
}

// ---- F.softmax flavour -------------------------------------------------------------------------
// PARTS threads share a pixel, each walking a contiguous quarter of the planes with its own online-softmax state; the
// states are merged through shared memory (max of the maxima, every sum rescaled by exp(m_i - M); on a tie the part
// with the lower planes wins, which keeps "the first plane reaching the maximum").  Why: one thread per pixel is one
// long-running thread per pixel -- 319 232 pixels are barely more threads than the GPU holds at once (303 104), so the
// grid ran as one full wave plus a 5 % straggler wave that cost as much as the first (0.185 ms for 490 MB = 40 % of the
// HBM peak, profiles/ncu_r1n.txt).  Four times as many threads, each a quarter as long: ~6 waves.
template <int HYPS, int PARTS>
__global__ void __launch_bounds__(256) regress_softmax_kernel(const RegressParams p) {
    constexpr int PIXB = 256 / PARTS;                  // pixels per CTA; consecutive lanes = consecutive pixels
    __shared__ float part_state[PARTS > 1 ? PARTS : 1][6][PIXB];
    const int pl = threadIdx.x % PIXB, part = threadIdx.x / PIXB;
    const long long pix_raw = (long long)blockIdx.x * PIXB + pl;
    const bool live = pix_raw < p.HW;
    if (PARTS == 1 && !live) return;
    const int pix = live ? (int)pix_raw : p.HW - 1;
    ResizeTap tap = {};
    if (HYPS == D3D_HYPS_RESIZED) tap = make_tap(pix / p.W, pix % p.W, p);
    const float* lg = p.logits + pix;

    float m = -INFINITY;       // running max
    float s = 0.f;             // sum exp(x-m)
    float sd = 0.f;            // sum exp(x-m) * (d - dref)
    float sk = 0.f;            // sum exp(x-m) * k
    float sdd = 0.f;           // sum exp(x-m) * (d - dref)^2
    int arg = 0;
    const float dref = hyp_at<HYPS>(p, 0, pix, tap);   // shift keeps the sums well conditioned
    const bool want_var = p.expvar != nullptr;

    const int per = PARTS == 1 ? p.D : ((p.D + PARTS * kGroup - 1) / (PARTS * kGroup)) * kGroup;   // planes per part
    const int kbeg = part * per, kend = min(p.D, kbeg + per);
    // two groups of planes in flight: the loads of group g + 1 are issued before group g is folded in (a thread then has
    // 16 logit loads outstanding instead of 8: the kernel is a pure stream and its speed is bytes in flight per SM)
    auto load_group = [&](int k0, float (&x)[kGroup], float (&d)[kGroup]) {
#pragma unroll
        for (int j = 0; j < kGroup; ++j) {
            const int k = k0 + j;
            const bool ok = k < kend;
            x[j] = ok ? __ldg(lg + (size_t)k * p.stride_d) : -INFINITY;
            d[j] = ok ? hyp_at<HYPS>(p, k, pix, tap) : dref;
        }
    };
    auto fold_group = [&](int k0, const float (&x)[kGroup], const float (&d)[kGroup]) {
        float gm = x[0];
#pragma unroll
        for (int j = 1; j < kGroup; ++j) gm = fmaxf(gm, x[j]);
        if (gm > m) {
            float r = expf(m - gm);   // 0 on the first group (m = -inf)
            s *= r; sd *= r; sk *= r; sdd *= r;
#pragma unroll
            for (int j = kGroup - 1; j >= 0; --j)
                if (x[j] == gm) arg = k0 + j;          // first plane reaching the new maximum
            m = gm;
        }
#pragma unroll
        for (int j = 0; j < kGroup; ++j) {
            float e = expf(x[j] - m);
            float dc = d[j] - dref;
            s += e;
            sd = fmaf(e, dc, sd);
            sk = fmaf(e, (float)(k0 + j), sk);
            if (want_var) sdd = fmaf(e * dc, dc, sdd);
        }
    };
    float xa[kGroup], da[kGroup], xb[kGroup], db[kGroup];
    if (kbeg < kend) load_group(kbeg, xa, da);
    for (int k0 = kbeg; k0 < kend; k0 += 2 * kGroup) {
        if (k0 + kGroup < kend) load_group(k0 + kGroup, xb, db);
        fold_group(k0, xa, da);
        if (k0 + kGroup < kend) {
            if (k0 + 2 * kGroup < kend) load_group(k0 + 2 * kGroup, xa, da);
            fold_group(k0 + kGroup, xb, db);
        }
    }
    if (PARTS > 1) {
        part_state[part][0][pl] = m; part_state[part][1][pl] = s; part_state[part][2][pl] = sd;
        part_state[part][3][pl] = sk; part_state[part][4][pl] = sdd; part_state[part][5][pl] = __int_as_float(arg);
        __syncthreads();
        if (part != 0 || !live) return;
#pragma unroll
        for (int i = 1; i < PARTS; ++i) {
            const float mi = part_state[i][0][pl];
            if (mi == -INFINITY) continue;             // a part without planes (D smaller than its start)
            if (mi > m) {
                const float r = expf(m - mi);
                s *= r; sd *= r; sk *= r; sdd *= r;
                m = mi;
                arg = __float_as_int(part_state[i][5][pl]);
            }
            const float r = expf(mi - m);
            s = fmaf(part_state[i][1][pl], r, s);
            sd = fmaf(part_state[i][2][pl], r, sd);
            sk = fmaf(part_state[i][3][pl], r, sk);
            sdd = fmaf(part_state[i][4][pl], r, sdd);
        }
    }
    const float inv = 1.f / s;
    const float mean_c = sd * inv;
    const float depth = dref + mean_c;
    p.depth[pix] = depth;
    float conf;
    int idx;
    if (p.conf_mode == D3D_CONF_MAX_PROB) {
        conf = inv;          // exp(max-max)/sum
        idx = arg;
    } else {
        // cas_mvsnet.py:72-76: i = clamp(long(sum p*k)); conf = p[i-1]+p[i]+p[i+1]+p[i+2]
        idx = (int)(sk * inv);
        idx = max(0, min(idx, p.D - 1));
        float w = 0.f;
#pragma unroll
        for (int j = -1; j <= 2; ++j) {
            int k = idx + j;
            if (k >= 0 && k < p.D) w += expf(__ldg(lg + (size_t)k * p.stride_d) - m);
        }
        conf = w * inv;
    }
    p.conf[pix] = conf;
    if (p.index) p.index[pix] = idx;
    if (want_var) {
        float var = fmaxf(sdd * inv - mean_c * mean_c, 0.f);
        p.expvar[pix] = p.lamb * sqrtf(var);
    }
    emit_next(p, pix, depth);
}

// The same, four consecutive pixels per thread: every logit (and per-pixel hypothesis) load is 16 bytes, a thread has
// 8 x 16 bytes in flight per group of planes.  A pure stream lives on bytes in flight per SM; with 4-byte loads the kernel
// above sits at 44 % of the HBM peak.  Needs H*W, the plane stride and the base addresses to be multiples of 4 floats;
// uniform or per-pixel hypotheses (resized ones keep the scalar kernel: their taps are gathers).
// kWin4 / kVar: the sums only the window-4 confidence (sum p*k) and exp_variance (sum p*d^2) need are compiled out
// otherwise.  exp() is ex2.approx(x * log2 e): every argument is <= 0 after the max subtraction and the terms that matter
// have small |x|, where its error is ~3e-7 relative -- three orders below the depth tolerance; the accurate expf() cost
// more instructions than the rest of the element put together (profiles: 63 % issue utilisation at 55 % of the HBM peak).
template <int HYPS, int PARTS, int G, bool kWin4, bool kVar>
__global__ void __launch_bounds__(256, (G == 8 && HYPS == D3D_HYPS_UNIFORM) ? 3 : 2) regress_softmax_vec_kernel(const RegressParams p) {
    constexpr int QB = 256 / PARTS;                    // pixel quads per CTA
    __shared__ float part_state[PARTS > 1 ? PARTS : 1][6][QB * 4];
    const int ql = threadIdx.x % QB, part = threadIdx.x / QB;
    const long long quad_raw = (long long)blockIdx.x * QB + ql;
    const bool live = quad_raw * 4 < p.HW;
    if (PARTS == 1 && !live) return;
    const int pix0 = live ? (int)(quad_raw * 4) : p.HW - 4;
    const float* lg = p.logits + pix0;
    const ResizeTap tap = {};

    float m[4], s[4], sd[4], sk[4], sdd[4], dref[4];
    int arg[4];
    if (HYPS == D3D_HYPS_UNIFORM) {
        const float d0 = __ldg(p.hyps);
#pragma unroll
        for (int c = 0; c < 4; ++c) dref[c] = d0;
    } else {
        const float4 d0 = ldg4(p.hyps + pix0);
        dref[0] = d0.x; dref[1] = d0.y; dref[2] = d0.z; dref[3] = d0.w;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) { m[c] = -INFINITY; s[c] = sd[c] = sk[c] = sdd[c] = 0.f; arg[c] = 0; }
    constexpr bool want_var = kVar;
    const int per = PARTS == 1 ? p.D : ((p.D + PARTS * G - 1) / (PARTS * G)) * G;
    const int kbeg = part * per, kend = min(p.D, kbeg + per);
    for (int k0 = kbeg; k0 < kend; k0 += G) {
        float x[G][4], d[G][4];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int k = k0 + j;
            const bool ok = k < kend;
            const float4 v = ok ? ldg4(lg + (size_t)k * p.stride_d) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            x[j][0] = v.x; x[j][1] = v.y; x[j][2] = v.z; x[j][3] = v.w;
            if (HYPS == D3D_HYPS_UNIFORM) {
                const float dk = ok ? __ldg(p.hyps + k) : dref[0];
#pragma unroll
                for (int c = 0; c < 4; ++c) d[j][c] = dk;
            } else {
                const float4 w = ok ? ldg4(p.hyps + (size_t)k * p.HW + pix0) : make_float4(dref[0], dref[1], dref[2], dref[3]);
                d[j][0] = w.x; d[j][1] = w.y; d[j][2] = w.z; d[j][3] = w.w;
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float gm = x[0][c];
#pragma unroll
            for (int j = 1; j < G; ++j) gm = fmaxf(gm, x[j][c]);
            if (gm > m[c]) {
                const float r = __expf(m[c] - gm);
                s[c] *= r; sd[c] *= r; sk[c] *= r; sdd[c] *= r;
#pragma unroll
                for (int j = G - 1; j >= 0; --j)
                    if (x[j][c] == gm) arg[c] = k0 + j;
                m[c] = gm;
            }
            const float kf = (float)k0;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const float e = __expf(x[j][c] - m[c]);
                const float dc = d[j][c] - dref[c];
                s[c] += e;
                sd[c] = fmaf(e, dc, sd[c]);
                if (kWin4) sk[c] = fmaf(e, kf + (float)j, sk[c]);
                if (want_var) sdd[c] = fmaf(e * dc, dc, sdd[c]);
            }
        }
    }
    if (PARTS > 1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int o = ql * 4 + c;
            part_state[part][0][o] = m[c]; part_state[part][1][o] = s[c]; part_state[part][2][o] = sd[c];
            part_state[part][3][o] = sk[c]; part_state[part][4][o] = sdd[c]; part_state[part][5][o] = __int_as_float(arg[c]);
        }
        __syncthreads();
        if (part != 0 || !live) return;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int o = ql * 4 + c;
#pragma unroll
            for (int i = 1; i < PARTS; ++i) {
                const float mi = part_state[i][0][o];
                if (mi == -INFINITY) continue;
                if (mi > m[c]) {
                    const float r = __expf(m[c] - mi);
                    s[c] *= r; sd[c] *= r; sk[c] *= r; sdd[c] *= r;
                    m[c] = mi;
                    arg[c] = __float_as_int(part_state[i][5][o]);
                }
                const float r = __expf(mi - m[c]);
                s[c] = fmaf(part_state[i][1][o], r, s[c]);
                sd[c] = fmaf(part_state[i][2][o], r, sd[c]);
                sk[c] = fmaf(part_state[i][3][o], r, sk[c]);
                sdd[c] = fmaf(part_state[i][4][o], r, sdd[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int pix = pix0 + c;
        const float inv = 1.f / s[c];
        const float mean_c = sd[c] * inv;
        const float depth = dref[c] + mean_c;
        p.depth[pix] = depth;
        float conf;
        int idx;
        if (!kWin4) {
            conf = inv;
            idx = arg[c];
        } else {
            idx = (int)(sk[c] * inv);
            idx = max(0, min(idx, p.D - 1));
            float w = 0.f;
#pragma unroll
            for (int j = -1; j <= 2; ++j) {
                const int k = idx + j;
                if (k >= 0 && k < p.D) w += __expf(__ldg(p.logits + pix + (size_t)k * p.stride_d) - m[c]);
            }
            conf = w * inv;
        }
        p.conf[pix] = conf;
        if (p.index) p.index[pix] = idx;
        if (want_var) {
            const float var = fmaxf(sdd[c] * inv - mean_c * mean_c, 0.f);
            p.expvar[pix] = p.lamb * sqrtf(var);
        }
        emit_next(p, pix, depth);
    }
    (void)tap;
}

// ---- streaming un-normalised exp flavour ---------------------------------------------------------
template <int HYPS>
__global__ void __launch_bounds__(256) regress_rawexp_kernel(const RegressParams p) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p.HW) return;
    ResizeTap tap = {};
    if (HYPS == D3D_HYPS_RESIZED) tap = make_tap(pix / p.W, pix % p.W, p);
    const float* lg = p.logits + pix;
    float s = 0.f, acc = 0.f, mx = 0.f;      // adamvs.py:456-462: zero-initialised accumulators
    if (p.d_begin > 0) {
        s = p.state[pix];
        acc = p.state[(size_t)p.HW + pix];
        mx = p.state[2 * (size_t)p.HW + pix];
    }
    // Resized hypotheses cost four tap loads per plane on top of the logit: all five loads of every plane of a
    // group are issued before anything is combined (a thread then has 5 * G loads in flight; written the obvious
    // way the compiler interleaved loads and the tap arithmetic, and the kernel sat on the long scoreboard for 11
    // cycles per issued instruction, profiles/ncu_r1_cfg3.txt).
    constexpr int G = (HYPS == D3D_HYPS_RESIZED) ? 4 : kGroup;
    const size_t hplane = (size_t)p.hh * p.hw;
    const float* row0 = p.hyps + (size_t)p.d_begin * hplane + tap.o00;       // resized hypotheses: north-west / south-west tap
    const float* row1 = p.hyps + (size_t)p.d_begin * hplane + tap.o10;
    const int xstep = tap.o01 - tap.o00;                                     // 1, or 0 in the last column
    const bool step1 = __all_sync(__activemask(), xstep == 1);
    for (int k0 = 0; k0 < p.d_count; k0 += G) {
        float x[G], d[G];
        if (HYPS == D3D_HYPS_RESIZED) {
            // Round 2: with the loads batched the kernel was ISSUE-bound (79 % issue utilisation at 36 % of the HBM rate, 76
            // instructions per plane and warp, a third of them address arithmetic).  The four taps of a plane sit at fixed
            // offsets from its north-west tap: two row pointers run from plane to plane, and when every lane of the warp
            // has a right-hand neighbour (everywhere but the last column) the east taps are an immediate +1 -- 4 address
            // instructions per plane where forming each tap's address from the plane index took 20.
            float t00[G], t01[G], t10[G], t11[G];
            auto load_taps = [&](auto step_c) {
                constexpr bool kStep1 = decltype(step_c)::value;
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const int k = k0 + j;
                    const bool ok = k < p.d_count;
                    const float* src = p.logits ? lg + (size_t)k * p.stride_d : p.planes[ok ? k : 0] + pix;
                    x[j] = ok ? __ldg(src) : -INFINITY;
                    t00[j] = t01[j] = t10[j] = t11[j] = 0.f;
                    if (ok) {                              // (planes past the slice are not dereferenced)
                        t00[j] = __ldg(row0); t01[j] = __ldg(row0 + (kStep1 ? 1 : xstep));
                        t10[j] = __ldg(row1); t11[j] = __ldg(row1 + (kStep1 ? 1 : xstep));
                    }
                    row0 += hplane;
                    row1 += hplane;
                }
            };
            if (step1) load_taps(std::true_type()); else load_taps(std::false_type());
#pragma unroll
            for (int j = 0; j < G; ++j)                  // the expression of hyp_at<RESIZED>
                d[j] = tap.h0 * (tap.w0 * t00[j] + tap.w1 * t01[j]) + tap.h1 * (tap.w0 * t10[j] + tap.w1 * t11[j]);
        } else {
#pragma unroll
            for (int j = 0; j < G; ++j) {
                int k = k0 + j;
                bool ok = k < p.d_count;
                const float* src = p.logits ? lg + (size_t)k * p.stride_d : p.planes[ok ? k : 0] + pix;
                x[j] = ok ? __ldg(src) : -INFINITY;
                d[j] = ok ? hyp_at<HYPS>(p, p.d_begin + k, pix, tap) : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            if (k0 + j < p.d_count) {
                float e = p.identity ? x[j] : expf(x[j]);   // adamvs.py:514, no max subtraction
                mx = (mx < e) ? e : mx;            // :516-517
                acc = __fadd_rn(__fmul_rn(d[j], e), acc);   // :522
                s = s + e;                         // :525
            }
        }
    }
    if (p.state) {
        p.state[pix] = s;
        p.state[(size_t)p.HW + pix] = acc;
        p.state[2 * (size_t)p.HW + pix] = mx;
    }
    if (p.finalize) {
        float tot = s + 1e-10f;                    // :527-529
        float depth = p.identity ? acc : acc / tot;
        p.depth[pix] = depth;
        p.conf[pix] = p.identity ? mx : mx / tot;
        emit_next(p, pix, depth);
    }
}

template <int HYPS>
static int launch_regress(const RegressParams& p, int softmax_mode, cudaStream_t stream) {
    dim3 grid((p.HW + 255) / 256);
    if (softmax_mode == D3D_SOFTMAX_STABLE) {
        // 16-byte loads over four pixels per thread where the layout allows it (uniform / per-pixel hypotheses)
        const bool vec = HYPS != D3D_HYPS_RESIZED && (p.HW & 3) == 0 && (p.stride_d & 3) == 0 && p.HW >= 4 &&
                         ((reinterpret_cast<uintptr_t>(p.logits) | reinterpret_cast<uintptr_t>(p.hyps)) & 15) == 0;
        if (vec) {
            constexpr int VH = HYPS == D3D_HYPS_RESIZED ? D3D_HYPS_UNIFORM : HYPS;
            const bool win4 = p.conf_mode != D3D_CONF_MAX_PROB, var = p.expvar != nullptr;
            const dim3 g4((p.HW / 4 + 63) / 64), g1((p.HW / 4 + 255) / 256);
            // (uniform hypotheses fit 85 registers: three CTAs per SM, 8 x 16 bytes in flight per thread = 98 KB of loads
            // outstanding per SM -- 0.108 ms at cfg2; 12 planes in flight at two CTAs per SM: 0.115 ms; four CTAs spill: 0.161 ms)
#define D3D_VEC(PARTS, G, GRID)                                                                              \
            do {                                                                                             \
                if (win4 && var) regress_softmax_vec_kernel<VH, PARTS, G, true, true><<<GRID, 256, 0, stream>>>(p);        \
                else if (win4) regress_softmax_vec_kernel<VH, PARTS, G, true, false><<<GRID, 256, 0, stream>>>(p);         \
                else if (var) regress_softmax_vec_kernel<VH, PARTS, G, false, true><<<GRID, 256, 0, stream>>>(p);          \
                else regress_softmax_vec_kernel<VH, PARTS, G, false, false><<<GRID, 256, 0, stream>>>(p);                  \
            } while (0)
            if (p.D >= 32) D3D_VEC(4, 8, g4);
            else D3D_VEC(1, 8, g1);
#undef D3D_VEC
        }
        // long sweeps: four threads per pixel (a quarter of the planes each); short ones: one thread per pixel
        else if (p.D >= 32) regress_softmax_kernel<HYPS, 4><<<dim3((p.HW + 63) / 64), 256, 0, stream>>>(p);
        else regress_softmax_kernel<HYPS, 1><<<grid, 256, 0, stream>>>(p);
    } else
        regress_rawexp_kernel<HYPS><<<grid, 256, 0, stream>>>(p);
    count_launch();
    return check_launch("regress_kernel");
}

}  // namespace d3d

using namespace d3d;

extern "C" int d3d_depth_regress(const D3dRegressArgs* a, void* cuda_stream) {
    if (!a) return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: args is NULL");
    if (a->struct_size != sizeof(D3dRegressArgs))
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: struct_size %u != %zu", a->struct_size,
                    sizeof(D3dRegressArgs));
    if (a->num_depth <= 0 || a->height <= 0 || a->width <= 0)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: non-positive extent D=%d H=%d W=%d", a->num_depth,
                    a->height, a->width);
    if ((long long)a->height * a->width > INT32_MAX)
        return fail(D3D_ERR_UNSUPPORTED, "d3d_depth_regress: H*W exceeds 2^31-1");
    if (!a->hyps) return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: hyps is NULL");
    if (!a->logits) {                                   // scattered planes: streaming flavours only
        if (a->softmax_mode == D3D_SOFTMAX_STABLE || a->d_count <= 0 || a->d_count > D3D_REGRESS_MAX_PLANES)
            return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: logits is NULL (logit_planes need RAW_EXP / NONE and "
                        "1 <= d_count <= %d)", D3D_REGRESS_MAX_PLANES);
        for (int k = 0; k < a->d_count; ++k)
            if (!a->logit_planes[k]) return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: logit_planes[%d] is NULL", k);
    }
    if (a->softmax_mode < D3D_SOFTMAX_STABLE || a->softmax_mode > D3D_SOFTMAX_NONE)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: unknown softmax_mode %d", a->softmax_mode);
    if (a->conf_mode != D3D_CONF_MAX_PROB && a->conf_mode != D3D_CONF_WINDOW4)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: unknown conf_mode %d", a->conf_mode);
    if (a->hyps_mode < D3D_HYPS_UNIFORM || a->hyps_mode > D3D_HYPS_RESIZED)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: unknown hyps_mode %d", a->hyps_mode);
    if (a->hyps_mode == D3D_HYPS_RESIZED && (a->hyps_height <= 0 || a->hyps_width <= 0))
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: resized hypotheses need hyps_height/width");
    int d_begin = a->d_begin, d_count = a->d_count <= 0 ? a->num_depth - a->d_begin : a->d_count;
    if (d_begin < 0 || d_count <= 0 || d_begin + d_count > a->num_depth)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: plane slice [%d,+%d) outside 0..%d", d_begin, d_count,
                    a->num_depth);
    const bool raw = a->softmax_mode != D3D_SOFTMAX_STABLE;
    const bool whole = d_begin == 0 && d_count == a->num_depth;
    if (!raw && !whole)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: plane slices need D3D_SOFTMAX_RAW_EXP");
    if (raw && a->conf_mode != D3D_CONF_MAX_PROB)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: RAW_EXP only defines the max-prob confidence");
    if (raw && !whole && !a->state)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: a plane slice needs the state buffer");
    if (raw && a->exp_variance)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: exp_variance needs D3D_SOFTMAX_STABLE");
    const int finalize = raw ? (a->finalize != 0) : 1;
    if (finalize && (!a->depth || !a->conf))
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: depth/conf output is NULL");
    if (a->next_num_depth > 0 && !a->next_hyps)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: next_hyps is NULL");
    if (a->next_num_depth == 1)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_regress: next_num_depth must be >= 2");

    RegressParams p;
    p.logits = a->logits; p.hyps = a->hyps; p.depth = a->depth; p.conf = a->conf; p.index = a->index;
    p.state = a->state; p.expvar = a->exp_variance; p.next_hyps = a->next_hyps;
    p.D = a->num_depth; p.H = a->height; p.W = a->width; p.HW = a->height * a->width;
    p.stride_d = a->logits_stride_d > 0 ? a->logits_stride_d : p.HW;
    p.d_begin = d_begin; p.d_count = d_count;
    p.conf_mode = a->conf_mode; p.hyps_mode = a->hyps_mode;
    p.hh = a->hyps_height; p.hw = a->hyps_width; p.finalize = finalize;
    p.identity = a->softmax_mode == D3D_SOFTMAX_NONE;
    p.next_nd = a->next_num_depth > 0 ? a->next_num_depth : 0;
    p.next_half_span = (float)((double)p.next_nd / 2.0 * a->next_interval);
    p.lamb = a->lamb;
    p.scale_h = p.hh > 0 ? (float)p.hh / (float)p.H : 1.f;
    p.scale_w = p.hw > 0 ? (float)p.hw / (float)p.W : 1.f;
    for (int k = 0; k < D3D_REGRESS_MAX_PLANES; ++k) p.planes[k] = (!a->logits && k < d_count) ? a->logit_planes[k] : nullptr;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    switch (a->hyps_mode) {
        case D3D_HYPS_UNIFORM: return launch_regress<D3D_HYPS_UNIFORM>(p, a->softmax_mode, stream);
        case D3D_HYPS_PER_PIXEL: return launch_regress<D3D_HYPS_PER_PIXEL>(p, a->softmax_mode, stream);
        default: return launch_regress<D3D_HYPS_RESIZED>(p, a->softmax_mode, stream);
    }
}

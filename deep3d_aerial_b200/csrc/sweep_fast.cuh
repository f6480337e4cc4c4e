// Fused plane sweep, production kernel ("variant 0").
//
// Decomposition (C = CPT*LPP channels; CPT = 4 or 8 channels per lane, LPP in {1,2,4,8} lanes per pixel):
//   lane   = (pixel q of the warp, channel group cg of CPT channels);  LPP lanes share a pixel
//   warp   = 32/LPP consecutive reference pixels x all channels
//   group  = LPP warps = 32 consecutive pixels (one 128-byte row segment per output channel)
//   CTA    = 8 warps = 8/LPP groups;  blockIdx.y = chunk of depth planes
// Every thread walks its depth planes in order and keeps, per source view, the 2x2 texel footprint of
// its CPT channels in registers (64 / 128 registers at 4 source views); a footprint is re-fetched only
// when floor(ix), floor(iy) move, so the bilinear gather runs out of registers for ~1/travel planes.
// CPT = 4 fits 128 registers per thread -> 16 resident warps per SM instead of 8 (the kernel is
// latency-bound at 8, profiles/).
//
// What makes it fast (numbers from tools/microbench.cu and the ncu captures under profiles/):
//  * the projection is computed ONCE per (pixel, view, plane): the LPP lanes of a pixel split the source
//    views (and, when there are more lanes than views, two consecutive planes) and publish
//    {fx, fy, fx*fy, key} through a per-warp double-buffered shared-memory table, one pass ahead;
//    consumers fetch it with one broadcast LDS.128 per view;
//  * bilinear interpolation as  A + fx*B + fy*C + fx*fy*D  (B=b-a, C=c-a, D=a-b-c+d, rebuilt whenever a
//    footprint arrives): 3 FMAs per channel instead of 4;
//  * packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2, new on sm_100) over channel pairs.  FFMA2 has the
//    FP32 pipe throughput of two FFMAs but takes ONE issue slot; the kernel is issue-bound, so this is
//    what makes room for the address, shared-memory and store instructions;
//  * the whole plane is ONE basic block: the four views' footprint keys are compared with a single
//    branch, the (rare) re-fetch is an opaque PTX block that updates the cache in place;
//  * stores: 32-byte row segments reach only 2.3 TB/s on B200, 128-byte ones 7.3 TB/s -> results go
//    through an XOR-swizzled [planes][C][32 px] shared-memory tile per group (conflict-free STS.32 in,
//    LDS.128 out) and leave as full 128-byte rows with STG.128; the LPP warps of a group meet at a named
//    barrier once per kTilePlanes planes (double-buffered tile).
#pragma once
#include "common.cuh"
#include "sweep_refetch.cuh"

namespace d3d {

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 splat(float a) { return make_float2(a, a); }

// key = [27:14] y0+4 | [13:0] x0+4  (floor corner of the 2x2 footprint)
constexpr float kMagic = 12582912.f;           // 1.5 * 2^23: float(kMagic + n) has bits 0x4B400000 + n
constexpr int kMagicBits = 0x4B400000;
constexpr int kTilePlanes = 8;                 // planes staged per group barrier

// Two quotients by the same divisor, correctly rounded in all but pathological cases: one MUFU.RCP,
// one Newton step on the reciprocal, and the residual correction  q' = q + (x - q*z) * r  that IEEE
// division itself ends with -- without its special-case path (z is a depth-like positive number
// here; garbage in gives garbage that the caller clamps out of bounds).
__device__ __forceinline__ void div2(float x, float y, float z, float& u, float& v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
    r = fmaf(fmaf(-z, r, 1.f), r, r);
    float qu = x * r, qv = y * r;
    u = fmaf(fmaf(-qu, z, x), r, qu);
    v = fmaf(fmaf(-qv, z, y), r, qv);
}

// Projection of one reference pixel onto one source view at one depth, with the fp32 operation order
// of the reference + CUDA-ATen (tools/diag_coords.py checked each step bit for bit on B200):
//   module.py:538      ray = rot @ [x,y,1]        fma(r2,1,fma(r1,y,r0*x))         (cuBLAS)
//   module.py:539-541  X = ray*d (rounded) + t (rounded)
//   module.py:542      u = X/Z  correctly rounded division
//   module.py:543      g = u * f32(1/((W-1)/2)) - 1          (ATen-CUDA multiplies by the reciprocal)
//   GridSampler.h:31   ix = ((g + 1) * 0.5) * (W-1)
template <bool kIeeeDiv>
__device__ __forceinline__ float4 project_frac(float rx, float ry, float rz, float tx, float ty, float tz, float d,
                                               const SweepParams& p) {
    float X = __fadd_rn(__fmul_rn(rx, d), tx);
    float Y = __fadd_rn(__fmul_rn(ry, d), ty);
    float Z = __fadd_rn(__fmul_rn(rz, d), tz);
    float u, v;
    if (kIeeeDiv) {
        u = __fdiv_rn(X, Z);
        v = __fdiv_rn(Y, Z);
    } else {
        div2(X, Y, Z, u, v);
    }
    float ix = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(u, p.inv_half_w), 1.f), 1.f), 0.5f), p.wm1);
    float iy = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(v, p.inv_half_h), 1.f), 1.f), 0.5f), p.hm1);
    ix = fminf(fmaxf(ix, -2.f), p.wm1 + 2.f);   // NaN -> -2: every corner out of bounds
    iy = fminf(fmaxf(iy, -2.f), p.hm1 + 2.f);
    // floor without the XU pipe: round to nearest through the magic constant, step down if above
    float mx = __fadd_rn(ix, kMagic), my = __fadd_rn(iy, kMagic);
    float fx0 = __fsub_rn(mx, kMagic), fy0 = __fsub_rn(my, kMagic);
    int xi = __float_as_int(mx) - kMagicBits, yi = __float_as_int(my) - kMagicBits;
    if (fx0 > ix) { fx0 -= 1.f; xi -= 1; }
    if (fy0 > iy) { fy0 -= 1.f; yi -= 1; }
    const float fx = __fsub_rn(ix, fx0), fy = __fsub_rn(iy, fy0);
    // key: packed floor corner; the re-fetch block derives the in-bounds mask of the four corners from it
    const unsigned key = ((unsigned)(yi + 4) << 14) | (unsigned)(xi + 4);
    return make_float4(fx, fy, fx * fy, __uint_as_float(key));
}

__device__ __forceinline__ void group_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int CPT, int NV, int LPP, int MODE, bool kIeeeDiv>
__global__ void __launch_bounds__(256, (CPT == 4 ? 2 : 1)) sweep_fast_kernel(const SweepParams p) {
    constexpr int PPW = 32 / LPP;                          // pixels per warp
    constexpr int PB = (LPP >= 2 * NV) ? 2 : 1;            // planes published per geometry pass
    constexpr int KV = (LPP >= NV) ? 1 : (NV + LPP - 1) / LPP;   // projections a lane owns per pass
    constexpr int NP = CPT / 2;                            // channel pairs per lane
    constexpr int C = CPT * LPP;
    constexpr int GROUPS = 8 / LPP;                        // 32-pixel groups per CTA
    constexpr int TILE = kTilePlanes * C * 32;             // floats per tile buffer
    constexpr int NRO = CPT / 4;                           // float4 each lane moves per plane at read-out
    extern __shared__ float4 smem4[];
    float4(*geo)[2][PB][NV][PPW] = reinterpret_cast<float4(*)[2][PB][NV][PPW]>(smem4);   // [8 warps]
    float* tile = reinterpret_cast<float*>(smem4 + 8 * 2 * PB * NV * PPW);               // [GROUPS][2][TILE]

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int cg = lane % LPP;
    const int q = lane / LPP;
    const int grp = warp / LPP, wq = warp % LPP;           // group of the CTA, warp within the group
    const int choff = cg * CPT;
    const long long grp_base = ((long long)blockIdx.x * GROUPS + grp) * 32;
    const long long pix_raw = grp_base + wq * PPW + q;
    const int pix = pix_raw < p.HW ? (int)pix_raw : p.HW - 1;    // clamp: the warp stays whole
    const int py = pix / p.W, px = pix - py * p.W;

    const int d0 = p.d_begin + blockIdx.y * p.d_chunk;
    const int d1 = min(d0 + p.d_chunk, p.d_end);
    if (d0 >= d1) return;

    // projection ownership: with LPP >= NV a lane owns (view cg % NV, plane offset cg / NV) of each pass;
    // otherwise it owns views cg, cg + LPP, ... of the pass's single plane
    const int po = (LPP >= NV) ? cg / NV : 0;
    const bool owner = (LPP >= NV) ? (cg < PB * NV) : true;
    float rx[KV], ry[KV], rz[KV], tx[KV], ty[KV], tz[KV];
#pragma unroll
    for (int k = 0; k < KV; ++k) {
        const int v = (LPP >= NV) ? cg % NV : min(cg + k * LPP, NV - 1);
        const float* m = p.pose + v * 16;
        rx[k] = fmaf(m[2], 1.f, fmaf(m[1], (float)py, m[0] * (float)px));
        ry[k] = fmaf(m[6], 1.f, fmaf(m[5], (float)py, m[4] * (float)px));
        rz[k] = fmaf(m[10], 1.f, fmaf(m[9], (float)py, m[8] * (float)px));
        if (p.rays) {   // the reference's own rot @ [x,y,1] (cuBLAS), whatever order it rounded in
            const float* rr = p.rays + (size_t)(v) * 3 * p.HW + pix;
            rx[k] = __ldg(rr); ry[k] = __ldg(rr + p.HW); rz[k] = __ldg(rr + 2 * (size_t)p.HW);
        }
        tx[k] = m[3]; ty[k] = m[7]; tz[k] = m[11];
    }

    float2 rf[NP], rf2[NP];
    {
        const float* t = p.feats + (size_t)pix * C + choff;
#pragma unroll
        for (int k = 0; k < CPT; k += 4) {
            float4 w = ldg4(t + k);
            rf[k / 2] = f2(w.x, w.y);
            rf[k / 2 + 1] = f2(w.z, w.w);
        }
#pragma unroll
        for (int j = 0; j < NP; ++j) rf2[j] = __fmul2_rn(rf[j], rf[j]);
    }

    float2 tex[NV][4][NP];      // per view: A, B, C, D of the current 2x2 footprint
    unsigned ckey[NV];
    const float* vbase[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        ckey[v] = 0xffffffffu;
        vbase[v] = p.feats + (size_t)(v + 1) * p.HW * C + choff;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < NP; ++j) tex[v][k][j] = f2(0.f, 0.f);
    }
    const int row_bytes = p.W * C * 4;

    const size_t hyp_stride = p.perpix ? (size_t)p.HW : 1;
    const float* hp = p.hyps + (p.perpix ? (size_t)pix : 0);

    // projections of this lane's share of planes [base, base + PB): pure arithmetic, no branches, so the
    // compiler can interleave this long dependent chain with the packed interpolation arithmetic
    auto project_mine = [&](float4 (&gn)[KV], float depth) {
#pragma unroll
        for (int k = 0; k < KV; ++k)
            gn[k] = project_frac<kIeeeDiv>(rx[k], ry[k], rz[k], tx[k], ty[k], tz[k], depth, p);
    };
    auto store_mine = [&](const float4 (&gn)[KV], int buf, int base) {
        if (owner && base + po < d1) {
#pragma unroll
            for (int k = 0; k < KV; ++k) {
                const int v = (LPP >= NV) ? cg % NV : cg + k * LPP;
                if (LPP >= NV || NV % LPP == 0 || v < NV) geo[warp][buf][po][v][q] = gn[k];
            }
        }
    };
    auto depth_of = [&](int plane) { return __ldg(hp + (size_t)min(plane, d1 - 1) * hyp_stride); };

    // prologue: geometry of the first pass
    float dnext = depth_of(d0 + po);
    {
        float4 gn[KV];
        project_mine(gn, dnext);
        store_mine(gn, 0, d0);
    }
    dnext = depth_of(d0 + PB + po);

    const float invV = 1.f / (float)(NV + 1);
    // staging tile: element (plane t, channel row r, pixel column c) lives at  t*C*32 + r*32 + (c ^ swz(r)),
    // swz(r) = PPW*(r/CPT) & 31, which spreads the LPP channel groups of one pixel over distinct banks and
    // keeps every aligned run of 4 pixels contiguous for the 16-byte read-out
    float* tbase = tile + (size_t)grp * 2 * TILE;
    const int col = wq * PPW + q;
    const int s0 = choff * 32 + (col ^ ((PPW * cg) & 31));        // this lane's rows: s0 + 32*k
    const int t_in_grp = wq * 32 + lane;
    int rrow[NRO], rc4[NRO], roff[NRO];
#pragma unroll
    for (int i = 0; i < NRO; ++i) {
        const int idx = t_in_grp + i * LPP * 32;
        rrow[i] = idx >> 3;
        rc4[i] = (idx & 7) * 4;
        roff[i] = rrow[i] * 32 + (rc4[i] ^ ((PPW * (rrow[i] / CPT)) & 31));
    }
    const bool vec_ok = ((p.out_sc | p.out_sd) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
    const bool fast_rows = vec_ok && grp_base + 32 <= p.HW;      // whole group in range, 16-byte stores

    // one staged plane (slot tp of tile buffer buf, depth plane `plane`) -> global, as 128-byte rows
    auto flush_plane = [&](int buf, int tp, int plane) {
        const float* t = tbase + (size_t)buf * TILE + tp * C * 32;
#pragma unroll
        for (int i = 0; i < NRO; ++i) {
            const long long gp = grp_base + rc4[i];
            float* o = p.out + (size_t)rrow[i] * p.out_sc + (size_t)(plane - p.d_begin) * p.out_sd + gp;
            const float4 w = *reinterpret_cast<const float4*>(t + roff[i]);
            if (fast_rows) {
                *reinterpret_cast<float4*>(o) = w;
            } else {
                if (gp < p.HW) o[0] = w.x;
                if (gp + 1 < p.HW) o[1] = w.y;
                if (gp + 2 < p.HW) o[2] = w.z;
                if (gp + 3 < p.HW) o[3] = w.w;
            }
        }
    };

    // Staging protocol: batch b (kTilePlanes planes) is written into tile buffer b&1; the group meets at a
    // named barrier when the batch is complete; the batch is then drained ONE plane per iteration while
    // batch b+1 is being computed into the other buffer (no store burst, and the barrier of batch b+1
    // orders those reads before the buffer is rewritten by batch b+2).
    int staged = 0, batch = 0;                       // planes in the current tile buffer, buffer parity
    int drain_left = 0, drain_plane = 0;             // planes of the previous batch still to be written out

    // one depth plane; with `ahead`, also this lane's projections for the next pass (table nbuf, planes from nbase)
    auto process = [&](int dd, const float4 (*gsrc)[PPW], bool ahead, int nbuf, int nbase) {
        float4 g[NV];
        unsigned moved = 0;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            g[v] = gsrc[v][q];
            moved |= __float_as_uint(g[v].w) ^ ckey[v];
        }
        if (moved) {                                 // some footprint moved: re-fetch those (in place)
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const unsigned key = __float_as_uint(g[v].w);
                refetch_footprint(tex[v], key, ckey[v], vbase[v], p.W - 1, p.H - 1, row_bytes, C * 4);
                ckey[v] = key;
            }
        }

        float4 gn[KV];
        if (ahead) project_mine(gn, dnext);

        float2 s[NP], sq[NP];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const float2 fx = splat(g[v].x), fy = splat(g[v].y), fxy = splat(g[v].z);
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                float2 o = __ffma2_rn(fx, tex[v][1][j], tex[v][0][j]);
                o = __ffma2_rn(fy, tex[v][2][j], o);
                o = __ffma2_rn(fxy, tex[v][3][j], o);
                if (MODE == D3D_AGG_VARIANCE) {
                    if (v == 0) {
                        s[j] = __fadd2_rn(rf[j], o);
                        sq[j] = __ffma2_rn(o, o, rf2[j]);
                    } else {
                        s[j] = __fadd2_rn(s[j], o);
                        sq[j] = __ffma2_rn(o, o, sq[j]);
                    }
                }
            }
        }

        float2 r[NP];
        if (MODE == D3D_AGG_VARIANCE) {
            const float2 ninv = splat(-invV), pinv = splat(invV);
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                float2 t = __fmul2_rn(s[j], ninv);                  // -sum/V
                float2 w = __ffma2_rn(t, s[j], sq[j]);               // sq - sum^2/V
                r[j] = __fmul2_rn(w, pinv);                          // sq/V - (sum/V)^2
            }
        }
        if (ahead) store_mine(gn, nbuf, nbase);
        if (LPP == 1 && CPT == 8) {                  // the warp already owns 32 consecutive pixels
            float* oplane = p.out + (size_t)(dd - p.d_begin) * p.out_sd;
            if (pix_raw < p.HW) {
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    oplane[(size_t)(2 * j) * p.out_sc + pix] = r[j].x;
                    oplane[(size_t)(2 * j + 1) * p.out_sc + pix] = r[j].y;
                }
            }
        } else {
            float* t = tbase + (size_t)batch * TILE + staged * C * 32;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                t[s0 + 64 * j] = r[j].x;
                t[s0 + 64 * j + 32] = r[j].y;
            }
            if (drain_left > 0) {                    // previous batch: one plane per iteration
                flush_plane(batch ^ 1, kTilePlanes - drain_left, drain_plane++);
                --drain_left;
            }
            if (++staged == kTilePlanes || dd + 1 == d1) {
                group_barrier(1 + grp, LPP * 32);
                if (dd + 1 == d1) {                  // last batch: nothing left to hide it behind
                    for (; drain_left > 0; --drain_left)
                        flush_plane(batch ^ 1, kTilePlanes - drain_left, drain_plane++);
                    for (int tp = 0; tp < staged; ++tp) flush_plane(batch, tp, dd + 1 - staged + tp);
                } else {
                    drain_left = kTilePlanes;
                    drain_plane = dd + 1 - kTilePlanes;
                }
                staged = 0;
                batch ^= 1;
            }
        }
    };

    int pass = 0;
    for (int dd = d0; dd < d1; dd += PB, ++pass) {
        __syncwarp();                                // table pass&1 is complete; the other one is free
        process(dd, geo[warp][pass & 1][0], true, (pass + 1) & 1, dd + PB);
        dnext = depth_of(dd + 2 * PB + po);
#pragma unroll
        for (int t = 1; t < PB; ++t)
            if (dd + t < d1) process(dd + t, geo[warp][pass & 1][t], false, 0, 0);
    }
}

template <int CPT, int NV, int LPP>
constexpr size_t sweep_fast_smem() {
    constexpr int PB = (LPP >= 2 * NV) ? 2 : 1;
    return (size_t)8 * 2 * PB * NV * (32 / LPP) * sizeof(float4) +
           ((LPP == 1 && CPT == 8) ? 0 : (size_t)(8 / LPP) * 2 * kTilePlanes * CPT * LPP * 32 * sizeof(float));
}

template <int CPT, int NV, int LPP, int MODE>
int launch_sweep_fast(const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee_div) {
    constexpr size_t smem = sweep_fast_smem<CPT, NV, LPP>();
    auto kern = ieee_div ? sweep_fast_kernel<CPT, NV, LPP, MODE, true> : sweep_fast_kernel<CPT, NV, LPP, MODE, false>;
    static SmemOptIn opted[2];                       // per instantiation, per division flavour
    if (int rc = opted[ieee_div].ensure(kern, smem)) return rc;
    kern<<<grid, 256, smem, stream>>>(p);
    count_launch();
    return check_launch("sweep_fast_kernel");
}

}  // namespace d3d

// Fused plane sweep, production kernel for 32-channel features -- four planes per projection pass.
//
// Decomposition as in sweep_lean.cuh (lane = pixel q of the warp x group cg of 4 channels; 8 lanes share a
// pixel, a warp owns 4 consecutive reference pixels, the 8 warps of a CTA own 32 pixels = one 128-byte row
// segment per output channel).  What changed, and why (profiles/ncu_r1l.txt: the FP32 arithmetic of the
// variance volume is 11.5 of 35 warp instructions per voxel; the projection was 7.9):
//
//   * the 8 lanes of a pixel cover one PASS of 4 planes x 4 views with ONE projection chain each: lane cg
//     projects view cg & 3 for the plane pair cg >> 2, the two planes packed in fp32x2 (FFMA2 / FADD2 / FMUL2,
//     FADD2.RM for the floor); every operation is the same IEEE operation on each half, so the coordinates stay
//     bit-identical to the reference chain (project_frac in sweep_util.cuh);
//   * the published footprint key is just the floor corner as it falls out of the round-down adds
//     ((y0 & 0xffff) << 16 | x0 & 0xffff, one PRMT); address arithmetic, the interior test and zeros padding
//     moved into the re-fetch block (refetch_tok, sweep_refetch.cuh), which runs for ~9 % of the (view, plane)s;
//   * a pass is also the staged batch: one mbarrier arrive / wait per four planes, no inner pass loop.
//
// The same kernel serves 16- and 8-channel features (cascade stages 2 and 3): LPP = 4 or 2 lanes per pixel,
// 8 or 16 pixels per warp, 64 or 128 pixels (a 256 / 512-byte row segment per channel) per CTA; a lane then
// runs 2 or 4 projection chains per pass.
#pragma once
#include "sweep_lean.cuh"

namespace d3d {

__device__ __forceinline__ float2 lds64(unsigned addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
constexpr bool kQuadThreeCtas = true;   // <= 2 source views: 32 footprint registers, three CTAs per SM
constexpr int kQuadPlanes = 4;       // planes per pass = planes per staged batch
constexpr int kQuadBuffers = 4;      // staging ring: batches
constexpr int kQuadLag = kQuadBuffers / 2;   // a batch is drained while the batch `lag` later is computed
__host__ __device__ constexpr int quad_dot_rows(int mode, int gs, int c) {
    return mode == D3D_AGG_PAIR_MEAN ? 4 : (gs == 4 ? c / 4 : c);
}

template <int NV, int TEXB, int V = 0>
struct RefetchTok {
    static __device__ __forceinline__ void run(float2 (&tex)[NV][4][2], unsigned (&ckey)[NV], const float4 (&g)[NV],
                                               const float* base, unsigned row_bytes, const int (&vt)[kMaxSrc + 1], int W, int H) {
        refetch_tok<V + 1, TEXB>(tex[V], ckey[V], __float_as_uint(g[V].w), base, row_bytes, vt[V + 1], W, H);
        RefetchTok<NV, TEXB, V + 1>::run(tex, ckey, g, base, row_bytes, vt, W, H);
    }
    // variance volume of the long sweeps: A = a - ref (the reference view then contributes nothing to sum / sum of squares)
    static __device__ __forceinline__ void run_shift(float2 (&tex)[NV][4][2], unsigned (&ckey)[NV], const float4 (&g)[NV],
                                                     const float* base, unsigned row_bytes, const int (&vt)[kMaxSrc + 1], int W, int H,
                                                     const float2 (&ref)[2]) {
        refetch_tok_shift<V + 1, TEXB>(tex[V], ckey[V], __float_as_uint(g[V].w), base, row_bytes, vt[V + 1], W, H, ref);
        RefetchTok<NV, TEXB, V + 1>::run_shift(tex, ckey, g, base, row_bytes, vt, W, H, ref);
    }
};
template <int NV, int TEXB>
struct RefetchTok<NV, TEXB, NV> {
    static __device__ __forceinline__ void run(float2 (&)[NV][4][2], unsigned (&)[NV], const float4 (&)[NV], const float*,
                                               unsigned, const int (&)[kMaxSrc + 1], int, int) {}
    static __device__ __forceinline__ void run_shift(float2 (&)[NV][4][2], unsigned (&)[NV], const float4 (&)[NV], const float*,
                                                     unsigned, const int (&)[kMaxSrc + 1], int, int, const float2 (&)[2]) {}
};

// The same in two phases (issue_tok / rebuild_tok): every moved view's loads are started before any view is rebuilt,
// so their latencies overlap instead of adding up.
template <int NV, int TEXB, int V = 0>
struct RefetchIssue {
    static __device__ __forceinline__ void run(float2 (&tex)[NV][4][2], unsigned (&ckey)[NV], unsigned& mv, const float4 (&g)[NV],
                                               const float* base, unsigned row_bytes, const int (&vt)[kMaxSrc + 1], int W, int H) {
        issue_tok<V + 1, TEXB>(tex[V], ckey[V], mv, __float_as_uint(g[V].w), base, row_bytes, vt[V + 1], W, H);
        RefetchIssue<NV, TEXB, V + 1>::run(tex, ckey, mv, g, base, row_bytes, vt, W, H);
    }
};
template <int NV, int TEXB>
struct RefetchIssue<NV, TEXB, NV> {
    static __device__ __forceinline__ void run(float2 (&)[NV][4][2], unsigned (&)[NV], unsigned&, const float4 (&)[NV],
                                               const float*, unsigned, const int (&)[kMaxSrc + 1], int, int) {}
};
template <int NV, int V = 0>
struct RefetchRebuild {
    static __device__ __forceinline__ void run(float2 (&tex)[NV][4][2], unsigned mv) {
        rebuild_tok<V + 1>(tex[V], mv);
        RefetchRebuild<NV, V + 1>::run(tex, mv);
    }
};
template <int NV>
struct RefetchRebuild<NV, NV> {
    static __device__ __forceinline__ void run(float2 (&)[NV][4][2], unsigned) {}
};

// Correlation volumes with cached corner dot products (kDot): per view the lane keeps PA..PD = sum over its channels of
// ref_c * {A_c, B_c, C_c, D_c} (refetch_dot, sweep_refetch.cuh) instead of the 16 coefficients themselves.
template <int NV, int TEXB, int V = 0>
struct RefetchDot {
    static __device__ __forceinline__ void run(float (&dot)[NV][4], unsigned (&ckey)[NV], const float4 (&g)[NV],
                                               const float* base, unsigned row_bytes, const int (&vt)[kMaxSrc + 1], int W, int H,
                                               const float (&ref)[4]) {
        refetch_dot<V + 1, TEXB>(dot[V], ckey[V], __float_as_uint(g[V].w), base, row_bytes, vt[V + 1], W, H, ref);
        RefetchDot<NV, TEXB, V + 1>::run(dot, ckey, g, base, row_bytes, vt, W, H, ref);
    }
};
template <int NV, int TEXB>
struct RefetchDot<NV, TEXB, NV> {
    static __device__ __forceinline__ void run(float (&)[NV][4], unsigned (&)[NV], const float4 (&)[NV], const float*, unsigned,
                                               const int (&)[kMaxSrc + 1], int, int, const float (&)[4]) {}
};

// MODE: D3D_AGG_VARIANCE, D3D_AGG_WEIGHTED_PRODUCT (both write C rows per plane) or D3D_AGG_GROUP_CORR
// (p.groups rows; a lane's 4 channels are whole groups, one group, or a slice of a group that spans lanes).
// D3D_AGG_PAIR_MEAN writes one row per source view (mean over all channels of ref * warped).
// GS: channels per group known at compile time (4 = the G=8 configuration of BASELINE.json), 0 = read p.groups.
// kSplit: two-phase re-fetch (short sweeps, where most planes re-fetch: cascade stages 1 and 2).
// kDot: the correlation modes (group-wise with >= 4 channels per group, pair mean) on cached corner dot products: what a
// plane costs a lane is 3 FMAs per view, whatever the group size (cfg4: 5.15 -> see DESIGN.md), and the footprint cache is
// 16 registers instead of 64, so three or four CTAs fit an SM.
template <int NV, int MODE, bool kIeeeDiv, bool kPerPix, int GS = 0, int LPP = 8, bool kSplit = false, bool kDot = false>
__global__ void __launch_bounds__(256, kDot ? 3 : (NV <= 2 && kQuadThreeCtas) ? 3 : 2)
sweep_quad_kernel(const SweepParams p) {
    static_assert(!kDot || ((MODE == D3D_AGG_GROUP_CORR || MODE == D3D_AGG_PAIR_MEAN) && LPP == 8 && !kSplit),
                  "dot-product cache: correlation volumes of 32-channel features, one-block re-fetch");
    constexpr int CPT = 4, PPW = 32 / LPP, NP = 2, C = CPT * LPP;
    // long variance sweeps of 32-channel features: footprints are kept relative to the reference texel (A = a - ref at
    // re-fetch time; variance is shift invariant), so the reference view costs nothing per plane: one packed add per
    // channel pair and plane less, 4 registers less (ref^2), and less cancellation than the reference's own form
    constexpr bool kShift = MODE == D3D_AGG_VARIANCE && !kSplit && LPP == 8 && !kIeeeDiv;
    constexpr int PIX = 8 * PPW;                           // pixels per CTA: 32, 64 or 128
    constexpr int JPL = 8 / LPP;                           // projection chains a lane runs per pass
    constexpr int NVL = JPL > 1 ? JPL / 2 : 1;             // distinct views among them
    constexpr int KT = kQuadPlanes, NBUF = kQuadBuffers;
    // Table layout.  LPP = 8 (kRec): one 64-byte RECORD per pixel and plane, {key[4 views] | fx[4] | fy[4] | fx*fy[4]} -- the
    // four values of a kind land in adjacent registers with one LDS.128 (what packs the correlation modes' arithmetic over
    // view pairs).  Otherwise: one 16-byte entry {fx, fy, fx*fy, key} per (view, pixel).
    constexpr bool kRec = LPP == 8;
    // fx * fy formed by the consumer instead of read from the record (see load_table): cfg4 3.34 -> 3.30 ms; the variance
    // sweep, whose FP32 pipe is the busier unit, loses by it (cfg2 5.19 -> 5.23 ms)
    constexpr bool kFxyLocal = kRec && kDot;
    constexpr int VS = kRec ? 4 : (NV <= 2 ? 2 : 4);       // view slots of the table
    constexpr unsigned GEO_PLANE = VS * PPW * 16;          // bytes: one plane's table of one warp
    constexpr unsigned GEO_BUF = KT * GEO_PLANE;
    // rows of a staged plane: the channels, or -- with the dot-product cache -- only the rows the mode writes (8 groups of
    // cfg4, the <= 4 pair volumes), which leaves room for a fourth CTA per SM
    constexpr int ROWS = kDot ? quad_dot_rows(MODE, GS, C) : C;
    constexpr unsigned TILE_PLANE = ROWS * PIX * 4;        // bytes: one staged plane (4 KB for ROWS = C, whatever LPP)
    static_assert(MODE != D3D_AGG_PAIR_MEAN || LPP == 8, "pair-mean volumes are built from 32-channel features");
    constexpr unsigned TILE_BUF = KT * TILE_PLANE;
    constexpr unsigned TILE_RING = NBUF * TILE_BUF;
    extern __shared__ float4 smem4[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int cg = lane % LPP;
    const int q = lane / LPP;
    const int choff = cg * CPT;
    const long long grp_base = (long long)blockIdx.x * PIX;

    const int d0 = p.d_begin + blockIdx.y * p.d_chunk;
    const int d1 = min(d0 + p.d_chunk, p.d_end);
    if (d0 >= d1) return;

    // ---- shared-memory map (bytes): mbarriers | projection tables | staging ring | hypotheses
    const unsigned bar0 = smem_u32(smem4);
    const unsigned sm0 = bar0 + 64;
    const unsigned geo_w = sm0 + warp * 2 * GEO_BUF;
    const unsigned tile_g = sm0 + 8 * 2 * GEO_BUF;
    const unsigned hyp_s = tile_g + TILE_RING;

    if (threadIdx.x < NBUF) mbar_init(bar0 + threadIdx.x * 8, 8);      // one arrival per warp
    if (!kPerPix) {                                        // fronto-parallel sweep: stage the chunk's depths
        const int n = d1 - d0 + kLeanHypPad;
        for (int i = threadIdx.x; i < n; i += 256) sts32(hyp_s + i * 4, __ldg(p.hyps + min(d0 + i, d1 - 1)));
    }
    __syncthreads();

    // ---- this lane's projection jobs.  Job ids 0..7 = (plane pair, view); lane cg runs ids cg, cg + LPP, ...:
    //   LPP = 8: one job, view cg & 3, pair cg >> 2;   LPP = 4: view cg, both pairs;   LPP = 2: views cg and cg + 2,
    //   both pairs.  Chain k of a lane: view slot k % NVL, pair (LPP == 8 ? cg >> 2 : k / NVL).
    float rx[NVL], ry[NVL], rz[NVL], tx[NVL], ty[NVL], tz[NVL];
    bool owner[NVL];
    int jview[NVL];
    float2 rf[NP], rf2[NP];
    const float* hp = p.hyps;                              // kPerPix: this pixel's hypotheses
    {
        const long long pix_raw = grp_base + warp * PPW + q;
        const int pix = pix_raw < p.HW ? (int)pix_raw : p.HW - 1;    // clamp: the warp stays whole
        const int py = pix / p.W, px = pix - py * p.W;
#pragma unroll
        for (int i = 0; i < NVL; ++i) {
            const int view = (LPP == 8) ? (cg & 3) : cg + i * LPP;
            owner[i] = view < NV;
            jview[i] = min(view, NV - 1);
            const float* m = p.pose + jview[i] * 16;
            rx[i] = fmaf(m[2], 1.f, fmaf(m[1], (float)py, m[0] * (float)px));
            ry[i] = fmaf(m[6], 1.f, fmaf(m[5], (float)py, m[4] * (float)px));
            rz[i] = fmaf(m[10], 1.f, fmaf(m[9], (float)py, m[8] * (float)px));
            if (p.rays) {   // the reference's own rot @ [x,y,1] (cuBLAS), whatever order it rounded in
                const float* rr = p.rays + (size_t)(jview[i]) * 3 * p.HW + pix;
                rx[i] = __ldg(rr); ry[i] = __ldg(rr + p.HW); rz[i] = __ldg(rr + 2 * (size_t)p.HW);
            }
            tx[i] = m[3]; ty[i] = m[7]; tz[i] = m[11];
        }
        const float4 w = ldg4(p.feats + ((size_t)p.view_tex[0] + pix) * C + choff);
        rf[0] = f2(w.x, w.y);
        rf[1] = f2(w.z, w.w);
        rf2[0] = __fmul2_rn(rf[0], rf[0]);                 // variance: ref^2 once per pixel, not once per plane
        rf2[1] = __fmul2_rn(rf[1], rf[1]);
        if (kPerPix) hp = p.hyps + (size_t)pix + (size_t)d0 * p.HW;
    }
    const int pp0 = (LPP == 8) ? (cg >> 2) : 0;            // plane pair of chain 0 (chains k >= NVL: pair 1)

    float2 tex[NV][4][NP];      // per view: A, B, C, D of the current 2x2 footprint
    float dotc[NV][4];          // kDot: their dot products with the reference texel instead
    unsigned ckey[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        ckey[v] = 0x7fff7fffu;  // no footprint has this corner
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dotc[v][k] = 0.f;
#pragma unroll
            for (int j = 0; j < NP; ++j) tex[v][k][j] = f2(0.f, 0.f);
        }
    }
    const float rfs[4] = {rf[0].x, rf[0].y, rf[1].x, rf[1].y};
    const float* feats_c = p.feats + choff;
    const unsigned row_bytes = (unsigned)p.W * (unsigned)(C * 4);

    // ---- running shared-memory addresses
    unsigned gr = geo_w + q * (kRec ? 64 : 16);                            // read: + t*GEO_PLANE (+ v*PPW*16 | + field*16)
    // write (other buffer): chain 0's entry; chain k (same view, LPP = 4) is 2*k planes further
    static_assert(NVL == 1, "one view per lane: LPP = 8 or 4");
    unsigned gw = geo_w + GEO_BUF + 2 * ((LPP == 8) ? (cg >> 2) : 0) * GEO_PLANE +
                  (kRec ? q * 64 + jview[0] * 4 : q * 16 + jview[0] * PPW * 16);
    unsigned gflip = GEO_BUF;                                              // +/- distance between the buffers
    // output rows of this lane: variance / weighted product / groups of 1 channel: its 4 channels; groups of 2:
    // two rows; wider groups: one row, written by the first lane of the group.  Row r of a staged plane is
    // swizzled by the lane that owns it (4 * (r / rows per lane)), which keeps the staging stores conflict free
    // and every aligned run of 4 pixels contiguous for the 16-byte read-out.
    const int gs = (MODE != D3D_AGG_GROUP_CORR) ? 1 : (GS ? GS : C / p.groups);   // channels per output row
    const int rpl_log2 = gs >= 4 ? 0 : (gs == 2 ? 1 : 2);                  // log2(rows per lane)
    // (pair mean: after the reduce-scatter over the 8 lanes of a pixel, lane cg holds source view cg >> 1)
    const int row0 = (MODE == D3D_AGG_PAIR_MEAN) ? (cg >> 1) : gs >= 4 ? choff / gs : (choff >> (gs == 2 ? 1 : 0));
    const int n_rows = (MODE == D3D_AGG_GROUP_CORR) ? p.groups : (MODE == D3D_AGG_PAIR_MEAN) ? NV : C;
    // staged plane: [rows][PIX px]; the pixel column is swizzled inside its aligned group of 32 by the lane
    // that owns the row (PPW * lane-in-pixel), which spreads a warp's stores over all 32 banks
    const unsigned swz_w = (MODE == D3D_AGG_PAIR_MEAN) ? (unsigned)((PPW * (cg >> 1) * 2) & 31)
                                                       : (unsigned)((PPW * (row0 >> rpl_log2)) & 31);
    unsigned tw = tile_g + (row0 * PIX + ((warp * PPW + q) ^ swz_w)) * 4;  // + k*PIX*4 per row
    unsigned dr;                                           // drain: staged row chunk this lane moves
    // plane stride of the volume in bytes, opaque to the compiler: written as `optr += p.out_sd` the advance is re-derived
    // from the uniform register every plane (two moves, LEA, LEA.HI.X and another move: 5 instructions; now IADD3 + IADD3.X)
    unsigned long long sd_bytes;
    asm volatile("shl.b64 %0, %1, 2;" : "=l"(sd_bytes) : "l"(p.out_sd));
    float* optr;                                           // drain: where it goes
    bool drain_row;                                        // fewer than 32 rows (group-wise correlation)
    {
        constexpr int CPR = PIX / 4;                       // 16-byte chunks per staged row
        const int row = threadIdx.x / CPR, c4 = (threadIdx.x % CPR) * 4;
        const unsigned swz_r = (MODE == D3D_AGG_PAIR_MEAN) ? (unsigned)((PPW * row * 2) & 31)
                                                           : (unsigned)((PPW * (row >> rpl_log2)) & 31);
        dr = tile_g + (row * PIX + (c4 ^ swz_r)) * 4;
        optr = p.out + ((long long)row * p.out_sc + (long long)(d0 - p.d_begin) * p.out_sd + grp_base + c4);
        drain_row = row < n_rows;
        // H*W is a multiple of PIX (checked by the launcher): every CTA owns whole row segments
    }
    float wt[NV];                                          // weighted product: this pixel's view weights
    float winv = 0.f;
    if (MODE == D3D_AGG_WEIGHTED_PRODUCT) {
        const long long pix_raw = grp_base + warp * PPW + q;
        const int pix = pix_raw < p.HW ? (int)pix_raw : p.HW - 1;
        float wsum = p.eps_num ? 0.f : 1e-5f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            wt[v] = __ldg(p.weights + (size_t)v * p.HW + pix);
            wsum += wt[v];                                 // adamvs.py:494,506 accumulation order
        }
        winv = __frcp_rn(wsum);
    }
    const float gscale = 1.f / ((float)gs * (float)NV);    // group-wise correlation: mean over group and views

    // ---- hypotheses of the planes this lane projects in the next pass: both plane pairs unless LPP = 8
    unsigned hs = hyp_s + 2 * pp0 * 4;                     // !kPerPix: running shared-memory address
    int hplane = d0 + 2 * pp0;                             // kPerPix: plane `hp` points at
    if (kPerPix) hp += (size_t)(2 * pp0) * p.HW;
    auto next_depths = [&](float2 (&d)[2]) {
        if (kPerPix) {
            const size_t hw = (size_t)p.HW;
#pragma unroll
            for (int k = 0; k < (LPP == 8 ? 2 : 4); ++k) {
                const int pl = min(hplane + k, d1 - 1);    // planes past the end repeat the last one
                const float v = __ldg(hp + (long long)(pl - hplane) * (long long)hw);
                if (k == 0) d[0].x = v; else if (k == 1) d[0].y = v; else if (k == 2) d[1].x = v; else d[1].y = v;
            }
            hp += (size_t)KT * hw;
            hplane += KT;
        } else {
            if (LPP == 8) {
                d[0] = lds64(hs);
            } else {
                const float4 v = lds128(hs);
                d[0] = f2(v.x, v.y);
                d[1] = f2(v.z, v.w);
            }
            hs += KT * 4;
        }
    };

    // Packed projection of the lane's view at two depths; every packed operation is the reference's IEEE
    // operation on each half (see project_frac in sweep_util.cuh for the order and why it is that order).
    auto project2 = [&](int i, float2 d, float4& ea, float4& eb) {
        // (nvcc contracts __fmul2_rn + __fadd2_rn into one FFMA2, which would round once where the reference
        // rounds twice: every add that follows a multiply is a scalar __fadd_rn)
        const float2 Xm = __fmul2_rn(splat(rx[i]), d), Ym = __fmul2_rn(splat(ry[i]), d), Zm = __fmul2_rn(splat(rz[i]), d);
        const float2 X = f2(__fadd_rn(Xm.x, tx[i]), __fadd_rn(Xm.y, tx[i]));
        const float2 Y = f2(__fadd_rn(Ym.x, ty[i]), __fadd_rn(Ym.y, ty[i]));
        const float2 Z = f2(__fadd_rn(Zm.x, tz[i]), __fadd_rn(Zm.y, tz[i]));
        float2 u, v;
        if (kIeeeDiv) {
            u = f2(__fdiv_rn(X.x, Z.x), __fdiv_rn(X.y, Z.y));
            v = f2(__fdiv_rn(Y.x, Z.x), __fdiv_rn(Y.y, Z.y));
        } else {                                           // div2() of sweep_util.cuh, both planes at once
            float2 r;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(Z.x));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(Z.y));
            const float2 nZ = neg2(Z);
            r = __ffma2_rn(__ffma2_rn(nZ, r, splat(1.f)), r, r);
            const float2 qu = __fmul2_rn(X, r), qv = __fmul2_rn(Y, r);
            u = __ffma2_rn(__ffma2_rn(nZ, qu, X), r, qu);
            v = __ffma2_rn(__ffma2_rn(nZ, qv, Y), r, qv);
        }
        float2 ix = __fmul2_rn(u, splat(p.inv_half_w));
        float2 iy = __fmul2_rn(v, splat(p.inv_half_h));
        ix = f2(__fsub_rn(ix.x, 1.f), __fsub_rn(ix.y, 1.f));
        iy = f2(__fsub_rn(iy.x, 1.f), __fsub_rn(iy.y, 1.f));
        ix = __fadd2_rn(ix, splat(1.f));
        iy = __fadd2_rn(iy, splat(1.f));
        ix = __fmul2_rn(ix, splat(0.5f));
        iy = __fmul2_rn(iy, splat(0.5f));
        ix = __fmul2_rn(ix, splat(p.wm1));
        iy = __fmul2_rn(iy, splat(p.hm1));
        const float xhi = p.wm1 + 2.f, yhi = p.hm1 + 2.f;
        ix = f2(fminf(fmaxf(ix.x, -2.f), xhi), fminf(fmaxf(ix.y, -2.f), xhi));   // NaN -> -2: out of bounds
        iy = f2(fminf(fmaxf(iy.x, -2.f), yhi), fminf(fmaxf(iy.y, -2.f), yhi));
        // floor: a round-down add against 1.5*2^23 leaves floor() in the low mantissa bits, exactly
        const float2 mx = fadd2_rd(ix, splat(kMagic)), my = fadd2_rd(iy, splat(kMagic));
        const float2 fx = __fadd2_rn(ix, neg2(__fadd2_rn(mx, splat(-kMagic))));
        const float2 fy = __fadd2_rn(iy, neg2(__fadd2_rn(my, splat(-kMagic))));
        const float2 fxy = __fmul2_rn(fx, fy);
        const unsigned ka = __byte_perm(__float_as_uint(mx.x), __float_as_uint(my.x), 0x5410);
        const unsigned kb = __byte_perm(__float_as_uint(mx.y), __float_as_uint(my.y), 0x5410);
        ea = make_float4(fx.x, fy.x, fxy.x, __uint_as_float(ka));
        eb = make_float4(fx.y, fy.y, fxy.y, __uint_as_float(kb));
    };

    // chain k of the lane's pass: view slot k % NVL, plane pair (LPP == 8 ? pp0 : k / NVL); its two table entries
    // go to the buffer at `base`.  In the sweep, chain k runs inside plane k of the pass (interleaved with that
    // plane's arithmetic), so every plane carries at most one chain.
    auto store_chain = [&](int k, const float4& ea, const float4& eb, unsigned base) {
        if (owner[0]) {
            const unsigned a = base + (LPP == 8 ? 0 : 2 * k) * GEO_PLANE;
            if constexpr (kRec) {                    // this view's word of the four fields of the pixel's record
                sts32(a, ea.w); sts32(a + 16, ea.x); sts32(a + 32, ea.y); sts32(a + 48, ea.z);
                sts32(a + GEO_PLANE, eb.w); sts32(a + GEO_PLANE + 16, eb.x); sts32(a + GEO_PLANE + 32, eb.y);
                sts32(a + GEO_PLANE + 48, eb.z);
            } else {
                sts128(a, ea);
                sts128(a + GEO_PLANE, eb);
            }
        }
    };
    auto project_chain = [&](int k, const float2 (&d)[2], float4& ea, float4& eb) {
        project2(k % NVL, d[LPP == 8 ? 0 : k / NVL], ea, eb);
    };
    // ---- prologue: projections of the first pass into table buffer 0
    float2 dnext[2];
    next_depths(dnext);
#pragma unroll
    for (int k = 0; k < JPL; ++k) {
        float4 ea, eb;
        project_chain(k, dnext, ea, eb);
        store_chain(k, ea, eb, gw - GEO_BUF);
    }
    next_depths(dnext);

    const float invV = 1.f / (float)(NV + 1);
    const float2 ninv = splat(-invV), pinv = splat(invV);

    // Staging protocol (see sweep_lean.cuh): batch n is computed into ring slot n % NBUF and drained, one plane
    // per iteration, while batch n+2 is computed; the drain pointers simply keep running.
    unsigned bar_c = bar0, bar_d = bar0, par_d = 0;
    int slot_c = 0, slot_d = 0;

    auto drain_one = [&]() {
        // the volume is write-once: keep it out of L1, which holds the texels the re-fetches hit
        if ((MODE != D3D_AGG_GROUP_CORR && MODE != D3D_AGG_PAIR_MEAN) || drain_row) {
            const float4 w = lds128(dr);             // (rows past the mode's own are not even staged when ROWS < C)
            asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(optr), "f"(w.x), "f"(w.y),
                         "f"(w.z), "f"(w.w) : "memory");
        }
        dr += TILE_PLANE;
        optr = reinterpret_cast<float*>(reinterpret_cast<unsigned long long>(optr) + sd_bytes);
    };
    // Long sweeps (one-block re-fetch) read the staged chunk at the top of a plane and store it at the bottom, so the
    // store never waits on shared memory (cfg2: 5.555 -> 5.49 ms); the short-sweep flavour has no registers to
    // spare for it (stage-2 shape: 1.81 -> 1.90 ms with it) and the group-wise volume drains a quarter of the rows
    // (cfg4: 5.15 -> 5.26 ms with it): both keep the back-to-back form.
    constexpr bool kEarlyDrain = !kSplit && MODE == D3D_AGG_VARIANCE;
    float4 dw;                                       // the staged chunk this plane's iteration writes out
    auto drain_load = [&](int t) {                   // top of a plane: the LDS has the whole plane to land
        dw = lds128(dr + t * TILE_PLANE);            // (an immediate offset; `dr` advances once per batch)
    };
    auto drain_store = [&]() {                       // bottom of the plane
        if ((MODE != D3D_AGG_GROUP_CORR && MODE != D3D_AGG_PAIR_MEAN) || drain_row)
            asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(optr), "f"(dw.x), "f"(dw.y),
                         "f"(dw.z), "f"(dw.w) : "memory");
        optr = reinterpret_cast<float*>(reinterpret_cast<unsigned long long>(optr) + sd_bytes);
    };
    auto begin_drain = [&]() {
        mbar_wait(bar_d, par_d);
        bar_d += 8;
        if (++slot_d == NBUF) { slot_d = 0; bar_d = bar0; par_d ^= 1; }
    };

    float4 g[NV];
    // record layout, long sweeps: the keys of the NEXT plane are read right after this plane's move test (they are all that
    // test needs, and nothing between the two tests writes them), the fractions after the arithmetic as before: the test at
    // the top of a plane no longer waits on the table read (short-scoreboard stalls were 7 % of the samples)
    constexpr bool kEarlyKeys = kRec && !kSplit && !kDot && MODE == D3D_AGG_VARIANCE;
    float4 knext;
    auto load_table = [&](unsigned base) {
        if constexpr (kRec) {
            const float4 K = kEarlyKeys ? knext : lds128(base);
            const float4 FX = lds128(base + 16), FY = lds128(base + 32);
            // fx * fy: read from the record, or -- where the FP32 pipe has room and the shared-memory data path does not
            // (the correlation modes: 16 of a warp's ~20 wavefronts per plane are these table reads) -- formed again
            // here, the same IEEE product of the same two operands
            float4 FXY;
            if constexpr (kFxyLocal) {
                const float2 lo = __fmul2_rn(f2(FX.x, FX.y), f2(FY.x, FY.y)), hi = __fmul2_rn(f2(FX.z, FX.w), f2(FY.z, FY.w));
                FXY = make_float4(lo.x, lo.y, hi.x, hi.y);
            } else {
                FXY = lds128(base + 48);
            }
            const float kk[4] = {K.x, K.y, K.z, K.w}, fx[4] = {FX.x, FX.y, FX.z, FX.w}, fy[4] = {FY.x, FY.y, FY.z, FY.w},
                        fxy[4] = {FXY.x, FXY.y, FXY.z, FXY.w};
#pragma unroll
            for (int v = 0; v < NV; ++v) g[v] = make_float4(fx[v], fy[v], fxy[v], kk[v]);     // register renaming only
        } else {
#pragma unroll
            for (int v = 0; v < NV; ++v) g[v] = lds128(base + v * PPW * 16);
        }
    };
    __syncwarp();                                    // table buffer 0 is complete
    if (kEarlyKeys) knext = lds128(gr);
    load_table(gr);

    int n = 0;
#pragma unroll 1
    for (int b0 = d0; b0 < d1; b0 += KT, ++n) {
        const bool draining = n >= kQuadLag;         // every batch but the last is full
        if (draining) begin_drain();
#pragma unroll
        for (int t = 0; t < KT; ++t) {
            if (kEarlyDrain && draining) drain_load(t);
            unsigned moved = 0;
#pragma unroll
            for (int v = 0; v < NV; ++v) moved |= __float_as_uint(g[v].w) ^ ckey[v];
            if (moved) {                             // some footprint moved: re-fetch those (in place)
                if constexpr (kDot) {
                    RefetchDot<NV, C * 4>::run(dotc, ckey, g, feats_c, row_bytes, p.view_tex, p.W, p.H, rfs);
                } else if constexpr (kSplit) {
                    unsigned mv = 0;
                    RefetchIssue<NV, C * 4>::run(tex, ckey, mv, g, feats_c, row_bytes, p.view_tex, p.W, p.H);
                    RefetchRebuild<NV>::run(tex, mv);
                } else if constexpr (kShift) {
                    RefetchTok<NV, C * 4>::run_shift(tex, ckey, g, feats_c, row_bytes, p.view_tex, p.W, p.H, rf);
                } else {
                    RefetchTok<NV, C * 4>::run(tex, ckey, g, feats_c, row_bytes, p.view_tex, p.W, p.H);
                }
            }
            if constexpr (kEarlyKeys) {
                if (t + 1 < KT) {
                    knext = lds128(gr + (t + 1) * GEO_PLANE);
                } else {                             // last plane of the pass: the other table (complete since plane JPL - 1)
                    __syncwarp();
                    knext = lds128(gr + gflip);
                }
            }
            float4 ea, eb;
            if (t < JPL) project_chain(t, dnext, ea, eb);    // next pass, interleaved with the arithmetic

            float2 s[NP], sq[NP];                    // variance: sum, sum of squares; otherwise s = accumulator
            float pm[4] = {0.f, 0.f, 0.f, 0.f};      // pair mean: this lane's partial dot product per view
            float dsum = 0.f;                        // kDot, group-wise: sum over the views
            if constexpr (kDot) {
                // PA + fx*PB + fy*PC + fxy*PD per view, two views per packed operation (the record layout hands fx, fy, fxy
                // of views (0,1) and (2,3) over in adjacent registers)
#pragma unroll
                for (int v = 0; v + 1 < NV; v += 2) {
                    float2 o = __ffma2_rn(f2(g[v].x, g[v + 1].x), f2(dotc[v][1], dotc[v + 1][1]), f2(dotc[v][0], dotc[v + 1][0]));
                    o = __ffma2_rn(f2(g[v].y, g[v + 1].y), f2(dotc[v][2], dotc[v + 1][2]), o);
                    o = __ffma2_rn(f2(g[v].z, g[v + 1].z), f2(dotc[v][3], dotc[v + 1][3]), o);
                    pm[v] = o.x;
                    pm[v + 1] = o.y;
                }
                if (NV & 1) {
                    constexpr int v = NV - 1;
                    pm[v] = fmaf(g[v].z, dotc[v][3], fmaf(g[v].y, dotc[v][2], fmaf(g[v].x, dotc[v][1], dotc[v][0])));
                }
                dsum = pm[0];
#pragma unroll
                for (int v = 1; v < NV; ++v) dsum += pm[v];
            }
#pragma unroll
            for (int v = 0; v < (kDot ? 0 : NV); ++v) {
                const float2 fx = splat(g[v].x), fy = splat(g[v].y), fxy = splat(g[v].z);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    float2 o = __ffma2_rn(fx, tex[v][1][j], tex[v][0][j]);
                    o = __ffma2_rn(fy, tex[v][2][j], o);
                    o = __ffma2_rn(fxy, tex[v][3][j], o);
                    if (MODE == D3D_AGG_VARIANCE) {
                        if (v == 0 && kShift) {
                            s[j] = o;
                            sq[j] = __fmul2_rn(o, o);
                        } else if (v == 0) {
                            s[j] = __fadd2_rn(rf[j], o);
                            sq[j] = __ffma2_rn(o, o, rf2[j]);
                        } else {
                            s[j] = __fadd2_rn(s[j], o);
                            sq[j] = __ffma2_rn(o, o, sq[j]);
                        }
                    } else if (MODE == D3D_AGG_GROUP_CORR) {           // sum_v ref * warped_v
                        s[j] = (v == 0) ? __fmul2_rn(rf[j], o) : __ffma2_rn(rf[j], o, s[j]);
                    } else if (MODE == D3D_AGG_PAIR_MEAN) {            // per view: sum_c ref * warped_v
                        s[0] = (j == 0) ? __fmul2_rn(rf[j], o) : __ffma2_rn(rf[j], o, s[0]);
                        if (j == NP - 1) pm[v] = s[0].x + s[0].y;
                    } else {                                           // sum_v (warped_v * ref) * weight_v
                        s[j] = __ffma2_rn(__fmul2_rn(o, rf[j]), splat(wt[v]),
                                          v == 0 ? splat(p.eps_num ? 1e-5f : 0.f) : s[j]);
                    }
                }
            }
            if (t < JPL) store_chain(t, ea, eb, gw);
            if (t == JPL - 1) next_depths(dnext);
            if (t + 1 < KT) {
                load_table(gr + (t + 1) * GEO_PLANE);
            } else {                                 // last plane of the pass: swap the table buffers
                if (!kEarlyKeys) __syncwarp();       // the other table is complete; this one is free
                gr += gflip;
                gw -= gflip;
                gflip = 0u - gflip;
                load_table(gr);
            }
            if (MODE == D3D_AGG_VARIANCE) {
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const float2 tneg = __fmul2_rn(s[j], ninv);            // -sum/V
                    const float2 w = __ffma2_rn(tneg, s[j], sq[j]);        // sq - sum^2/V
                    const float2 r = __fmul2_rn(w, pinv);                  // sq/V - (sum/V)^2
                    sts32(tw + (2 * j) * PIX * 4 + t * TILE_PLANE, r.x);
                    sts32(tw + (2 * j + 1) * PIX * 4 + t * TILE_PLANE, r.y);
                }
            } else if (MODE == D3D_AGG_WEIGHTED_PRODUCT) {
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const float2 r = __fmul2_rn(s[j], splat(winv));
                    sts32(tw + (2 * j) * PIX * 4 + t * TILE_PLANE, r.x);
                    sts32(tw + (2 * j + 1) * PIX * 4 + t * TILE_PLANE, r.y);
                }
            } else if (MODE == D3D_AGG_PAIR_MEAN) {
                // reduce-scatter over the 8 lanes of the pixel: 4 partial sums -> lane cg ends with the total of
                // view cg >> 1 (4 shuffles instead of a 3-step butterfly per view)
                const bool hi4 = (cg & 4) != 0, hi2 = (cg & 2) != 0;
                float a0 = hi4 ? pm[2] : pm[0], a1 = hi4 ? pm[3] : pm[1];          // keep views (0,1) or (2,3)
                const float b0 = hi4 ? pm[0] : pm[2], b1 = hi4 ? pm[1] : pm[3];    // send the other two
                a0 += __shfl_xor_sync(0xffffffffu, b0, 4);
                a1 += __shfl_xor_sync(0xffffffffu, b1, 4);
                float c0 = hi2 ? a1 : a0;
                const float c1 = hi2 ? a0 : a1;
                c0 += __shfl_xor_sync(0xffffffffu, c1, 2);
                c0 += __shfl_xor_sync(0xffffffffu, c0, 1);
                if (!(cg & 1) && (cg >> 1) < NV) sts32(tw + t * TILE_PLANE, c0 * (1.f / (float)C));
            } else {                                 // group-wise correlation: mean over the group and the views
                if (gs == 1) {
#pragma unroll
                    for (int j = 0; j < NP; ++j) {
                        sts32(tw + (2 * j) * PIX * 4 + t * TILE_PLANE, s[j].x * gscale);
                        sts32(tw + (2 * j + 1) * PIX * 4 + t * TILE_PLANE, s[j].y * gscale);
                    }
                } else if (gs == 2) {
                    sts32(tw + t * TILE_PLANE, (0.f + s[0].x + s[0].y) * gscale);
                    sts32(tw + PIX * 4 + t * TILE_PLANE, (0.f + s[1].x + s[1].y) * gscale);
                } else {
                    float a = kDot ? dsum : ((s[0].x + s[0].y) + s[1].x) + s[1].y;   // the lane's 4 channels, in order
                    for (int o = 1; o < gs / 4; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                    if ((cg & (gs / 4 - 1)) == 0) sts32(tw + t * TILE_PLANE, a * gscale);
                }
            }
            if (draining) {
                if (kEarlyDrain) drain_store(); else drain_one();
            }
        }
        // batch n is staged in ring slot slot_c: announce it (one arrival per warp)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_c);
        bar_c += 8;
        tw += TILE_BUF;
        if (++slot_c == NBUF) { slot_c = 0; bar_c = bar0; tw -= TILE_RING; }
        if (kEarlyDrain && draining) dr += KT * TILE_PLANE;
        if (draining && slot_d == 0) dr -= TILE_RING;   // the drained slot was the last of the ring
    }
    // the sweep is over: the last two batches have nothing left to hide behind
    for (int m = max(n - kQuadLag, 0); m < n; ++m) {
        begin_drain();
        for (int k = min(KT, d1 - (d0 + m * KT)); k > 0; --k) drain_one();
        if (slot_d == 0) dr -= TILE_RING;
    }
}

template <int NV, int MODE, int LPP>
int launch_sweep_quad(const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee_div) {
    const size_t smem = 64 + (size_t)8 * 2 * kQuadPlanes * ((LPP == 8 || NV > 2) ? 4 : 2) * (32 / LPP) * 16 +
                        (size_t)kQuadBuffers * kQuadPlanes * 32 * 32 * 4 +
                        (p.perpix ? 0 : (size_t)(p.d_chunk + kLeanHypPad) * 4);
    if (smem > 110 * 1024) return -1;                // keep two CTAs per SM; absurd depth chunks go elsewhere
    const size_t smem_req = (p.flags & 8) ? 120 * 1024 : smem;   // DIAGNOSTIC (variant 24): one CTA per SM
    void (*kern)(const SweepParams);
    const int which = (ieee_div ? 2 : 0) + (p.perpix ? 1 : 0);
    const bool g8 = MODE == D3D_AGG_GROUP_CORR && LPP == 8 && p.groups == 8 && !ieee_div;   // groups of 4 = one lane
    // two-phase re-fetch for the short sweeps of the cascade (48 / 32 / 8 planes: a footprint lasts 1-3 planes, most
    // planes re-fetch: -13 % / -10 % on the stage-1 / stage-2 shapes); the one-block form for long sweeps (384 planes,
    // 9 % re-fetch rate: the two-phase form costs 1 % there).  A/B: variant 18 forces it, variant 32 forbids it.
    // correlation volumes whose groups span at least a lane's 4 channels (and the pair volumes): the dot-product cache;
    // A/B: variant 48 (flags bit 5) keeps the coefficient cache
    if constexpr ((MODE == D3D_AGG_GROUP_CORR || MODE == D3D_AGG_PAIR_MEAN) && LPP == 8) {
        const bool wide = MODE == D3D_AGG_PAIR_MEAN || (p.groups > 0 && p.C / p.groups >= 4);
        if (wide && !ieee_div && !(p.flags & 32)) {
            const size_t smem_dot = smem - (size_t)kQuadBuffers * kQuadPlanes * 32 * 4 * (32 - quad_dot_rows(MODE, g8 ? 4 : 0, 32));
            const size_t smem_req = (p.flags & 8) ? 120 * 1024 : smem_dot;
            void (*kd)(const SweepParams);
            if (g8) kd = p.perpix ? sweep_quad_kernel<NV, MODE, false, true, 4, LPP, false, true>
                                  : sweep_quad_kernel<NV, MODE, false, false, 4, LPP, false, true>;
            else    kd = p.perpix ? sweep_quad_kernel<NV, MODE, false, true, 0, LPP, false, true>
                                  : sweep_quad_kernel<NV, MODE, false, false, 0, LPP, false, true>;
            static SmemOptIn opted_dot[2][2];
            if (int rc = opted_dot[g8 ? 1 : 0][p.perpix ? 1 : 0].ensure(kd, smem_req)) return rc;
            kd<<<grid, 256, smem_req, stream>>>(p);
            count_launch();
            return check_launch("sweep_quad_kernel (dot cache)");
        }
    }
    const bool split = !(p.flags & 16) && ((p.flags & 2) != 0 || p.d_end - p.d_begin <= 128) && !ieee_div && !g8;
    if (split) {
        kern = p.perpix ? sweep_quad_kernel<NV, MODE, false, true, 0, LPP, true>
                        : sweep_quad_kernel<NV, MODE, false, false, 0, LPP, true>;
        static SmemOptIn opted_split[2];
        if (int rc = opted_split[p.perpix ? 1 : 0].ensure(kern, smem_req)) return rc;
        kern<<<grid, 256, smem_req, stream>>>(p);
        count_launch();
        return check_launch("sweep_quad_kernel");
    }
    switch (which) {
        case 0: kern = g8 ? sweep_quad_kernel<NV, MODE, false, false, (LPP == 8 ? 4 : 0), LPP>
                          : sweep_quad_kernel<NV, MODE, false, false, 0, LPP>; break;
        case 1: kern = g8 ? sweep_quad_kernel<NV, MODE, false, true, (LPP == 8 ? 4 : 0), LPP>
                          : sweep_quad_kernel<NV, MODE, false, true, 0, LPP>; break;
        case 2: kern = sweep_quad_kernel<NV, MODE, true, false, 0, LPP>; break;
        default: kern = sweep_quad_kernel<NV, MODE, true, true, 0, LPP>; break;
    }
    static SmemOptIn opted[2][4];                    // per instantiation, group flavour and kernel flavour
    if (int rc = opted[g8 ? 1 : 0][which].ensure(kern, smem_req)) return rc;
    kern<<<grid, 256, smem_req, stream>>>(p);
    count_launch();
    return check_launch("sweep_quad_kernel");
}

template <int MODE, int LPP>
int sweep_quad_by_views(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee) {
    switch (nv) {
        case 1: return launch_sweep_quad<1, MODE, LPP>(p, grid, stream, ieee);
        case 2: return launch_sweep_quad<2, MODE, LPP>(p, grid, stream, ieee);
        case 3: return launch_sweep_quad<3, MODE, LPP>(p, grid, stream, ieee);
        case 4: return launch_sweep_quad<4, MODE, LPP>(p, grid, stream, ieee);
        default: return -1;
    }
}

// returns -1 when the shape is not covered (the caller falls back to sweep_lean / sweep_base)
template <int MODE>
int sweep_quad_dispatch(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee) {
    if (p.W > 16000 || p.H > 16000) return -1;       // 16-bit corner fields in the footprint key
    // rows leave as 16-byte chunks, every CTA owns whole row segments of 32 / 64 pixels
    if (((p.out_sc | p.out_sd) & 3) != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15) != 0) return -1;
    if (p.C == 32 && (p.HW & 31) == 0) return sweep_quad_by_views<MODE, 8>(nv, p, grid, stream, ieee);
    if constexpr (MODE != D3D_AGG_PAIR_MEAN && MODE != D3D_AGG_GROUP_CORR) {          // cascade stages 2 and 3
        if (p.C == 16 && (p.HW & 63) == 0) return sweep_quad_by_views<MODE, 4>(nv, p, grid, stream, ieee);
        // (8-channel features, LPP = 2: the projection chains outweigh the arithmetic, one CTA per SM, and stage 3
        // sweeps only 8 planes -- measured slower than sweep_base_kernel, 3.7 vs 2.9 ms on the cfg3 shape)
    }
    return -1;
}

}  // namespace d3d

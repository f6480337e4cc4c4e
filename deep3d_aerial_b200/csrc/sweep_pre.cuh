// Fused plane sweep for the long variance sweeps of 32-channel features: sweep_quad.cuh's formulation with the moved
// footprints PREFETCHED into shared memory by TMA a whole pass (4 planes) ahead of their use.
//
// What changes against sweep_quad_kernel (profiles/ncu_r1n.txt: 19 % of all stall samples sit on the global loads of
// the re-fetch block, which a warp issues and needs in the same breath):
//   * the lane that projects a (pixel, view, plane pair) for the NEXT pass also knows the floor corner of the plane
//     before (its own previous result, or its partner lane's: one SHFL per pass), so it sees every footprint move a pass
//     early.  For each move it takes a 512-byte slot of the warp's region for that pass and issues ONE
//     cp.async.bulk.tensor.4d: the 2 x 2 x 32-channel box of the texel tensor at (x0, y0, view); texels outside the
//     image arrive as zeros (the tensor map's out-of-bounds fill = grid_sample's zeros padding), so there is no border
//     case.  The table entry {fx, fy, fx*fy, info} carries the slot address instead of the corner key;
//   * the consuming lanes pick the footprint up with 4 LDS.128 (consume_ws_shift, sweep_refetch.cuh): no address
//     arithmetic, no global load, no per-view "current key" registers.  A warp whose region is full (7 slots per pass;
//     the cfg2 rig averages 4 moves per warp and pass) publishes the corner instead and the consumer falls back to the
//     global loads;
//   * the copies of a pass complete on an mbarrier private to the warp and the pass parity; the warp waits on it at
//     the top of the pass that consumes them -- a pass after they were issued;
//   * the staging ring is two batches deep (a batch is drained while the next one is computed), which pays for the slots.
// Everything else -- decomposition, packed projection, 3-FMA bilinear form relative to the reference texel, swizzled
// staging, 128-byte row stores -- is sweep_quad.cuh's.
#pragma once
#include "sweep_ws.cuh"

namespace d3d {

constexpr int kPreSlots = 7;            // footprint slots per warp and pass parity
constexpr bool kPreBulk = false;        // plain bulk copies (2 x 256 bytes) instead of one tensor-map box per footprint
constexpr unsigned kPreBarBytes = 256;  // ring barriers [2] at +0, per-warp copy barriers [8][2] at +64
constexpr unsigned kPreRegion = kPreSlots * 512;
constexpr unsigned kPreSlotBytes = 8 * 2 * kPreRegion;
constexpr unsigned kPreGeoPlane = 4 * 4 * 16;               // one plane's table of one warp: 4 views x 4 pixels
constexpr unsigned kPreGeoBuf = 4 * kPreGeoPlane;
constexpr unsigned kPreTableBytes = 8 * 2 * kPreGeoBuf;
constexpr unsigned kPreTilePlane = 32 * 32 * 4;
constexpr unsigned kPreTileBuf = 4 * kPreTilePlane;
constexpr unsigned kPreTileRing = 2 * kPreTileBuf;

template <int NV, bool kPerPix>
__global__ void __launch_bounds__(256, 2) sweep_pre_kernel(const SweepParams p, const __grid_constant__ CUtensorMap texmap) {
    constexpr int KT = 4, NBUF = 2, NP = 2, C = 32, PIX = 32, PPW = 4;
    extern __shared__ __align__(1024) float4 smem4[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int cg = lane & 7;
    const int q = lane >> 3;
    const int choff = cg * 4;
    const long long grp_base = (long long)blockIdx.x * PIX;

    const int d0 = p.d_begin + blockIdx.y * p.d_chunk;
    const int d1 = min(d0 + p.d_chunk, p.d_end);
    if (d0 >= d1) return;

    // ---- shared-memory map (bytes): mbarriers | footprint slots | projection tables | staging ring | hypotheses
    const unsigned bar0 = smem_u32(smem4);
    const unsigned wbar = bar0 + 64 + warp * 16;           // this warp's copy barriers, one per pass parity
    const unsigned slot_w = bar0 + kPreBarBytes + warp * 2 * kPreRegion;
    const unsigned geo_w = bar0 + kPreBarBytes + kPreSlotBytes + warp * 2 * kPreGeoBuf;
    const unsigned tile_g = bar0 + kPreBarBytes + kPreSlotBytes + kPreTableBytes;
    const unsigned hyp_s = tile_g + kPreTileRing;

    if (threadIdx.x < NBUF) mbar_init(bar0 + threadIdx.x * 8, 8);      // one arrival per warp
    else if (threadIdx.x >= 32 && threadIdx.x < 48) mbar_init(bar0 + 64 + (threadIdx.x - 32) * 8, 1);
    if (!kPerPix) {                                        // fronto-parallel sweep: stage the chunk's depths
        const int n = d1 - d0 + kLeanHypPad;
        for (int i = threadIdx.x; i < n; i += 256) sts32(hyp_s + i * 4, __ldg(p.hyps + min(d0 + i, d1 - 1)));
    }
    __syncthreads();

    // ---- this lane's projection job: view cg & 3, plane pair cg >> 2 of every pass
    float rx, ry, rz, tx, ty, tz;
    const int view = cg & 3;
    const bool owner = view < NV;
    const int jview = min(view, NV - 1);
    float2 rf[NP];
    const float* hp = p.hyps;                              // kPerPix: this pixel's hypotheses
    {
        const long long pix_raw = grp_base + warp * PPW + q;
        const int pix = pix_raw < p.HW ? (int)pix_raw : p.HW - 1;    // clamp: the warp stays whole
        const int py = pix / p.W, px = pix - py * p.W;
        const float* m = p.pose + jview * 16;
        rx = fmaf(m[2], 1.f, fmaf(m[1], (float)py, m[0] * (float)px));
        ry = fmaf(m[6], 1.f, fmaf(m[5], (float)py, m[4] * (float)px));
        rz = fmaf(m[10], 1.f, fmaf(m[9], (float)py, m[8] * (float)px));
        if (p.rays) {   // the reference's own rot @ [x,y,1] (cuBLAS), whatever order it rounded in
            const float* rr = p.rays + (size_t)jview * 3 * p.HW + pix;
            rx = __ldg(rr); ry = __ldg(rr + p.HW); rz = __ldg(rr + 2 * (size_t)p.HW);
        }
        tx = m[3]; ty = m[7]; tz = m[11];
        const float4 w = ldg4(p.feats + (size_t)pix * C + choff);
        rf[0] = f2(w.x, w.y);
        rf[1] = f2(w.z, w.w);
        if (kPerPix) hp = p.hyps + (size_t)pix + (size_t)d0 * p.HW;
    }
    const int pp0 = cg >> 2;

    float2 tex[NV][4][NP];      // per view: A - ref, B, C, D of the current 2x2 footprint
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < NP; ++j) tex[v][k][j] = f2(0.f, 0.f);
    const float* feats_c = p.feats + choff;
    const unsigned row_bytes = (unsigned)p.W * (unsigned)(C * 4);
    const unsigned lane_off = (unsigned)choff * 4u - 1u;   // slot address = (info - 1) + the lane's channel offset

    // ---- running shared-memory addresses (as in sweep_quad.cuh)
    unsigned gr = geo_w + q * 16;                                          // read: + t*GEO_PLANE + v*PPW*16
    unsigned gw = geo_w + kPreGeoBuf + q * 16 + 2 * pp0 * kPreGeoPlane + jview * PPW * 16;   // write: the other buffer
    unsigned gflip = kPreGeoBuf;
    unsigned tw = tile_g + (choff * PIX + ((warp * PPW + q) ^ ((4 * cg) & 31))) * 4;
    unsigned dr;
    float* optr;
    {
        const int row = threadIdx.x / 8, c4 = (threadIdx.x % 8) * 4;
        dr = tile_g + (row * PIX + (c4 ^ ((4 * (row >> 2)) & 31))) * 4;
        optr = p.out + ((long long)row * p.out_sc + (long long)(d0 - p.d_begin) * p.out_sd + grp_base + c4);
    }

    unsigned hs = hyp_s + 2 * pp0 * 4;
    int hplane = d0 + 2 * pp0;
    if (kPerPix) hp += (size_t)(2 * pp0) * p.HW;
    auto next_depths = [&](float2& d) {
        if (kPerPix) {
            const size_t hw = (size_t)p.HW;
            const int pa = min(hplane, d1 - 1), pb = min(hplane + 1, d1 - 1);
            d.x = __ldg(hp + (long long)(pa - hplane) * (long long)hw);
            d.y = __ldg(hp + (long long)(pb - hplane) * (long long)hw);
            hp += (size_t)KT * hw;
            hplane += KT;
        } else {
            d = lds64(hs);
            hs += KT * 4;
        }
    };

    unsigned klast = 0x7fff7fffu;      // floor corner of the last plane this lane projected (no footprint has this one)

    // One projection round: this lane's (view, plane pair) of the pass whose table is at `gwb`, slots at `region`,
    // copies completing on `cbar`.
    auto project_round = [&](float2 d, unsigned gwb, unsigned region, unsigned cbar) {
        const float2 Xm = __fmul2_rn(splat(rx), d), Ym = __fmul2_rn(splat(ry), d), Zm = __fmul2_rn(splat(rz), d);
        const float2 X = f2(__fadd_rn(Xm.x, tx), __fadd_rn(Xm.y, tx));
        const float2 Y = f2(__fadd_rn(Ym.x, ty), __fadd_rn(Ym.y, ty));
        const float2 Z = f2(__fadd_rn(Zm.x, tz), __fadd_rn(Zm.y, tz));
        float2 r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(Z.x));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(Z.y));
        const float2 nZ = neg2(Z);
        r = __ffma2_rn(__ffma2_rn(nZ, r, splat(1.f)), r, r);
        const float2 qu = __fmul2_rn(X, r), qv = __fmul2_rn(Y, r);
        const float2 u = __ffma2_rn(__ffma2_rn(nZ, qu, X), r, qu);
        const float2 v = __ffma2_rn(__ffma2_rn(nZ, qv, Y), r, qv);
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
        const float2 mx = fadd2_rd(ix, splat(kMagic)), my = fadd2_rd(iy, splat(kMagic));
        const float2 fx = __fadd2_rn(ix, neg2(__fadd2_rn(mx, splat(-kMagic))));
        const float2 fy = __fadd2_rn(iy, neg2(__fadd2_rn(my, splat(-kMagic))));
        const float2 fxy = __fmul2_rn(fx, fy);
        const unsigned ka = __byte_perm(__float_as_uint(mx.x), __float_as_uint(my.x), 0x5410);
        const unsigned kb = __byte_perm(__float_as_uint(mx.y), __float_as_uint(my.y), 0x5410);
        // the plane before this lane's first one: plane 3 of the pass before (pair 0: the partner's result of the LAST
        // round) or plane 1 of this pass (pair 1: the partner's result of THIS round)
        const unsigned prev = __shfl_xor_sync(0xffffffffu, pp0 ? klast : kb, 4);
        klast = kb;
        const bool moved_a = owner && ka != prev, moved_b = owner && kb != ka;
        const unsigned ma = __ballot_sync(0xffffffffu, moved_a), mb = __ballot_sync(0xffffffffu, moved_b);
        unsigned info_a = 0, info_b = 0;
        if (ma | mb) {
            const unsigned below = (1u << lane) - 1u;
            if (moved_a) {
                const int x0 = (int)(short)(ka & 0xffffu), y0 = (int)(short)(ka >> 16);
                const int mine = __popc(ma & below);
                const bool inside = (unsigned)x0 < (unsigned)(p.W - 1) && (unsigned)y0 < (unsigned)(p.H - 1);
                if (mine < kPreSlots && (kPreBulk ? inside : true)) {
                    info_a = (region + mine * 512) | 1u;
                    if (kPreBulk)
                        bulk_footprint(region + mine * 512, p.feats + ((size_t)(jview + 1) * p.HW + (size_t)y0 * p.W + x0) * C,
                                       row_bytes, cbar);
                    else
                        tma_footprint(region + mine * 512, &texmap, x0, y0, jview + 1, cbar);
                } else {
                    info_a = 2u | ((unsigned)(x0 + 8) << 2) | ((unsigned)(y0 + 8) << 17);
                }
            }
            if (moved_b) {
                const int x0 = (int)(short)(kb & 0xffffu), y0 = (int)(short)(kb >> 16);
                const int mine = __popc(ma) + __popc(mb & below);
                const bool inside = (unsigned)x0 < (unsigned)(p.W - 1) && (unsigned)y0 < (unsigned)(p.H - 1);
                if (mine < kPreSlots && (kPreBulk ? inside : true)) {
                    info_b = (region + mine * 512) | 1u;
                    if (kPreBulk)
                        bulk_footprint(region + mine * 512, p.feats + ((size_t)(jview + 1) * p.HW + (size_t)y0 * p.W + x0) * C,
                                       row_bytes, cbar);
                    else
                        tma_footprint(region + mine * 512, &texmap, x0, y0, jview + 1, cbar);
                } else {
                    info_b = 2u | ((unsigned)(x0 + 8) << 2) | ((unsigned)(y0 + 8) << 17);
                }
            }
        }
        if (owner) {
            sts128(gwb, make_float4(fx.x, fy.x, fxy.x, __uint_as_float(info_a)));
            sts128(gwb + kPreGeoPlane, make_float4(fx.y, fy.y, fxy.y, __uint_as_float(info_b)));
        }
    };

    // ---- prologue: the first pass's table into buffer 0, its footprints into region 0
    float2 dnext;
    next_depths(dnext);
    project_round(dnext, gw - kPreGeoBuf, slot_w, wbar);
    next_depths(dnext);
    __syncwarp();
    if (lane == 0) mbar_arrive(wbar);

    const float invV = 1.f / (float)(NV + 1);
    const float2 ninv = splat(-invV), pinv = splat(invV);

    // Staging: batch n is computed into ring slot n % 2 while batch n - 1 is drained, one plane per plane.
    unsigned bar_c = bar0, bar_d = bar0, par_d = 0;
    int slot_c = 0, slot_d = 0;
    float4 dw;
    auto begin_drain = [&]() {
        mbar_wait(bar_d, par_d);
        bar_d += 8;
        if (++slot_d == NBUF) { slot_d = 0; bar_d = bar0; par_d ^= 1; }
    };
    auto drain_one = [&]() {
        const float4 w = lds128(dr);
        dr += kPreTilePlane;
        stg128_na(optr, w);
        optr += p.out_sd;
    };

    float4 g[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) g[v] = lds128(gr + v * PPW * 16);      // table buffer 0 is complete (__syncwarp above)
    int n = 0;
#pragma unroll 1
    for (int b0 = d0; b0 < d1; b0 += KT, ++n) {
        const int par = n & 1;
        mbar_wait(wbar + par * 8, (n >> 1) & 1);     // this pass's footprints have landed (issued a pass ago)
        const bool draining = n >= 1;
        if (draining) begin_drain();
#pragma unroll
        for (int t = 0; t < KT; ++t) {
            if (draining) {                          // the staged chunk this plane writes out: read now, stored at the bottom
                dw = lds128(dr);
                dr += kPreTilePlane;
            }
            unsigned moved = 0;
#pragma unroll
            for (int v = 0; v < NV; ++v) moved |= __float_as_uint(g[v].w);
            if (moved) {
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    consume_ws_shift<128>(tex[v], __float_as_uint(g[v].w), lane_off,
                                          feats_c + (size_t)(v + 1) * p.HW * C, row_bytes, p.W, p.H, rf);
            }
            if (t == 0)                              // the next pass: table into the other buffer, copies into the other region
                project_round(dnext, gw, slot_w + (par ^ 1) * kPreRegion, wbar + (par ^ 1) * 8);

            float2 s[NP], sq[NP];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float2 fx = splat(g[v].x), fy = splat(g[v].y), fxy = splat(g[v].z);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    float2 o = __ffma2_rn(fx, tex[v][1][j], tex[v][0][j]);
                    o = __ffma2_rn(fy, tex[v][2][j], o);
                    o = __ffma2_rn(fxy, tex[v][3][j], o);
                    if (v == 0) {
                        s[j] = o;
                        sq[j] = __fmul2_rn(o, o);
                    } else {
                        s[j] = __fadd2_rn(s[j], o);
                        sq[j] = __ffma2_rn(o, o, sq[j]);
                    }
                }
            }
            if (t == 0) next_depths(dnext);
            if (t + 1 < KT) {
#pragma unroll
                for (int v = 0; v < NV; ++v) g[v] = lds128(gr + (t + 1) * kPreGeoPlane + v * PPW * 16);
            } else {                                 // last plane of the pass: swap the table buffers
                __syncwarp();                        // the other table is complete; this one is free
                if (lane == 0) mbar_arrive(wbar + (par ^ 1) * 8);   // ... and the next pass's copies are all announced
                gr += gflip;
                gw -= gflip;
                gflip = 0u - gflip;
#pragma unroll
                for (int v = 0; v < NV; ++v) g[v] = lds128(gr + v * PPW * 16);
            }
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const float2 tneg = __fmul2_rn(s[j], ninv);            // -sum/V
                const float2 w = __ffma2_rn(tneg, s[j], sq[j]);        // sq - sum^2/V
                const float2 r = __fmul2_rn(w, pinv);                  // sq/V - (sum/V)^2
                sts32(tw + (2 * j) * PIX * 4 + t * kPreTilePlane, r.x);
                sts32(tw + (2 * j + 1) * PIX * 4 + t * kPreTilePlane, r.y);
            }
            if (draining) {
                stg128_na(optr, dw);
                optr += p.out_sd;
            }
        }
        // batch n is staged in ring slot slot_c: announce it (one arrival per warp)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_c);
        bar_c += 8;
        tw += kPreTileBuf;
        if (++slot_c == NBUF) { slot_c = 0; bar_c = bar0; tw -= kPreTileRing; }
        if (draining && slot_d == 0) dr -= kPreTileRing;   // the drained slot was the last of the ring
    }
    // the sweep is over: the last batch has nothing left to hide behind
    if (n >= 1) {
        begin_drain();
        for (int k = min(KT, d1 - (d0 + (n - 1) * KT)); k > 0; --k) drain_one();
    }
}

template <int NV>
int launch_sweep_pre(const SweepParams& p, dim3 grid, cudaStream_t stream) {
    const size_t smem = kPreBarBytes + kPreSlotBytes + kPreTableBytes + kPreTileRing +
                        (p.perpix ? 0 : (size_t)(p.d_chunk + kLeanHypPad) * 4);
    if (smem > 112 * 1024) return -1;                // two CTAs per SM
    if ((reinterpret_cast<uintptr_t>(p.feats) & 15) != 0) return -1;
    CUtensorMap map;
    if (int rc = make_texel_map(&map, p, NV + 1)) return rc;
    void (*kern)(const SweepParams, const CUtensorMap) = p.perpix ? sweep_pre_kernel<NV, true> : sweep_pre_kernel<NV, false>;
    static SmemOptIn opted[2];
    if (int rc = opted[p.perpix ? 1 : 0].ensure(kern, smem)) return rc;
    kern<<<grid, 256, smem, stream>>>(p, map);
    count_launch();
    return check_launch("sweep_pre_kernel");
}

// returns -1 when the shape is not covered (the caller falls back to sweep_quad)
inline int sweep_pre_dispatch(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream) {
    if (p.pooled) return -1;                         // views addressed by stride here: dense [V,H,W,C] texels only
    if (p.W > 16000 || p.H > 16000) return -1;       // 15-bit corner fields in the table entry
    if (p.C != 32 || (p.HW & 31) != 0) return -1;
    if (((p.out_sc | p.out_sd) & 3) != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15) != 0) return -1;
    switch (nv) {
        case 1: return launch_sweep_pre<1>(p, grid, stream);
        case 2: return launch_sweep_pre<2>(p, grid, stream);
        case 3: return launch_sweep_pre<3>(p, grid, stream);
        case 4: return launch_sweep_pre<4>(p, grid, stream);
        default: return -1;
    }
}

}  // namespace d3d

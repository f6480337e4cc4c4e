// Fused plane sweep, warp-specialised form for 32-channel features (round-2 experiment, `variant` 8; NOT the default:
// 6.54 ms at cfg2 against sweep_quad's 5.37 ms -- DESIGN.md 2.7 has the account).
//
// Why it was built: profiles/ncu_r1n.txt -- in sweep_quad_kernel the FP32 arithmetic of the volume is 46 of the ~120 warp
// instructions a warp issues per plane; the rest is the projection chain, the footprint-key tests, the re-fetch (address
// arithmetic + 4 global loads whose latency the warp eats: 19 % of all stall samples), the drain of the staged rows and
// the pointer upkeep around them, and with 122 registers only 4 warps per scheduler hide it.  Here the CTA is split by
// role (setmaxnreg moves the registers to where the footprints live):
//
//   * 8 CONSUMER warps (104 registers) own the register-resident footprint cache (lane = pixel q of the warp x group cg
//     of 4 channels, as in sweep_quad.cuh) and do nothing but: 4 LDS.128 of the plane's table entries, one test "did
//     anything move", the packed FFMA2 arithmetic, 4 STS.32 into the staging tile -- 56 instructions per plane, 42 of them
//     arithmetic.  A moved footprint is picked up from a shared-memory SLOT (4 LDS.128 + the packed rebuild,
//     consume_ws_shift): no address arithmetic, no global load, no long-scoreboard stall.
//   * 4 PRODUCER warps (32 registers), one lane per (pixel, source view) STREAM of 8 pixels: the packed projection chain
//     (two planes per FFMA2/FADD2, operation order of the reference, see project2 in sweep_quad.cuh), the move test
//     against the stream's previous floor corner (lane-local), and for every moved footprint ONE
//     cp.async.bulk.tensor.4d -- the 2 x 2 x 32-channel box of the texel tensor at (x0, y0, view), texels outside the
//     image arriving as zeros (the tensor map's out-of-bounds fill = grid_sample's zeros padding) -- into a slot of the
//     pass's region, a whole pass (4 planes) ahead of its use.  The table entry {fx, fy, fx*fy, info} tells the consumer
//     where the slot is.  When a region is full (14 slots per producer warp and pass; the cfg2 rig averages 8 moves)
//     the entry carries the floor corner instead and the consumer fetches from global memory itself (rare).  The same
//     warps drain the staged rows (LDS.128 -> 128-byte STG.128 rows, L1::no_allocate), after the pass is produced, so
//     the drain overlaps the flight time of the pass's copies.
//
// Synchronisation is all mbarriers (no bar.sync in the loop), two-deep rings:
//     tfull[pw][2]  producer warp pw -> its two consumer warps: table + slots of a pass are complete (one arrive that
//                   releases the table stores + the bytes of the pass's copies, expect_tx / complete_tx)
//     tempty[pw][2] the two consumer warps -> producer: pass consumed, table + slot region may be overwritten
//     sfull[2]      8 consumer warps -> producers: a batch of 4 planes is staged
//     sempty[2]     4 producer warps -> consumers: the batch has left, the staging buffer is free
// The producer runs one pass ahead: produce(n + 1) happens while the consumers compute pass n.
//
// What it measured (profiles/ncu_r2_ws_variant8.txt): the consumers run at ~85 % FP32-pipe utilisation when fed, but wait
// on tfull for 21-32 % of their time: a producer warp needs ~390 instructions per pass and gets one issue slot in six.
//
// Variance volume: the reference texel is subtracted from every footprint's A at rebuild time (variance is shift
// invariant), so the reference view contributes nothing to sum and sum of squares: one packed add per channel pair
// and plane less, and less cancellation than the reference's own sum-of-squares form.
#pragma once
#include <cuda.h>

#include "sweep_quad.cuh"

namespace d3d {

constexpr int kWsSlots = 14;            // footprint slots per producer warp and pass parity
constexpr int kWsConsumerRegs = 104;
constexpr int kWsProducerRegs = 32;
constexpr unsigned kWsBarBytes = 256;
constexpr unsigned kWsTablePlane = 32 * 16;                  // one plane of one producer warp: 32 streams x 16 B
constexpr unsigned kWsTablePass = 4 * kWsTablePlane;         // 2 KB
constexpr unsigned kWsTableBytes = 4 * 2 * kWsTablePass;     // 4 producer warps x 2 parities
constexpr unsigned kWsSlotBytes = 512;
constexpr unsigned kWsRegion = kWsSlots * kWsSlotBytes;
constexpr unsigned kWsSlotsBytes = 4 * 2 * kWsRegion;
constexpr unsigned kWsStagePlane = 32 * 32 * 4;              // [32 channels][32 pixels] fp32
constexpr unsigned kWsStageBatch = 4 * kWsStagePlane;        // 16 KB
constexpr unsigned kWsStageBytes = 2 * kWsStageBatch;
constexpr unsigned kWsRefBytes = 32 * 128;

// One footprint = the 2 x 2 x 32-channel box of the texel tensor at (x0, y0) of view `view`, into a 512-byte slot;
// texels outside the image arrive as zeros (the tensor map's out-of-bounds fill = grid_sample's zeros padding).
__device__ __forceinline__ void tma_footprint(unsigned dst, const CUtensorMap* map, int x0, int y0, int view, unsigned bar) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], 512;" ::"r"(bar) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(map), "r"(0), "r"(x0), "r"(y0), "r"(view), "r"(bar) : "memory");
}
__device__ __forceinline__ void stg128_na(float* ptr, const float4& w) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(w.x), "f"(w.y), "f"(w.z),
                 "f"(w.w) : "memory");
}

// The same footprint as two plain bulk copies (rows y0 and y0 + 1, two adjacent texels = 256 contiguous bytes each): no
// tensor-map lookup, no coordinate arithmetic in the copy engine -- for footprints wholly inside the image only.
__device__ __forceinline__ void bulk_footprint(unsigned dst, const float* nw_texel, unsigned row_bytes, unsigned bar) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], 512;" ::"r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];"
                 ::"r"(dst), "l"(nw_texel), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];"
                 ::"r"(dst + 256), "l"(reinterpret_cast<const char*>(nw_texel) + row_bytes), "r"(bar) : "memory");
}

// MODE: D3D_AGG_VARIANCE (others: sweep_quad.cuh).  kPerPix: per-pixel hypotheses [D,H,W].
template <int NV, int MODE, bool kPerPix>
__global__ void __launch_bounds__(384, 2) sweep_ws_kernel(const SweepParams p, const __grid_constant__ CUtensorMap texmap) {
    static_assert(MODE == D3D_AGG_VARIANCE, "warp-specialised sweep: variance volume");
    constexpr int KT = 4;
    extern __shared__ __align__(1024) float4 smem4[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    const int d0 = p.d_begin + blockIdx.y * p.d_chunk;
    const int d1 = min(d0 + p.d_chunk, p.d_end);
    if (d0 >= d1) return;
    const int npass = (d1 - d0 + KT - 1) / KT;
    const long long grp_base = (long long)blockIdx.x * 32;

    // ---- shared-memory map: mbarriers | tables | slots | staging | hypotheses
    const unsigned bar0 = smem_u32(smem4);
    const unsigned tab0 = bar0 + kWsBarBytes;
    const unsigned slot0 = tab0 + kWsTableBytes;
    const unsigned stg0 = slot0 + kWsSlotsBytes;
    const unsigned ref0 = stg0 + kWsStageBytes;            // the CTA's 32 reference texels (read at rebuild time only)
    const unsigned trn0 = ref0 + kWsRefBytes;               // per source view: translation column of the relative pose
    const unsigned hyp_s = trn0 + 64;
    // barrier addresses: tfull[pw][par] = bar0 + (pw*2+par)*8; tempty = +64; sfull[par] = +128; sempty[par] = +144
    if (threadIdx.x < 8) mbar_init(bar0 + threadIdx.x * 8, 1);        // + the bytes of the pass's footprint copies
    else if (threadIdx.x < 16) mbar_init(bar0 + threadIdx.x * 8, 2);
    else if (threadIdx.x < 18) mbar_init(bar0 + 128 + (threadIdx.x - 16) * 8, 8);
    else if (threadIdx.x < 20) mbar_init(bar0 + 144 + (threadIdx.x - 18) * 8, 4);
    if (threadIdx.x >= 32 && threadIdx.x < 32 + NV)
        sts128(trn0 + (threadIdx.x - 32) * 16, make_float4(p.pose[(threadIdx.x - 32) * 16 + 3], p.pose[(threadIdx.x - 32) * 16 + 7],
                                                           p.pose[(threadIdx.x - 32) * 16 + 11], 0.f));
    if (!kPerPix) {
        const int n = d1 - d0 + kLeanHypPad;
        for (int i = threadIdx.x; i < n; i += 384) sts32(hyp_s + i * 4, __ldg(p.hyps + min(d0 + i, d1 - 1)));
    }
    __syncthreads();

    if (warp >= 8) {
        // =============================================================== PRODUCER: one lane per (pixel, view) stream
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kWsProducerRegs));
        const int pw = warp - 8;
        const int v = (lane >> 2) & 3;                     // stream = (consumer warp of the pair, view, pixel)
        const bool active = v < NV;
        const int jv = min(v, NV - 1);
        float rx, ry, rz;
        {
            const long long pix_raw = grp_base + (2 * pw + (lane >> 4)) * 4 + (lane & 3);
            const int pix = pix_raw < p.HW ? (int)pix_raw : p.HW - 1;
            const int py = pix / p.W, px = pix - py * p.W;
            const float* m = p.pose + jv * 16;
            rx = fmaf(m[2], 1.f, fmaf(m[1], (float)py, m[0] * (float)px));
            ry = fmaf(m[6], 1.f, fmaf(m[5], (float)py, m[4] * (float)px));
            rz = fmaf(m[10], 1.f, fmaf(m[9], (float)py, m[8] * (float)px));
            if (p.rays) {   // the reference's own rot @ [x,y,1] (cuBLAS), whatever order it rounded in
                const float* rr = p.rays + (size_t)jv * 3 * p.HW + pix;
                rx = __ldg(rr); ry = __ldg(rr + p.HW); rz = __ldg(rr + 2 * (size_t)p.HW);
            }
        }
        const unsigned tvec = trn0 + jv * 16;              // this view's translation (staged: 3 registers less)
        unsigned kprev = 0x7fff7fffu;                      // no footprint has this corner: the first plane always moves
        unsigned hs = hyp_s;
        const float* hp = kPerPix ? p.hyps + (size_t)min((long long)p.HW - 1, grp_base + (2 * pw + (lane >> 4)) * 4 + (lane & 3)) +
                                        (size_t)d0 * p.HW : p.hyps;
        int hplane = d0;
        // drain: the 128 producer threads move a staged batch (4 planes x 32 rows x 8 chunks of 16 bytes), 8 chunks each
        float* optr;
        {
            const int tid_p = pw * 32 + lane;
            const int c0 = tid_p >> 3, c4 = (tid_p & 7) * 4;
            optr = p.out + ((long long)c0 * p.out_sc + (long long)(d0 - p.d_begin) * p.out_sd + grp_base + c4);
        }

        auto drain = [&](int m) {
            const int par = m & 1;
            mbar_wait(bar0 + 128 + par * 8, (m >> 1) & 1);
            const int planes = min(KT, d1 - (d0 + m * KT));
            const int tid_p = pw * 32 + lane;
            const int c0 = tid_p >> 3, c4 = (tid_p & 7) * 4;
            // rows c0 and c0 + 16 of a staged plane: the second one's swizzle differs by 16 pixels (= 64 bytes)
            unsigned sa = stg0 + par * kWsStageBatch + (c0 * 32 + (c4 ^ ((4 * (c0 >> 2)) & 31))) * 4;
            const long long sc16 = 16 * p.out_sc;
            if (planes == KT) {                            // whole batch: four loads in flight, then their four stores
#pragma unroll
                for (int t = 0; t < KT; t += 2) {
                    const float4 a0 = lds128(sa), b0 = lds128((sa ^ 64u) + 2048u);
                    const float4 a1 = lds128(sa + kWsStagePlane), b1 = lds128(((sa + kWsStagePlane) ^ 64u) + 2048u);
                    stg128_na(optr, a0);
                    stg128_na(optr + sc16, b0);
                    stg128_na(optr + p.out_sd, a1);
                    stg128_na(optr + p.out_sd + sc16, b1);
                    sa += 2 * kWsStagePlane;
                    optr += 2 * p.out_sd;
                }
            } else {
                for (int t = 0; t < KT; ++t) {
                    if (t < planes) {
                        const float4 a = lds128(sa), b = lds128((sa ^ 64u) + 2048u);
                        stg128_na(optr, a);
                        stg128_na(optr + sc16, b);
                    }
                    sa += kWsStagePlane;
                    optr += p.out_sd;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 144 + par * 8);
        };

        for (int n = 0; n < npass; ++n) {
            const int par = n & 1;
            if (n >= 2) mbar_wait(bar0 + 64 + pw * 16 + par * 8, ((n >> 1) & 1) ^ 1);
            const unsigned region = slot0 + (pw * 2 + par) * kWsRegion;
            unsigned tw = tab0 + (pw * 2 + par) * kWsTablePass + lane * 16;
            unsigned dst = region;                         // next free slot
            const unsigned tfull = bar0 + pw * 16 + par * 8;
            int room = kWsSlots;
#pragma unroll 1
            for (int pair = 0; pair < 2; ++pair) {
                float2 d;
                if (kPerPix) {
                    const size_t hw = (size_t)p.HW;
                    const int pa = min(hplane, d1 - 1), pb = min(hplane + 1, d1 - 1);
                    d.x = __ldg(hp + (long long)(pa - d0) * (long long)hw);
                    d.y = __ldg(hp + (long long)(pb - d0) * (long long)hw);
                    hplane += 2;
                } else {
                    d = lds64(hs);
                    hs += 8;
                }
                const float4 tr = lds128(tvec);
                // packed projection, the reference's IEEE operation on each half (project2 in sweep_quad.cuh)
                const float2 Xm = __fmul2_rn(splat(rx), d), Ym = __fmul2_rn(splat(ry), d), Zm = __fmul2_rn(splat(rz), d);
                const float2 X = f2(__fadd_rn(Xm.x, tr.x), __fadd_rn(Xm.y, tr.x));
                const float2 Y = f2(__fadd_rn(Ym.x, tr.y), __fadd_rn(Ym.y, tr.y));
                const float2 Z = f2(__fadd_rn(Zm.x, tr.z), __fadd_rn(Zm.y, tr.z));
                float2 r;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(Z.x));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(Z.y));
                const float2 nZ = neg2(Z);
                r = __ffma2_rn(__ffma2_rn(nZ, r, splat(1.f)), r, r);
                const float2 qu = __fmul2_rn(X, r), qv = __fmul2_rn(Y, r);
                const float2 u = __ffma2_rn(__ffma2_rn(nZ, qu, X), r, qu);
                const float2 w = __ffma2_rn(__ffma2_rn(nZ, qv, Y), r, qv);
                float2 ix = __fmul2_rn(u, splat(p.inv_half_w));
                float2 iy = __fmul2_rn(w, splat(p.inv_half_h));
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

#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const unsigned key = half ? kb : ka;
                    const bool moved = active && key != kprev;
                    kprev = key;
                    unsigned info = 0;
                    const unsigned mask = __ballot_sync(0xffffffffu, moved);
                    if (mask) {
                        if (moved) {
                            const int x0 = (int)(short)(key & 0xffffu), y0 = (int)(short)(key >> 16);
                            const int mine = __popc(mask & ((1u << lane) - 1u));
                            if (mine < room) {
                                info = (dst + mine * kWsSlotBytes) | 1u;
                                tma_footprint(dst + mine * kWsSlotBytes, &texmap, x0, y0, jv + 1, tfull);
                            } else {
                                info = 2u | ((unsigned)(x0 + 8) << 2) | ((unsigned)(y0 + 8) << 17);
                            }
                        }
                        const int nm = __popc(mask);
                        dst += nm * kWsSlotBytes;
                        room -= nm;
                    }
                    sts128(tw, make_float4(half ? fx.y : fx.x, half ? fy.y : fy.x, half ? fxy.y : fxy.x,
                                           __uint_as_float(info)));
                    tw += kWsTablePlane;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(tfull);             // releases the table; the phase ends when the copies have landed too
            // the batch that left the consumers when this pass's table slot became free: draining it AFTER the pass is
            // produced overlaps the drain with the flight time of the pass's footprint copies
            if (n >= 2) drain(n - 2);
        }
        if (npass >= 2) drain(npass - 2);
        drain(npass - 1);
        return;
    }

    // =================================================================== CONSUMER: lane = (pixel q, channel group cg)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWsConsumerRegs));
    constexpr int NP = 2;
    const int cg = lane & 7, q = lane >> 3;
    const int choff = cg * 4;
    const int pw = warp >> 1;
    const unsigned ref_a = ref0 + (warp * 4 + q) * 128 + choff * 4;
    {
        const long long pix_raw = grp_base + warp * 4 + q;
        const int pix = pix_raw < p.HW ? (int)pix_raw : p.HW - 1;
        sts128(ref_a, ldg4(p.feats + (size_t)pix * 32 + choff));     // read back by this lane only
    }
    float2 tex[NV][4][NP];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < NP; ++j) tex[v][k][j] = f2(0.f, 0.f);
    const unsigned lane_off = (unsigned)choff * 4u - 1u;   // slot address = (info - 1) + the lane's channel offset
    const unsigned row_bytes = (unsigned)p.W * 128u;
    const unsigned tfull = bar0 + pw * 16, tempty = bar0 + 64 + pw * 16;
    const unsigned tabr = tab0 + pw * 2 * kWsTablePass + ((warp & 1) * 16 + q) * 16;    // + par, + t*plane, + v*64
    // staging: row = channel, the pixel column swizzled by the lane that owns the row (as in sweep_quad.cuh)
    const unsigned stw = stg0 + (choff * 32 + ((warp * 4 + q) ^ ((4 * cg) & 31))) * 4;
    const float invV = 1.f / (float)(NV + 1);
    const float2 ninv = splat(-invV), pinv = splat(invV);

    float4 g[NV];
    for (int n = 0; n < npass; ++n) {
        const int par = n & 1;
        mbar_wait(tfull + par * 8, (n >> 1) & 1);
        unsigned tr = tabr + par * kWsTablePass;
#pragma unroll
        for (int v = 0; v < NV; ++v) g[v] = lds128(tr + v * 64);
        if (n >= 2) mbar_wait(bar0 + 144 + par * 8, ((n >> 1) & 1) ^ 1);
        const unsigned tw = stw + par * kWsStageBatch;
#pragma unroll
        for (int t = 0; t < KT; ++t) {
            unsigned moved = 0;
#pragma unroll
            for (int v = 0; v < NV; ++v) moved |= __float_as_uint(g[v].w);
            if (moved) {
                const float4 rw = lds128(ref_a);
                const float2 rf[NP] = {f2(rw.x, rw.y), f2(rw.z, rw.w)};
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    consume_ws_shift<128>(tex[v], __float_as_uint(g[v].w), lane_off,
                                          p.feats + (size_t)(v + 1) * p.HW * 32 + choff, row_bytes, p.W, p.H, rf);
            }
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
            if (t + 1 < KT) {
#pragma unroll
                for (int v = 0; v < NV; ++v) g[v] = lds128(tr + (t + 1) * kWsTablePlane + v * 64);
            }
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const float2 tneg = __fmul2_rn(s[j], ninv);            // -sum/V
                const float2 w = __ffma2_rn(tneg, s[j], sq[j]);        // sq - sum^2/V
                const float2 r = __fmul2_rn(w, pinv);                  // sq/V - (sum/V)^2
                sts32(tw + (2 * j) * 128 + t * kWsStagePlane, r.x);
                sts32(tw + (2 * j + 1) * 128 + t * kWsStagePlane, r.y);
            }
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(bar0 + 128 + par * 8);
            mbar_arrive(tempty + par * 8);
        }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point: the library does not link libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// texels [V][H][W][32] fp32 as a 4-D tensor (innermost first: channel, x, y, view); box = one 2 x 2 footprint
inline int make_texel_map(CUtensorMap* map, const SweepParams& p, int views) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(D3D_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[4] = {32, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)views};
    const cuuint64_t strides[3] = {128, (cuuint64_t)p.W * 128, (cuuint64_t)p.HW * 128};
    const cuuint32_t box[4] = {32, 2, 2, 1}, estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.feats), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(D3D_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return D3D_OK;
}

template <int NV, int MODE>
int launch_sweep_ws(const SweepParams& p, dim3 grid, cudaStream_t stream) {
    const size_t smem = kWsBarBytes + kWsTableBytes + kWsSlotsBytes + kWsStageBytes + kWsRefBytes + 64 +
                        (p.perpix ? 0 : (size_t)(p.d_chunk + kLeanHypPad) * 4);
    if (smem > 112 * 1024) return -1;                // two CTAs per SM; absurd depth chunks go elsewhere
    if ((reinterpret_cast<uintptr_t>(p.feats) & 15) != 0) return -1;
    CUtensorMap map;
    if (int rc = make_texel_map(&map, p, NV + 1)) return rc;
    void (*kern)(const SweepParams, const CUtensorMap) = p.perpix ? sweep_ws_kernel<NV, MODE, true> : sweep_ws_kernel<NV, MODE, false>;
    static SmemOptIn opted[2];
    if (int rc = opted[p.perpix ? 1 : 0].ensure(kern, smem)) return rc;
    kern<<<grid, 384, smem, stream>>>(p, map);
    count_launch();
    return check_launch("sweep_ws_kernel");
}

// returns -1 when the shape is not covered (the caller falls back to sweep_quad / sweep_lean / sweep_base)
template <int MODE>
int sweep_ws_dispatch(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream) {
    if (p.pooled) return -1;                         // views addressed by stride here: dense [V,H,W,C] texels only
    if (p.W > 16000 || p.H > 16000) return -1;       // 15-bit corner fields in the table entry
    if (p.C != 32 || (p.HW & 31) != 0) return -1;
    if ((unsigned long long)(nv + 1) * (unsigned long long)p.HW >= (1ull << 25)) return -1;   // 32-bit byte offsets into `feats`
    if (((p.out_sc | p.out_sd) & 3) != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15) != 0) return -1;
    switch (nv) {
        case 1: return launch_sweep_ws<1, MODE>(p, grid, stream);
        case 2: return launch_sweep_ws<2, MODE>(p, grid, stream);
        case 3: return launch_sweep_ws<3, MODE>(p, grid, stream);
        case 4: return launch_sweep_ws<4, MODE>(p, grid, stream);
        default: return -1;
    }
}

}  // namespace d3d

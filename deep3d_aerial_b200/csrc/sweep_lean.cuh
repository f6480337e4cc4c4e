// Fused plane sweep, two planes per projection pass (variant 6) -- the fall-back of sweep_quad.cuh for the shapes that
// kernel is not instantiated for (C = 4, 8, 64; H*W not a multiple of the pixel tile).
//
// Decomposition (C = CPT*LPP channels; CPT = 4 or 8 channels per lane, LPP in {1,2,4,8} lanes per pixel):
//   lane   = (pixel q of the warp, channel group cg of CPT channels);  LPP lanes share a pixel
//   warp   = 32/LPP consecutive reference pixels x all channels
//   group  = LPP warps = 32 consecutive pixels (one 128-byte row segment per output channel)
//   CTA    = 8 warps = 8/LPP groups;  blockIdx.y = chunk of depth planes
// Every thread walks its depth planes in order and keeps, per source view, the 2x2 texel footprint of its CPT channels
// in registers; a footprint is re-fetched only when floor(ix), floor(iy) move.  The projection is computed ONCE per
// (pixel, view, plane): the LPP lanes of a pixel split the source views (and, when there are more lanes than views, two
// consecutive planes) and publish {fx, fy, fx*fy, key} through a per-warp double-buffered shared-memory table, one pass
// ahead; bilinear interpolation as A + fx*B + fy*C + fx*fy*D in packed fp32x2; results leave through an XOR-swizzled
// shared-memory tile as full 128-byte rows.  (The first kernel of this family, sweep_fast.cuh, is retired.)
//
// The per-plane loop is stated
// so that nothing is re-derived per plane or per batch: the kernel is issue- and FP32-pipe-bound, and under the
// 128-register cap (16 resident warps per SM) the compiler otherwise rematerialises addresses from threadIdx
// (profiles/ncu_r1k.txt: 45 warp instructions per voxel, 74 per staged batch spent on pointer arithmetic).
//
//   * footprint key = BYTE OFFSET.  The lane that projects a (pixel, view, plane) also turns the footprint into
//     the byte offset of its north-west texel (view offset included) and publishes that as the key; a consumer
//     whose key moved adds it to its base pointer and loads -- no decode, clamp or mask arithmetic in the
//     re-fetch (59 -> 22 instructions).  Footprints that touch the border or lie outside the image carry their
//     corner instead (bit 0 set; key 1 = nothing inside) and take a predicated path.
//   * floor() of the sample position is one FADD.RM against 1.5*2^23 (round-down add), exact for |x| < 2^22.
//   * fronto-parallel hypotheses are staged in shared memory once per CTA; per-pixel hypotheses are walked with
//     a running pointer (template flag kPerPix).
//   * staging tiles, the drain pointers and the mbarrier slot/parity are running state: a staged batch costs a
//     wait, an arrive and a wrap test.
//
// Per batch of kLeanTilePlanes planes, per lane:
//     for each pass of PB planes:    __syncwarp
//         LDS.128 x NV               projection table of this plane           (geo, per warp, 2 buffers)
//         1 compare/branch           any footprint key moved?  -> in-place PTX re-fetch (rare)
//         projection of the NEXT pass for the (view, plane) this lane owns, interleaved with
//         3 FFMA2 x NV x CPT/2       A + fx*B + fy*C + fxy*D
//         FADD2 + FFMA2              sum, sum of squares;  3 packed ops of variance epilogue
//         STS.32 x CPT               swizzled staging tile                    (per group, 4 buffers)
//         LDS.128 + STG.128 x CPT/4  one plane of an EARLIER batch leaves as 128-byte rows
//     mbarrier arrive (split phase: the wait happens two batches later, when that batch is drained)
#pragma once
#include "sweep_util.cuh"

namespace d3d {

__device__ __forceinline__ unsigned smem_u32(const void* ptr) {
    return (unsigned)__cvta_generic_to_shared(ptr);
}
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(unsigned addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

constexpr int kLeanTilePlanes = 4;
constexpr int kLeanTileBuffers = 4;
constexpr int kLeanHypPad = 16;      // staged hypotheses are padded so the look-ahead never needs a clamp

// ---- split-phase group synchronisation (mbarrier): a warp announces that its share of a staged batch is
// written and only waits, two batches later, when it starts draining that batch -- so one warp stalled on a
// re-fetch no longer stops the other warps of its group (a bar.sync rendezvous did: 12-16 % of all cycles)
__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned addr) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT;\n\t}" ::"r"(addr), "r"(parity) : "memory");
}

// Projection of one reference pixel onto one source view at one depth (operation order: see project_frac in
// sweep_util.cuh), returning {fx, fy, fx*fy, key} with the offset key described above.
//   view_off = byte offset of the source view inside `feats`;  texel_bytes = C*4
template <bool kIeeeDiv>
__device__ __forceinline__ float4 project_off(float rx, float ry, float rz, float tx, float ty, float tz, float d,
                                              const SweepParams& p, unsigned view_off, int texel_bytes) {
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
    // floor: a round-down add against 1.5*2^23 leaves floor(ix) in the low mantissa bits, exactly
    const float mx = __fadd_rd(ix, kMagic), my = __fadd_rd(iy, kMagic);
    const float fx = __fsub_rn(ix, __fsub_rn(mx, kMagic)), fy = __fsub_rn(iy, __fsub_rn(my, kMagic));
    const int xi = __float_as_int(mx) - kMagicBits, yi = __float_as_int(my) - kMagicBits;
    const bool interior = (unsigned)xi < (unsigned)(p.W - 1) && (unsigned)yi < (unsigned)(p.H - 1);
    const bool any = (unsigned)(xi + 1) <= (unsigned)p.W && (unsigned)(yi + 1) <= (unsigned)p.H;
    const unsigned off = view_off + (unsigned)(yi * p.W + xi) * (unsigned)texel_bytes;
    const unsigned corner = any ? (1u | ((unsigned)(xi + 4) << 4) | ((unsigned)(yi + 4) << 18)) : 1u;
    return make_float4(fx, fy, fx * fy, __uint_as_float(interior ? off : corner));
}

// the views' re-fetch blocks, unrolled at compile time (the view index is an immediate of the PTX block)
template <int NV, int TEXB, int V = 0>
struct RefetchAll {
    template <int NP>
    static __device__ __forceinline__ void run(float2 (&tex)[NV][4][NP], unsigned (&ckey)[NV], const float4 (&g)[NV],
                                               const float* base, unsigned row_bytes, unsigned view_bytes, int W,
                                               int H) {
        refetch_off<V + 1, TEXB>(tex[V], ckey[V], __float_as_uint(g[V].w), base, row_bytes, view_bytes, W, H);
        RefetchAll<NV, TEXB, V + 1>::run(tex, ckey, g, base, row_bytes, view_bytes, W, H);
    }
};
template <int NV, int TEXB>
struct RefetchAll<NV, TEXB, NV> {
    template <int NP>
    static __device__ __forceinline__ void run(float2 (&)[NV][4][NP], unsigned (&)[NV], const float4 (&)[NV],
                                               const float*, unsigned, unsigned, int, int) {}
};

template <int CPT, int NV, int LPP, int MODE, bool kIeeeDiv, bool kPerPix>
__global__ void __launch_bounds__(256, (CPT == 4 ? 2 : 1)) sweep_lean_kernel(const SweepParams p) {
    constexpr int PPW = 32 / LPP;                          // pixels per warp
    constexpr int PB = (LPP >= 2 * NV) ? 2 : 1;            // planes published per geometry pass
    constexpr int KV = (LPP >= NV) ? 1 : (NV + LPP - 1) / LPP;   // projections a lane owns per pass
    constexpr int NP = CPT / 2;                            // channel pairs per lane
    constexpr int C = CPT * LPP;
    constexpr int GROUPS = 8 / LPP;                        // 32-pixel groups per CTA
    constexpr int NRO = CPT / 4;                           // float4 each lane moves per plane at read-out
    constexpr int KT = kLeanTilePlanes;                    // planes per staged batch
    constexpr int NBUF = kLeanTileBuffers;                 // tile buffers in flight
    constexpr unsigned GEO_PLANE = NV * PPW * 16;          // bytes: one plane's table of one warp
    constexpr unsigned GEO_BUF = PB * GEO_PLANE;
    constexpr unsigned TILE_PLANE = C * 32 * 4;            // bytes: one staged plane of one group
    constexpr unsigned TILE_BUF = KT * TILE_PLANE;
    constexpr unsigned TILE_RING = NBUF * TILE_BUF;
    static_assert(KT % PB == 0, "batch must hold whole passes");
    extern __shared__ float4 smem4[];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int cg = lane % LPP;
    const int q = lane / LPP;
    const int grp = warp / LPP, wq = warp % LPP;           // group of the CTA, warp within the group
    const int choff = cg * CPT;
    const long long grp_base = ((long long)blockIdx.x * GROUPS + grp) * 32;

    const int d0 = p.d_begin + blockIdx.y * p.d_chunk;
    const int d1 = min(d0 + p.d_chunk, p.d_end);
    if (d0 >= d1) return;

    // ---- shared-memory map (bytes, shared window): mbarriers | projection tables | staging tiles | hypotheses
    const unsigned bar0 = smem_u32(smem4) + grp * NBUF * 8;        // this group's NBUF mbarriers
    const unsigned sm0 = smem_u32(smem4) + 256;
    const unsigned geo_w = sm0 + warp * 2 * GEO_BUF;
    const unsigned tile_g = sm0 + 8 * 2 * GEO_BUF + grp * TILE_RING;
    const unsigned hyp_s = sm0 + 8 * 2 * GEO_BUF + GROUPS * TILE_RING;

    if (threadIdx.x < GROUPS * NBUF) mbar_init(smem_u32(smem4) + threadIdx.x * 8, LPP);   // one arrival per warp
    if (!kPerPix) {                                        // fronto-parallel sweep: stage the chunk's depths
        const int n = d1 - d0 + kLeanHypPad;
        for (int i = threadIdx.x; i < n; i += 256) sts32(hyp_s + i * 4, __ldg(p.hyps + min(d0 + i, d1 - 1)));
    }
    __syncthreads();

    // ---- projection ownership (see the header) and per-lane constants
    const int po = (LPP >= NV) ? cg / NV : 0;
    const bool owner = (LPP >= NV) ? (cg < PB * NV) : true;
    float rx[KV], ry[KV], rz[KV], tx[KV], ty[KV], tz[KV];
    unsigned voff[KV];
    const unsigned view_bytes = (unsigned)p.HW * (unsigned)(C * 4);
    float2 rf[NP];
    const float* hp = p.hyps;                              // kPerPix: this pixel's hypothesis of the next pass
    {
        const long long pix_raw = grp_base + wq * PPW + q;
        const int pix = pix_raw < p.HW ? (int)pix_raw : p.HW - 1;    // clamp: the warp stays whole
        const int py = pix / p.W, px = pix - py * p.W;
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
            voff[k] = (unsigned)(v + 1) * view_bytes;
        }
        const float* t = p.feats + (size_t)pix * C + choff;
#pragma unroll
        for (int k = 0; k < CPT; k += 4) {
            float4 w = ldg4(t + k);
            rf[k / 2] = f2(w.x, w.y);
            rf[k / 2 + 1] = f2(w.z, w.w);
        }
        if (kPerPix) hp = p.hyps + (size_t)pix + (size_t)(d0 + po) * p.HW;
    }

    float2 tex[NV][4][NP];      // per view: A, B, C, D of the current 2x2 footprint
    unsigned ckey[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        ckey[v] = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < NP; ++j) tex[v][k][j] = f2(0.f, 0.f);
    }
    const float* feats_c = p.feats + choff;                // keys are byte offsets from here
    const unsigned row_bytes = (unsigned)p.W * (unsigned)(C * 4);

    // ---- running shared-memory addresses
    unsigned gr = geo_w + q * 16;                                          // read: + t*GEO_PLANE + v*PPW*16
    unsigned gw = geo_w + GEO_BUF + po * GEO_PLANE + q * 16;               // write (other buffer): + v*PPW*16
    unsigned gflip = GEO_BUF;                                              // +/- distance between the two buffers
    unsigned tw = tile_g + (choff * 32 + ((wq * PPW + q) ^ ((PPW * cg) & 31))) * 4;   // + k*128 per channel row
    const int t_in_grp = wq * 32 + lane;
    unsigned dr[NRO];                                      // drain: staged row chunk this lane moves
    float* optr[NRO];                                      // drain: where it goes
#pragma unroll
    for (int i = 0; i < NRO; ++i) {
        const int idx = t_in_grp + i * LPP * 32;
        const int row = idx >> 3, c4 = (idx & 7) * 4;
        dr[i] = tile_g + (row * 32 + (c4 ^ ((PPW * (row / CPT)) & 31))) * 4;
        optr[i] = p.out + ((long long)row * p.out_sc + (long long)(d0 - p.d_begin) * p.out_sd + grp_base + c4);
    }
    const bool vec_ok = ((p.out_sc | p.out_sd) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
    const bool fast_rows = vec_ok && grp_base + 32 <= p.HW;      // whole group in range, 16-byte stores

    // ---- hypotheses: depth of the plane this lane projects in the next pass
    unsigned hs = hyp_s + po * 4;                          // !kPerPix: running shared-memory address
    int hplane = d0 + po;                                  // kPerPix: plane `hp` points at
    auto next_depth = [&]() -> float {
        float d;
        if (kPerPix) {
            d = __ldg(hplane < d1 ? hp : hp - (size_t)(hplane - (d1 - 1)) * p.HW);
            hp += (size_t)PB * p.HW;
            hplane += PB;
        } else {
            d = lds32(hs);
            hs += PB * 4;
        }
        return d;
    };
    auto project_mine = [&](float4 (&gn)[KV], float depth) {
#pragma unroll
        for (int k = 0; k < KV; ++k)
            gn[k] = project_off<kIeeeDiv>(rx[k], ry[k], rz[k], tx[k], ty[k], tz[k], depth, p, voff[k], C * 4);
    };
    auto store_mine = [&](const float4 (&gn)[KV], unsigned base) {
        if (owner) {
#pragma unroll
            for (int k = 0; k < KV; ++k) {
                const int v = (LPP >= NV) ? cg % NV : cg + k * LPP;
                if (LPP >= NV || NV % LPP == 0 || v < NV) sts128(base + v * PPW * 16, gn[k]);
            }
        }
    };

    // ---- prologue: projections of the first pass into table buffer 0
    float dnext = next_depth();
    {
        float4 gn[KV];
        project_mine(gn, dnext);
        store_mine(gn, gw - GEO_BUF);
    }
    dnext = next_depth();

    const float invV = 1.f / (float)(NV + 1);
    const float2 ninv = splat(-invV), pinv = splat(invV);

    auto drain_one = [&]() {                         // one staged plane -> global, 128-byte rows
#pragma unroll
        for (int i = 0; i < NRO; ++i) {
            const float4 w = lds128(dr[i]);
            float* o = optr[i];
            if (fast_rows) {
                // the volume is write-once: keep it out of L1, which holds the texels the re-fetches hit
                asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(w.x), "f"(w.y),
                             "f"(w.z), "f"(w.w) : "memory");
            } else {
                const long long gp = grp_base + ((t_in_grp + i * LPP * 32) & 7) * 4;
                if (gp < p.HW) o[0] = w.x;
                if (gp + 1 < p.HW) o[1] = w.y;
                if (gp + 2 < p.HW) o[2] = w.z;
                if (gp + 3 < p.HW) o[3] = w.w;
            }
            dr[i] += TILE_PLANE;
            optr[i] += p.out_sd;
        }
    };

    // Staging protocol: batch n is computed into ring slot n % NBUF and drained, one plane per iteration, while
    // batch n+2 is computed.  Starting the drain of batch n-2 waits for it to be complete in every warp of the
    // group, which also implies that they are done draining batch n-4, the previous user of slot n % NBUF.
    // The drain pointers simply keep running (a full batch advances them by exactly one slot / KT planes).
    unsigned bar_c = bar0, bar_d = bar0;             // mbarrier of the slot being computed / drained
    unsigned par_d = 0;                              // phase parity of the drained slot
    int slot_c = 0, slot_d = 0;
    int drain_n = 0;                                 // planes of the draining batch still to write out
    auto begin_drain = [&]() {
        mbar_wait(bar_d, par_d);
        bar_d += 8;
        if (++slot_d == NBUF) { slot_d = 0; bar_d = bar0; par_d ^= 1; }
    };
    auto end_drain = [&]() {                         // the drained slot was the last of the ring: wrap
        if (slot_d == 0) {
#pragma unroll
            for (int i = 0; i < NRO; ++i) dr[i] -= TILE_RING;
        }
    };

    // The table entries of a plane are loaded right after the arithmetic of the plane before it (their registers
    // are dead by then), so the move test at the top of a plane never waits on shared memory.
    float4 g[NV];
    auto load_table = [&](unsigned base) {
#pragma unroll
        for (int v = 0; v < NV; ++v) g[v] = lds128(base + v * PPW * 16);
    };
    __syncwarp();                                    // table buffer 0 is complete
    load_table(gr);

    int n = 0;
    for (int b0 = d0; b0 < d1; b0 += KT) {
        if (n >= 2) {
            begin_drain();
            drain_n = KT;                            // every batch but the last is full
        }
#pragma unroll 1
        for (int t = 0; t < KT; t += PB) {
#pragma unroll
            for (int tt = 0; tt < PB; ++tt) {
                unsigned moved = 0;
#pragma unroll
                for (int v = 0; v < NV; ++v) moved |= __float_as_uint(g[v].w) ^ ckey[v];
                if (moved)                           // some footprint moved: re-fetch those (in place)
                    RefetchAll<NV, C * 4>::run(tex, ckey, g, feats_c, row_bytes, view_bytes, p.W, p.H);
                float4 gn[KV];
                if (tt == 0) project_mine(gn, dnext);          // next pass, interleaved with the arithmetic

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
                            s[j] = __fadd2_rn(rf[j], o);
                            sq[j] = __ffma2_rn(o, o, __fmul2_rn(rf[j], rf[j]));
                        } else {
                            s[j] = __fadd2_rn(s[j], o);
                            sq[j] = __ffma2_rn(o, o, sq[j]);
                        }
                    }
                }
                if (tt == 0) {
                    store_mine(gn, gw);
                    dnext = next_depth();
                }
                if (tt + 1 < PB) {
                    load_table(gr + (tt + 1) * GEO_PLANE);
                } else {                             // last plane of the pass: swap the table buffers
                    __syncwarp();                    // the other table is complete; this one is free
                    gr += gflip;
                    gw -= gflip;
                    gflip = 0u - gflip;
                    load_table(gr);
                }
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const float2 tneg = __fmul2_rn(s[j], ninv);            // -sum/V
                    const float2 w = __ffma2_rn(tneg, s[j], sq[j]);        // sq - sum^2/V
                    const float2 r = __fmul2_rn(w, pinv);                  // sq/V - (sum/V)^2
                    sts32(tw + (2 * j) * 128, r.x);
                    sts32(tw + (2 * j + 1) * 128, r.y);
                }
                tw += TILE_PLANE;
                if (drain_n > 0) { drain_one(); --drain_n; }
            }
        }
        // batch n is staged in ring slot slot_c: announce it (one arrival per warp)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_c);
        ++n;
        bar_c += 8;
        if (++slot_c == NBUF) { slot_c = 0; bar_c = bar0; tw -= TILE_RING; }
        if (drain_n == 0 && n > 2) end_drain();
    }
    // the sweep is over: whatever is still staged has nothing left to hide behind
    if (drain_n > 0) {
        for (; drain_n > 0; --drain_n) drain_one();
        end_drain();
    }
    for (int m = max(n - 2, 0); m < n; ++m) {
        begin_drain();
        for (drain_n = min(KT, d1 - (d0 + m * KT)); drain_n > 0; --drain_n) drain_one();
        end_drain();
    }
}

template <int CPT, int NV, int LPP>
constexpr size_t sweep_lean_smem_fixed() {
    constexpr int PB = (LPP >= 2 * NV) ? 2 : 1;
    return 256 + (size_t)8 * 2 * PB * NV * (32 / LPP) * 16 +
           (size_t)(8 / LPP) * kLeanTileBuffers * kLeanTilePlanes * CPT * LPP * 32 * 4;
}

template <int CPT, int NV, int LPP, int MODE>
int launch_sweep_lean(const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee_div) {
    if (p.pooled) return -1;                         // views addressed by stride here: dense [V,H,W,C] texels only
    // + the chunk's hypotheses (fronto-parallel sweeps)
    const size_t smem = sweep_lean_smem_fixed<CPT, NV, LPP>() + (p.perpix ? 0 : (size_t)(p.d_chunk + kLeanHypPad) * 4);
    if (smem > 200 * 1024) return -1;                // absurd depth chunk: let another kernel take it
    void (*kern)(const SweepParams);
    const int which = (ieee_div ? 2 : 0) + (p.perpix ? 1 : 0);
    switch (which) {
        case 0: kern = sweep_lean_kernel<CPT, NV, LPP, MODE, false, false>; break;
        case 1: kern = sweep_lean_kernel<CPT, NV, LPP, MODE, false, true>; break;
        case 2: kern = sweep_lean_kernel<CPT, NV, LPP, MODE, true, false>; break;
        default: kern = sweep_lean_kernel<CPT, NV, LPP, MODE, true, true>; break;
    }
    static SmemOptIn opted[4];                       // per instantiation and flavour
    if (int rc = opted[which].ensure(kern, smem)) return rc;
    kern<<<grid, 256, smem, stream>>>(p);
    count_launch();
    return check_launch("sweep_lean_kernel");
}

}  // namespace d3d

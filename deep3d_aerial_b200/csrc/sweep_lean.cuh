// Fused plane sweep, production kernel ("variant 0") -- instruction-lean formulation.
//
// Same algorithm and decomposition as sweep_fast.cuh (read its header first); this file re-states the
// per-plane loop with explicit 32-bit shared-memory addresses, running pointers and a fixed batch
// structure, because under the 128-register cap that buys 16 resident warps per SM the compiler
// re-derived every address from threadIdx each plane (56 instructions per voxel, 68 % issue-bound,
// profiles/ncu_r1j.txt).  Per batch of kTilePlanes planes, per lane:
//
//     for each pass of PB planes:    __syncwarp
//         LDS.128 x NV               projection table of this plane           (geo, per warp, 2 buffers)
//         1 compare/branch           any footprint key moved?  -> in-place PTX re-fetch (rare)
//         projection of the NEXT pass for the (view, plane) this lane owns, interleaved with
//         3 FFMA2 x NV x CPT/2       A + fx*B + fy*C + fxy*D
//         FADD2 + FFMA2              sum, sum of squares;  3 packed ops of variance epilogue
//         STS.32 x CPT               swizzled staging tile                    (per group, 2 buffers)
//         LDS.128 + STG.128 x CPT/4  one plane of the PREVIOUS batch leaves as 128-byte rows
//     named barrier of the group's LPP warps, swap tile buffers
#pragma once
#include "sweep_fast.cuh"

namespace d3d {

__device__ __forceinline__ unsigned smem_u32(const void* ptr) {
    return (unsigned)__cvta_generic_to_shared(ptr);
}
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
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

template <int CPT, int NV, int LPP, int MODE, bool kIeeeDiv>
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
    static_assert(KT % PB == 0, "batch must hold whole passes");
    extern __shared__ float4 smem4[];

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
    if (threadIdx.x < GROUPS * NBUF) mbar_init(smem_u32(smem4) + threadIdx.x * 8, LPP);   // one arrival per warp
    __syncthreads();

    // ---- projection ownership (see sweep_fast.cuh)
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
        tx[k] = m[3]; ty[k] = m[7]; tz[k] = m[11];
    }

    float2 rf[NP];
    {
        const float* t = p.feats + (size_t)pix * C + choff;
#pragma unroll
        for (int k = 0; k < CPT; k += 4) {
            float4 w = ldg4(t + k);
            rf[k / 2] = f2(w.x, w.y);
            rf[k / 2 + 1] = f2(w.z, w.w);
        }
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
    const float* feats_c = p.feats + choff;                // + (v+1)*HW*C per source view
    const size_t view_stride = (size_t)p.HW * C;
    const int row_bytes = p.W * C * 4;

    // ---- shared-memory addresses (bytes, shared window)
    const unsigned bar0 = smem_u32(smem4) + grp * NBUF * 8;        // this group's NBUF mbarriers
    const unsigned sm0 = smem_u32(smem4) + 256;
    const unsigned geo_w = sm0 + warp * 2 * GEO_BUF;
    unsigned gr = geo_w + q * 16;                                          // read: + t*GEO_PLANE + v*PPW*16
    unsigned gw = geo_w + GEO_BUF + po * GEO_PLANE + q * 16;               // write (other buffer): + v*PPW*16
    const unsigned tile_g = sm0 + 8 * 2 * GEO_BUF + grp * NBUF * TILE_BUF;
    const int col = wq * PPW + q;
    const unsigned tile_wr = tile_g + (choff * 32 + (col ^ ((PPW * cg) & 31))) * 4;   // + k*128 per channel row
    const int t_in_grp = wq * 32 + lane;
    unsigned tile_rd[NRO];
    long long row_off[NRO];                                                // element offset of this lane's rows
#pragma unroll
    for (int i = 0; i < NRO; ++i) {
        const int idx = t_in_grp + i * LPP * 32;
        const int row = idx >> 3, c4 = (idx & 7) * 4;
        tile_rd[i] = tile_g + (row * 32 + (c4 ^ ((PPW * (row / CPT)) & 31))) * 4;
        row_off[i] = (long long)row * p.out_sc + grp_base + c4;
    }
    const bool vec_ok = ((p.out_sc | p.out_sd) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
    const bool fast_rows = vec_ok && grp_base + 32 <= p.HW;      // whole group in range, 16-byte stores

    const size_t hyp_stride = p.perpix ? (size_t)p.HW : 1;
    const float* hp = p.hyps + (p.perpix ? (size_t)pix : 0);
    auto depth_of = [&](int plane) { return __ldg(hp + (size_t)min(plane, d1 - 1) * hyp_stride); };
    auto project_mine = [&](float4 (&gn)[KV], float depth) {
#pragma unroll
        for (int k = 0; k < KV; ++k)
            gn[k] = project_frac<kIeeeDiv>(rx[k], ry[k], rz[k], tx[k], ty[k], tz[k], depth, p);
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
    float dnext = depth_of(d0 + po);
    {
        float4 gn[KV];
        project_mine(gn, dnext);
        store_mine(gn, gw - GEO_BUF);
    }
    dnext = depth_of(d0 + PB + po);

    const float invV = 1.f / (float)(NV + 1);
    const float2 ninv = splat(-invV), pinv = splat(invV);
    unsigned tw = tile_wr;                           // staging write pointer (advances one plane per plane)
    unsigned dr[NRO];                                // drain read pointers (previous batch)
    float* optr[NRO];                                // drain global pointers
    int drain_n = 0;                                 // planes of the previous batch still to write out
#pragma unroll
    for (int i = 0; i < NRO; ++i) { dr[i] = tile_rd[i]; optr[i] = p.out; }
    unsigned gflip = GEO_BUF;                        // +/- distance between the two table buffers

    auto drain_one = [&]() {                         // one staged plane -> global, 128-byte rows
#pragma unroll
        for (int i = 0; i < NRO; ++i) {
            const float4 w = lds128(dr[i]);
            float* o = optr[i];
            if (fast_rows) {
                *reinterpret_cast<float4*>(o) = w;
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

    // batch m (planes d0 + m*KT ...) becomes the one to drain: wait until every warp of the group has staged it
    auto begin_drain = [&](int m) {
        mbar_wait(bar0 + (m % NBUF) * 8, (m / NBUF) & 1);
        const int base = d0 + m * KT;
        drain_n = min(KT, d1 - base);                // planes past the end of the sweep are dropped
#pragma unroll
        for (int i = 0; i < NRO; ++i) {
            dr[i] = tile_rd[i] + (m % NBUF) * TILE_BUF;
            optr[i] = p.out + row_off[i] + (long long)(base - p.d_begin) * p.out_sd;
        }
    };

    // Staging protocol: batch n is computed into tile buffer n % NBUF and drained, one plane per iteration,
    // while batch n+2 is computed.  Starting batch n waits for batch n-2 to be complete in every warp of the
    // group, which also implies that they are done draining batch n-4, the previous user of buffer n % NBUF.
    int n = 0;
    for (int b0 = d0; b0 < d1; b0 += KT) {
        if (n >= 2) begin_drain(n - 2);
#pragma unroll 1
        for (int t = 0; t < KT; t += PB) {
            __syncwarp();                            // table `gr` is complete; the other one is free
#pragma unroll
            for (int tt = 0; tt < PB; ++tt) {
                float4 g[NV];
                unsigned moved = 0;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    g[v] = lds128(gr + tt * GEO_PLANE + v * PPW * 16);
                    moved |= __float_as_uint(g[v].w) ^ ckey[v];
                }
                if (moved) {                         // some footprint moved: re-fetch those (in place)
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const unsigned key = __float_as_uint(g[v].w);
                        refetch_footprint(tex[v], key, ckey[v], feats_c + (size_t)(v + 1) * view_stride, p.W - 1,
                                          p.H - 1, row_bytes, C * 4);
                        ckey[v] = key;
                    }
                }
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
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const float2 tneg = __fmul2_rn(s[j], ninv);            // -sum/V
                    const float2 w = __ffma2_rn(tneg, s[j], sq[j]);        // sq - sum^2/V
                    const float2 r = __fmul2_rn(w, pinv);                  // sq/V - (sum/V)^2
                    sts32(tw + (2 * j) * 128, r.x);
                    sts32(tw + (2 * j + 1) * 128, r.y);
                }
                tw += TILE_PLANE;
                if (tt == 0) {
                    store_mine(gn, gw);
                    dnext = depth_of(b0 + t + 2 * PB + po);
                }
                if (drain_n > 0) { drain_one(); --drain_n; }
            }
            gr += gflip;                             // swap the table buffers
            gw -= gflip;
            gflip = 0u - gflip;
        }
        // batch n is staged in tile buffer n % NBUF: announce it (one arrival per warp)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + (n % NBUF) * 8);
        ++n;
        tw = tile_wr + (n % NBUF) * TILE_BUF;
    }
    // the last two batches have nothing left to hide behind
    for (int m = max(n - 2, 0); m < n; ++m) {
        for (; drain_n > 0; --drain_n) drain_one();
        begin_drain(m);
    }
    for (; drain_n > 0; --drain_n) drain_one();
}

template <int CPT, int NV, int LPP>
constexpr size_t sweep_lean_smem() {
    constexpr int PB = (LPP >= 2 * NV) ? 2 : 1;
    return 256 + (size_t)8 * 2 * PB * NV * (32 / LPP) * 16 +
           (size_t)(8 / LPP) * kLeanTileBuffers * kLeanTilePlanes * CPT * LPP * 32 * 4;
}

template <int CPT, int NV, int LPP, int MODE>
int launch_sweep_lean(const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee_div) {
    constexpr size_t smem = sweep_lean_smem<CPT, NV, LPP>();
    auto kern = ieee_div ? sweep_lean_kernel<CPT, NV, LPP, MODE, true> : sweep_lean_kernel<CPT, NV, LPP, MODE, false>;
    static bool configured[2] = {false, false};      // per instantiation, per division flavour
    if (!configured[ieee_div]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(D3D_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
        configured[ieee_div] = true;
    }
    kern<<<grid, 256, smem, stream>>>(p);
    count_launch();
    return check_launch("sweep_lean_kernel");
}

}  // namespace d3d

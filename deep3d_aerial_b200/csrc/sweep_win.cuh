// Fused plane sweep for the short per-pixel-hypothesis stages of the cascade, view-weighted product volume: sweep_acc.cuh's
// formulation (view outer, planes inner, a chunk of planes accumulated in registers, one 2x2 footprint alive per lane)
// with the footprints gathered from a SHARED-MEMORY WINDOW of the source image that TMA fetched for the whole CTA.
//
//   sweep_acc is bound by the latency of its dependent gathers: project -> corner key -> 4 x LDG.256 -> arithmetic, an L2
//   round trip on nearly every plane of every warp (some lane's footprint moves), 3 cycles of load stall per issued
//   instruction, insensitive to occupancy (profiles/ncu_r2_acc_stage3.txt).  Here the CTA is a 16 x 8 pixel tile; over
//   the chunk's planes the projections of a pixel run along a straight segment, so the footprints of the tile lie in the
//   bounding box of the segments' END points (first and last plane), a few texels larger than the tile.  Per source view:
//     1. every lane projects its two end planes; warp min/max (redux.sync) + one shared-memory exchange give the box;
//        a box larger than the window (rough depth map, oblique view) means no window for that view: global gathers;
//     2. one thread issues two `cp.async.bulk.tensor.4d` boxes (channels c0..c0+3 and c0+4..c0+7 of kWinW x kWinH
//        texels at the box origin of view v) into one of two window buffers, completing on an mbarrier -- the window of
//        view v+1 is requested before the arithmetic of view v starts, so it lands underneath it; texels outside the image
//        arrive as zeros (the tensor map's out-of-bounds fill = grid_sample's zeros padding): no border case;
//     3. a lane whose corner key changed reads its footprint with 8 x LDS.128 off ONE address register (the two 4-channel
//        planes of the window are dense 16-byte arrays: 8 consecutive pixels read 128 contiguous bytes, conflict-free);
//        a footprint outside the window (oblique views, boxes larger than the buffer, points behind the camera) takes
//        sweep_acc's global-memory path, so the window is only ever an accelerator.
//   Same operations in the same view order as the other kernels (adamvs.py:494-506); corner weights as in sweep_acc;
//   results bit-identical to sweep_acc's.  The lane's rays of all views are formed (or loaded) once and parked in shared
//   memory; CTAs of one tile (plane chunks, channel groups) are adjacent in launch order so they share windows in L2.
//
//   Measured on B200, stage-3 shape (C=8 D=8 2752x1856, V=5): with a smooth depth map every tile's box fits (20 x 10
//   texels on the synthetic rig) and the load stalls go from 2.96 to 0.63 cycles per issued instruction: 0.96 ms against
//   sweep_acc's 1.04 and sweep_direct's 1.16 (profiles/ncu_r2_win_stage3.txt) -- issue-limited at 12 warps per SM (157
//   registers: 64 accumulators + 32 footprint).  With the white-noise depth map that random regulariser outputs give
//   (bench.py cfg3) no box fits and every view falls back: 1.59 ms against sweep_acc's 1.42 (its 20 warps per SM hide
//   the gather better).  Stage-2 shape (C=16: two channel groups, the projection is done twice): 2.24 ms against
//   sweep_quad's 1.80.  Hence opt-in (`variant` 12 / 13), not the default.
#pragma once
#include <cuda.h>

#include "sweep_acc.cuh"
#include "sweep_lean.cuh"

namespace d3d {

constexpr int kWinTileW = 16, kWinTileH = 8;           // pixels per CTA (128 threads)
constexpr int kWinW = 24, kWinH = 16;                  // texels per window
constexpr unsigned kWinPlane = kWinW * kWinH * 16;     // one 4-channel plane of a window
constexpr unsigned kWinBytes = 2 * kWinPlane;

// cuTensorMapEncodeTiled through the runtime's driver entry point: the library does not link libcuda.
typedef CUresult (*WinEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline WinEncodeFn win_encode_fn() {
    static WinEncodeFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<WinEncodeFn>(f);
    }();
    return fn;
}

// texels [V][H][W][C] fp32 as a 4-D tensor (innermost first: channel, x, y, view); box = 4 channels of a window
inline int make_window_map(CUtensorMap* map, const SweepParams& p, int views) {
    WinEncodeFn enc = win_encode_fn();
    if (!enc) return -1;
    const cuuint64_t cb = (cuuint64_t)p.C * 4;
    const cuuint64_t dims[4] = {(cuuint64_t)p.C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)views};
    const cuuint64_t strides[3] = {cb, (cuuint64_t)p.W * cb, (cuuint64_t)p.HW * cb};
    const cuuint32_t box[4] = {4, kWinW, kWinH, 1}, estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.feats), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? D3D_OK : -1;
}

__device__ __forceinline__ void tma_window_plane(unsigned dst, const CUtensorMap* map, int c, int x, int y, int view, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(map), "r"(c), "r"(x), "r"(y), "r"(view), "r"(bar) : "memory");
}

template <int C, bool kPerPix, bool kIeeeDiv, int KP, int MB>
__global__ void __launch_bounds__(128, MB) sweep_win_kernel(const SweepParams p, const int nv, const int tiles_x, const int chunks,
                                                            const __grid_constant__ CUtensorMap texmap) {
    constexpr int NJ = 4;                                  // 8 channels per lane
    __shared__ __align__(128) unsigned char win[2][kWinBytes];
    __shared__ __align__(16) int box_part[2][4][4];        // [parity][warp]{min x, min y, max x, max y}
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ int win_org[2][2];
    __shared__ float ray_s[4][3][128];                     // the lane's rays of every view (a spill the lane reads back twice)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // blockIdx.x = (tile, chunk, channel group), group fastest: the CTAs that read the same windows, hypotheses, weights and
    // rays are launched together and share them through L2
    constexpr int GROUPS = C / 8;
    const int group = blockIdx.x % GROUPS, tc = blockIdx.x / GROUPS;
    const int chunk = tc % chunks, tile = tc / chunks;
    const int tile_y = tile / tiles_x, tile_x = tile - tile_y * tiles_x;
    const int px_raw = tile_x * kWinTileW + (lane & 15), py_raw = tile_y * kWinTileH + 2 * warp + (lane >> 4);
    const bool live = px_raw < p.W && py_raw < p.H;
    const int px = min(px_raw, p.W - 1), py = min(py_raw, p.H - 1);      // (dead lanes repeat a live pixel: same footprints)
    const int pix = py * p.W + px;
    const int d0 = p.d_begin + chunk * KP;
    const int d1 = min(d0 + KP, p.d_end);
    const int c0 = group * 8;
    const unsigned bar0 = smem_u32(bars), win0 = smem_u32(win);
    if (threadIdx.x < 2) mbar_init(bar0 + threadIdx.x * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");

    float2 rf[NJ];
    ldg8<8>(rf, p.feats + (size_t)pix * C + c0);
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

    const size_t view_stride = (size_t)p.HW * C;
    const int wmax = p.W - 1, hmax = p.H - 1;
    const size_t row = (size_t)p.W * C;
    // rays of every view (module.py:538), formed or loaded once, all loads in flight together
#pragma unroll 1
    for (int v = 0; v < nv; ++v) {
        const float* m = p.pose + v * 16;
        float rx = fmaf(m[2], 1.f, fmaf(m[1], (float)py, m[0] * (float)px));
        float ry = fmaf(m[6], 1.f, fmaf(m[5], (float)py, m[4] * (float)px));
        float rz = fmaf(m[10], 1.f, fmaf(m[9], (float)py, m[8] * (float)px));
        if (p.rays) {   // the reference's own rot @ [x,y,1] (cuBLAS), whatever order it rounded in
            const float* rr = p.rays + (size_t)v * 3 * p.HW + pix;
            rx = __ldg(rr); ry = __ldg(rr + p.HW); rz = __ldg(rr + 2 * (size_t)p.HW);
        }
        ray_s[v][0][threadIdx.x] = rx;
        ray_s[v][1][threadIdx.x] = ry;
        ray_s[v][2][threadIdx.x] = rz;
    }
    auto load_view = [&](int v, float& rx, float& ry, float& rz, float& tx, float& ty, float& tz) {
        const float* m = p.pose + v * 16;
        rx = ray_s[v][0][threadIdx.x]; ry = ray_s[v][1][threadIdx.x]; rz = ray_s[v][2][threadIdx.x];
        tx = m[3]; ty = m[7]; tz = m[11];                  // (module.py:539-541)
    };
    // steps 1 + 2 for view v: the tile's box from the end planes, then the two TMA boxes
    auto request_window = [&](int v) {
        float rx, ry, rz, tx, ty, tz;
        load_view(v, rx, ry, rz, tx, ty, tz);
        float4 ea, eb;
        project_pair<kIeeeDiv>(rx, ry, rz, tx, ty, tz, f2(depth[0], depth[KP - 1]), p, ea, eb);
        const unsigned ka = __float_as_uint(ea.w), kb = __float_as_uint(eb.w);
        const int xa = (int)(short)(ka & 0xffffu), ya = (int)(short)(ka >> 16);
        const int xb = (int)(short)(kb & 0xffffu), yb = (int)(short)(kb >> 16);
        const int lo_x = __reduce_min_sync(0xffffffffu, min(xa, xb)), lo_y = __reduce_min_sync(0xffffffffu, min(ya, yb));
        const int hi_x = __reduce_max_sync(0xffffffffu, max(xa, xb)), hi_y = __reduce_max_sync(0xffffffffu, max(ya, yb));
        if (lane == 0) *reinterpret_cast<int4*>(box_part[v & 1][warp]) = make_int4(lo_x, lo_y, hi_x, hi_y);
        __syncthreads();                                   // (also: every lane is done with the window buffer this view reuses)
        if (threadIdx.x == 0) {
            const int4 a = *reinterpret_cast<const int4*>(box_part[v & 1][0]), b = *reinterpret_cast<const int4*>(box_part[v & 1][1]);
            const int4 c = *reinterpret_cast<const int4*>(box_part[v & 1][2]), d = *reinterpret_cast<const int4*>(box_part[v & 1][3]);
            const int ox = min(min(a.x, b.x), min(c.x, d.x)), oy = min(min(a.y, b.y), min(c.y, d.y));
            const int hx = max(max(a.z, b.z), max(c.z, d.z)), hy = max(max(a.w, b.w), max(c.w, d.w));
            const unsigned bar = bar0 + (v & 1) * 8, dst = win0 + (v & 1) * kWinBytes;
            const bool fits = hx - ox + 2 <= kWinW && hy - oy + 2 <= kWinH;    // every footprint of the tile in one window
            // if not (rough depth map, oblique view): no window for this view, every lane gathers from global memory
            win_org[v & 1][0] = fits ? ox : -(1 << 20);    // the window's origin, for everybody (read after the wait:
            win_org[v & 1][1] = fits ? oy : -(1 << 20);    //  the arrive below releases these stores)
            if (fits) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kWinBytes) : "memory");
                tma_window_plane(dst, &texmap, c0, ox, oy, v + 1, bar);
                tma_window_plane(dst + kWinPlane, &texmap, c0 + 4, ox, oy, v + 1, bar);
            } else {
                mbar_arrive(bar);
            }
        }
    };
    request_window(0);
#pragma unroll 1
    for (int v = 0; v < nv; ++v) {
        if (v + 1 < nv) request_window(v + 1);             // lands underneath this view's arithmetic
        float rx, ry, rz, tx, ty, tz;
        load_view(v, rx, ry, rz, tx, ty, tz);
        mbar_wait(bar0 + (v & 1) * 8, (v >> 1) & 1);
        const int ox = win_org[v & 1][0], oy = win_org[v & 1][1];
        const unsigned wbase = win0 + (v & 1) * kWinBytes;
        const float wt = __ldg(p.weights + (size_t)v * p.HW + pix);
        const float* vf = p.feats + (size_t)(v + 1) * view_stride + c0;
        unsigned ckey = 0x7fff7fffu;                       // no footprint yet (x0 = y0 = 32767 cannot occur)
        float2 A[NJ], B[NJ], Cc[NJ], D[NJ];
        // two planes per packed projection chain; the chain of the NEXT pair sits in the same basic block as this plane's
        // arithmetic (after the fetch), so the scheduler interleaves its serial dependencies with the independent FMAs
        float4 e[2], en[2];
        project_pair<kIeeeDiv>(rx, ry, rz, tx, ty, tz, f2(depth[0], depth[1 < KP ? 1 : 0]), p, e[0], e[1]);
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const float4 ek = e[k & 1];
            const unsigned key = __float_as_uint(ek.w);
            if (key != ckey) {                             // the footprint moved (or first plane): fetch it
                ckey = key;
                const int x0 = (int)(short)(key & 0xffffu), y0 = (int)(short)(key >> 16);
                const int lx = x0 - ox, ly = y0 - oy;
                if ((unsigned)lx < (unsigned)(kWinW - 1) && (unsigned)ly < (unsigned)(kWinH - 1)) {     // inside the window
                    const unsigned a = wbase + (unsigned)(ly * kWinW + lx) * 16u;
                    const float4 al = lds128(a), bl = lds128(a + 16), cl = lds128(a + kWinW * 16), dl = lds128(a + kWinW * 16 + 16);
                    const float4 ah = lds128(a + kWinPlane), bh = lds128(a + kWinPlane + 16);
                    const float4 ch = lds128(a + kWinPlane + kWinW * 16), dh = lds128(a + kWinPlane + kWinW * 16 + 16);
                    A[0] = f2(al.x, al.y); A[1] = f2(al.z, al.w); A[2] = f2(ah.x, ah.y); A[3] = f2(ah.z, ah.w);
                    B[0] = f2(bl.x, bl.y); B[1] = f2(bl.z, bl.w); B[2] = f2(bh.x, bh.y); B[3] = f2(bh.z, bh.w);
                    Cc[0] = f2(cl.x, cl.y); Cc[1] = f2(cl.z, cl.w); Cc[2] = f2(ch.x, ch.y); Cc[3] = f2(ch.z, ch.w);
                    D[0] = f2(dl.x, dl.y); D[1] = f2(dl.z, dl.w); D[2] = f2(dh.x, dh.y); D[3] = f2(dh.z, dh.w);
                } else if ((unsigned)x0 < (unsigned)wmax && (unsigned)y0 < (unsigned)hmax) {   // global memory, all corners inside
                    const float* t = texel_ptr<C>(vf, (unsigned)(y0 * p.W + x0));
                    ldg8<8>(A, t);
                    ldg8<8>(B, t + C);
                    t += row;
                    ldg8<8>(Cc, t);
                    ldg8<8>(D, t + C);
                } else {                                   // global memory at the image border: clamped addresses, zeros padding
                    const int xa = min(max(x0, 0), wmax), xb = min(max(x0 + 1, 0), wmax);
                    const int ya = min(max(y0, 0), hmax), yb = min(max(y0 + 1, 0), hmax);
                    const int ra = ya * p.W, rb = yb * p.W;
                    ldg8<8>(A, texel_ptr<C>(vf, (unsigned)(ra + xa)));
                    ldg8<8>(B, texel_ptr<C>(vf, (unsigned)(ra + xb)));
                    ldg8<8>(Cc, texel_ptr<C>(vf, (unsigned)(rb + xa)));
                    ldg8<8>(D, texel_ptr<C>(vf, (unsigned)(rb + xb)));
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
            if ((k & 1) == 0 && k + 2 < KP)
                project_pair<kIeeeDiv>(rx, ry, rz, tx, ty, tz, f2(depth[k + 2], depth[k + 3 < KP ? k + 3 : k + 2]), p, en[0], en[1]);
            // grid_sample's corner weights (GridSampler.cuh: nw, ne, sw, se) from the fractions
            const float fx = ek.x, fy = ek.y, w11 = ek.z;
            const float w01 = fx - w11, w10 = fy - w11, w00 = (1.f - fx) - w10;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                float2 o = __fmul2_rn(splat(w00), A[j]);
                o = __ffma2_rn(splat(w01), B[j], o);
                o = __ffma2_rn(splat(w10), Cc[j], o);
                o = __ffma2_rn(splat(w11), D[j], o);
                acc[k][j] = __ffma2_rn(__fmul2_rn(o, rf[j]), splat(wt), acc[k][j]);   // (warped * ref) * weight
            }
            if (k & 1) { e[0] = en[0]; e[1] = en[1]; }
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

template <int C, int KP, int MB>
int launch_sweep_win(int nv, const SweepParams& p, cudaStream_t stream, bool ieee_div) {
    const int tiles_x = (p.W + kWinTileW - 1) / kWinTileW, tiles_y = (p.H + kWinTileH - 1) / kWinTileH;
    const int chunks = (p.d_end - p.d_begin + KP - 1) / KP;
    const long long ctas = (long long)tiles_x * tiles_y * chunks * (C / 8);
    if (ctas > 0x7fffffffLL || nv > 4) return -1;
    CUtensorMap map;
    if (make_window_map(&map, p, nv + 1)) return -1;
    const dim3 grid((unsigned)ctas);
    if (p.perpix) {
        if (ieee_div) sweep_win_kernel<C, true, true, KP, MB><<<grid, 128, 0, stream>>>(p, nv, tiles_x, chunks, map);
        else sweep_win_kernel<C, true, false, KP, MB><<<grid, 128, 0, stream>>>(p, nv, tiles_x, chunks, map);
    } else {
        if (ieee_div) sweep_win_kernel<C, false, true, KP, MB><<<grid, 128, 0, stream>>>(p, nv, tiles_x, chunks, map);
        else sweep_win_kernel<C, false, false, KP, MB><<<grid, 128, 0, stream>>>(p, nv, tiles_x, chunks, map);
    }
    count_launch();
    return check_launch("sweep_win_kernel");
}

// returns -1 when the shape is not covered
inline int sweep_win_dispatch(int nv, const SweepParams& p, cudaStream_t stream, bool ieee_div) {
    if (p.pooled) return -1;                         // views addressed by stride here: dense [V,H,W,C] texels only
    if (p.W > 16000 || p.H > 16000 || nv < 1 || !p.weights) return -1;
    if (reinterpret_cast<uintptr_t>(p.feats) & 31) return -1;      // 256-bit texel loads, 16-byte TMA rows
    if (p.C == 8) return (p.flags & 1) ? launch_sweep_win<8, 4, 4>(nv, p, stream, ieee_div) : launch_sweep_win<8, 8, 3>(nv, p, stream, ieee_div);
    if (p.C == 16) return (p.flags & 1) ? launch_sweep_win<16, 4, 4>(nv, p, stream, ieee_div) : launch_sweep_win<16, 8, 3>(nv, p, stream, ieee_div);
    return -1;
}

}  // namespace d3d

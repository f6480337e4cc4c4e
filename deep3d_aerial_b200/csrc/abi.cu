// C-ABI plumbing: error state, argument validation and kernel selection for d3d_cost_volume.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace d3d {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(D3D_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
    return D3D_OK;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// per-mode launchers (sweep_base_*.cu, sweep_*.cu)
int sweep_base_warp(int cpt, int nv, const SweepParams& p, dim3 grid, cudaStream_t stream);
int sweep_base_variance(int cpt, int nv, const SweepParams& p, dim3 grid, cudaStream_t stream);
int sweep_base_group_corr(int cpt, int nv, const SweepParams& p, dim3 grid, cudaStream_t stream);
int sweep_base_weighted_product(int cpt, int nv, const SweepParams& p, dim3 grid, cudaStream_t stream);
int sweep_base_pair_mean(int cpt, int nv, const SweepParams& p, dim3 grid, cudaStream_t stream);
int sweep_lean_variance(int cpt, int nv, const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee_div);
int sweep_quad_variance(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee_div);
int sweep_ws_variance(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream);
int sweep_pre_variance(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream);
int sweep_quad_group_corr(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee_div);
int sweep_quad_weighted_product(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee_div);
int sweep_quad_pair_mean(int nv, const SweepParams& p, dim3 grid, cudaStream_t stream, bool ieee_div);
int sweep_direct_variance(int nv, const SweepParams& p, cudaStream_t stream);
int sweep_direct_weighted_product(int nv, const SweepParams& p, cudaStream_t stream);
int sweep_acc_weighted_product(int nv, const SweepParams& p, cudaStream_t stream, bool ieee_div);
int sweep_win_weighted_product(int nv, const SweepParams& p, cudaStream_t stream, bool ieee_div);

static int sm_count() {
    static int cached = 0;
    if (!cached) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || cached <= 0)
            cached = 148;
    }
    return cached;
}

static int ilog2_exact(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return (1 << l) == v ? l : -1;
}

}  // namespace d3d

using namespace d3d;

extern "C" const char* d3d_last_error(void) { return g_err; }
extern "C" int d3d_version(void) { return D3D_VERSION; }
extern "C" int64_t d3d_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" int32_t d3d_abi_sizeof(int32_t which) {
    switch (which) {
        case 0: return (int32_t)sizeof(D3dCostVolumeArgs);
        case 1: return (int32_t)sizeof(D3dRegressArgs);
        case 2: return (int32_t)sizeof(D3dSamplesArgs);
        case 3: return (int32_t)sizeof(D3dFuseArgs);
        default: return -1;
    }
}

extern "C" int d3d_cost_volume(const D3dCostVolumeArgs* a, void* cuda_stream) {
    if (!a) return fail(D3D_ERR_BAD_ARGUMENT, "d3d_cost_volume: args is NULL");
    if (a->struct_size != sizeof(D3dCostVolumeArgs))
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_cost_volume: struct_size %u != %zu", a->struct_size,
                    sizeof(D3dCostVolumeArgs));
    if (a->mode < D3D_AGG_WARP || a->mode > D3D_AGG_PAIR_MEAN)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_cost_volume: unknown mode %d", a->mode);
    if (a->height < 2 || a->width < 2 || a->num_depth <= 0 || a->channels <= 0)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_cost_volume: bad extent C=%d D=%d H=%d W=%d (H, W >= 2)", a->channels,
                    a->num_depth, a->height, a->width);
    const int nv = a->num_views - 1;
    if (nv < 1) return fail(D3D_ERR_BAD_ARGUMENT, "d3d_cost_volume: need at least one source view (V=%d)", a->num_views);
    if (nv > kMaxSrc) return fail(D3D_ERR_UNSUPPORTED, "d3d_cost_volume: V=%d exceeds %d views", a->num_views, kMaxSrc + 1);
    if (a->mode == D3D_AGG_WARP && nv != 1)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_cost_volume: D3D_AGG_WARP takes exactly one source view");
    if (!a->feats || !a->pose || !a->hyps || !a->out)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_cost_volume: feats/pose/hyps/out is NULL");
    if (a->mode == D3D_AGG_WEIGHTED_PRODUCT && !a->weights)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_cost_volume: WEIGHTED_PRODUCT needs weights");
    if ((long long)a->height * a->width > INT32_MAX / 2)
        return fail(D3D_ERR_UNSUPPORTED, "d3d_cost_volume: H*W too large");
    const int C = a->channels;
    if (C % 4) return fail(D3D_ERR_UNSUPPORTED, "d3d_cost_volume: C=%d is not a multiple of 4", C);
    // channels per lane: 8 when the footprint cache (V-1)*4*8 registers fits, else 4
    int cpt = (C % 8 == 0 && nv <= 4) ? 8 : 4;
    int lpp_log2 = ilog2_exact(C / cpt);
    if (lpp_log2 < 0 && cpt == 8) { cpt = 4; lpp_log2 = ilog2_exact(C / 4); }
    if (lpp_log2 < 0 || lpp_log2 > 5)
        return fail(D3D_ERR_UNSUPPORTED, "d3d_cost_volume: C=%d: C/4 or C/8 must be a power of two <= 32", C);
    if (a->mode == D3D_AGG_GROUP_CORR) {
        if (a->groups <= 0 || C % a->groups) return fail(D3D_ERR_BAD_ARGUMENT, "d3d_cost_volume: C=%d not divisible into G=%d groups", C, a->groups);
        const int gs = C / a->groups;
        if (ilog2_exact(gs) < 0) return fail(D3D_ERR_UNSUPPORTED, "d3d_cost_volume: group size %d is not a power of two", gs);
    }
    const int d_begin = a->d_begin;
    const int d_count = a->d_count <= 0 ? a->num_depth - d_begin : a->d_count;
    if (d_begin < 0 || d_count <= 0 || d_begin + d_count > a->num_depth)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_cost_volume: plane slice [%d,+%d) outside 0..%d", d_begin, d_count,
                    a->num_depth);

    SweepParams p;
    p.feats = a->feats; p.pose = a->pose; p.hyps = a->hyps; p.weights = a->weights; p.out = a->out; p.rays = a->rays;
    p.C = C; p.H = a->height; p.W = a->width; p.HW = a->height * a->width;
    p.out_sd = a->out_stride_d > 0 ? a->out_stride_d : p.HW;
    p.out_sc = a->out_stride_c > 0 ? a->out_stride_c : (long long)d_count * p.out_sd;
    p.d_begin = d_begin; p.d_end = d_begin + d_count;
    p.perpix = a->hyps_per_pixel != 0;
    p.groups = a->groups; p.eps_num = a->eps_in_numerator != 0;
    p.flags = 0;
    p.inv_half_w = 1.f / ((float)(a->width - 1) / 2.f);
    p.inv_half_h = 1.f / ((float)(a->height - 1) / 2.f);
    p.wm1 = (float)(a->width - 1); p.hm1 = (float)(a->height - 1);
    // where each view starts in `feats` (texel index): the dense block, or slots of a texel pool
    p.pooled = 0;
    for (int v = 0; v <= kMaxSrc; ++v) p.view_tex[v] = 0;
    if (a->texel_slots < 0) return fail(D3D_ERR_BAD_ARGUMENT, "d3d_cost_volume: texel_slots %d < 0", a->texel_slots);
    for (int v = 0; v <= nv; ++v) {
        int slot = v;
        if (a->texel_slots > 0) {
            slot = a->view_slot[v];
            if (slot < 0 || slot >= a->texel_slots)
                return fail(D3D_ERR_BAD_ARGUMENT, "d3d_cost_volume: view_slot[%d] = %d outside the pool of %d", v, slot, a->texel_slots);
            if (slot != v) p.pooled = 1;
        }
        const long long start = (long long)slot * p.HW;
        if (start + p.HW > INT32_MAX) return fail(D3D_ERR_UNSUPPORTED, "d3d_cost_volume: texel pool beyond 2^31 texels (slot %d)", slot);
        p.view_tex[v] = (int)start;
    }

    // grid: x = pixel tiles (8 warps x 32/LPP pixels), y = depth chunks.  Depth is only split when
    // the pixel tiles alone leave SMs idle (each chunk re-warms its footprint registers).
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    auto make_grid = [&](int lpp_l2, dim3& grid) -> int {
        p.lpp_log2 = lpp_l2;
        const int pix_per_cta = 8 * (32 >> lpp_l2);
        const long long tiles = ((long long)p.HW + pix_per_cta - 1) / pix_per_cta;
        int chunks = 1;
        const long long want = 4LL * sm_count();
        if (tiles < want) chunks = (int)((want + tiles - 1) / tiles);
        if (chunks > d_count) chunks = d_count;
        p.d_chunk = (d_count + chunks - 1) / chunks;
        chunks = (d_count + p.d_chunk - 1) / p.d_chunk;
        if (tiles > 0x7fffffffLL || chunks > 65535) return fail(D3D_ERR_UNSUPPORTED, "d3d_cost_volume: grid too large");
        grid = dim3((unsigned)tiles, (unsigned)chunks);
        return D3D_OK;
    };
    dim3 grid;

    // variant 0: production kernels (sweep_quad for 32- / 16-channel features; 8 channels: sweep_acc for the weighted product,
    //            sweep_direct for the variance; sweep_lean / sweep_base for
    //            what those are not instantiated for);  1: baseline kernel (sweep_base) everywhere;  2: variant 0 with
    //            __fdiv_rn instead of the shared-reciprocal division;  6: sweep_lean where variant 0 picks sweep_quad;
    //            7: sweep_quad spelled out;  8 / 9: the TMA-prefetch experiments (sweep_ws / sweep_pre);  10 - 14: sweep_acc / sweep_win
    //            A/B (below);  >= 16: A/B flag
    //            bits of the production kernel (variant - 16 -> SweepParams.flags).
    // Shapes a kernel is not instantiated for fall through to the next one.
    int variant = a->variant;
    if (variant >= 16) { p.flags = variant - 16; variant = 0; }   // A/B switches of the production kernel
    // view-weighted product volume of the short per-pixel-hypothesis stages (cascade stages 2 and 3: 16 / 8 channels):
    // planes accumulated in registers, one footprint alive per lane, gathered from a TMA-fetched shared-memory window of
    // the source view (sweep_win.cuh) or straight from global memory (sweep_acc.cuh).  variant 12 / 13: sweep_win with
    // 8 / 4 planes per lane;  10 / 14: sweep_acc (4 planes per lane / its alternative configuration);  11: the kernels
    // that served these shapes before (sweep_direct, sweep_quad).
    if (a->mode == D3D_AGG_WEIGHTED_PRODUCT && (C == 8 || C == 16 || C == 32) && variant != 1 && variant != 11) {
        int rc = -1;
        const bool ieee = variant == 2;
        const int flags = p.flags;
        if ((variant == 12 || variant == 13) && C != 32) {
            p.flags = variant == 13;
            rc = sweep_win_weighted_product(nv, p, stream, false);
        } else if (variant == 10 || variant == 14) {
            p.flags = variant == 14;
            rc = sweep_acc_weighted_product(nv, p, stream, false);
        } else if ((variant == 0 || variant == 2) && C == 8) {
            rc = sweep_acc_weighted_product(nv, p, stream, ieee);
        }
        p.flags = flags;
        if (rc >= 0) return rc;
    }
    if (variant >= 10 && variant <= 14) variant = 0;
    // 8-channel features (cascade stage 3): direct gather, one lane per pixel (sweep_direct.cuh)
    if (variant == 0 && C == 8 && nv <= 4 && (a->mode == D3D_AGG_VARIANCE || a->mode == D3D_AGG_WEIGHTED_PRODUCT)) {
        const int rc = a->mode == D3D_AGG_VARIANCE ? sweep_direct_variance(nv, p, stream)
                                                   : sweep_direct_weighted_product(nv, p, stream);
        if (rc >= 0) return rc;
    }
    if (variant != 1 && a->mode == D3D_AGG_VARIANCE && nv <= 4) {
        const bool want8 = false;                 // (8 channels per lane: only where 4 would need more than 8 lanes per pixel, C = 64)
        int fcpt = want8 ? 8 : 4;
        int fl2 = ilog2_exact(C / fcpt);
        if ((fl2 < 0 || fl2 > 3) && C % 8 == 0) { fcpt = 8; fl2 = ilog2_exact(C / 8); }
        if (fl2 >= 0 && fl2 <= 3) {
            if (int rc = make_grid(fl2, grid)) return rc;
            int rc = -1;
            // 32-channel features: the four-planes-per-pass kernel (variant 6 keeps the two-plane one for A/B)
            // variant 8: the warp-specialised kernel (sweep_ws.cuh: producer warps project and prefetch moved footprints
            // into shared-memory slots with TMA, consumer warps only do the packed arithmetic).  Parity-green; 6.5 ms
            // against sweep_quad's 5.46 ms at cfg2 so far (producer-bound), hence not the default yet.  7 = sweep_quad.
            if (variant == 8 && fcpt == 4 && C == 32)
                rc = sweep_ws_variance(nv, p, grid, stream);
            // variant 9: sweep_quad's formulation with the moved footprints prefetched by TMA a pass ahead (sweep_pre.cuh)
            if (variant == 9 && fcpt == 4 && C == 32)
                rc = sweep_pre_variance(nv, p, grid, stream);
            if (rc < 0 && (variant == 0 || variant == 2 || variant == 7 || variant == 8 || variant == 9) && fcpt == 4 && (C == 32 || C == 16))
                rc = sweep_quad_variance(nv, p, grid, stream, variant == 2);
            if (rc < 0) rc = sweep_lean_variance(fcpt, nv, p, grid, stream, variant == 2);
            if (rc >= 0) return rc;
        }
    }
    // 32-channel features, <= 4 source views: the four-planes-per-pass kernel also builds the group-wise
    // correlation and the weighted-product volumes
    if ((variant == 0 || variant == 2) && (C == 32 || C == 16) && nv <= 4 &&
        (a->mode == D3D_AGG_GROUP_CORR || a->mode == D3D_AGG_WEIGHTED_PRODUCT || a->mode == D3D_AGG_PAIR_MEAN)) {
        if (int rc = make_grid(ilog2_exact(C / 4), grid)) return rc;
        const bool ieee = variant == 2;
        const int rc = a->mode == D3D_AGG_GROUP_CORR ? sweep_quad_group_corr(nv, p, grid, stream, ieee)
                     : a->mode == D3D_AGG_PAIR_MEAN  ? sweep_quad_pair_mean(nv, p, grid, stream, ieee)
                                                     : sweep_quad_weighted_product(nv, p, grid, stream, ieee);
        if (rc >= 0) return rc;
    }
    if (int rc = make_grid(lpp_log2, grid)) return rc;
    switch (a->mode) {
        case D3D_AGG_WARP: return sweep_base_warp(cpt, nv, p, grid, stream);
        case D3D_AGG_VARIANCE: return sweep_base_variance(cpt, nv, p, grid, stream);
        case D3D_AGG_GROUP_CORR: return sweep_base_group_corr(cpt, nv, p, grid, stream);
        case D3D_AGG_WEIGHTED_PRODUCT: return sweep_base_weighted_product(cpt, nv, p, grid, stream);
        default: return sweep_base_pair_mean(cpt, nv, p, grid, stream);
    }
}

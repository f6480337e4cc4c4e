// Per-stage depth-hypothesis resampling and the feature relayout.
//
//   d3d_depth_samples  get_depth_range_samples / get_cur_depth_range_samples (module.py:616-650) and the
//                      Cas-MVSNet / RED-Net stage glue around it (cas_mvsnet.py:206-226,
//                      msrednet.py:495-515): bilinear up-sampling of the previous depth to full
//                      resolution, sampling there, tri-linear down-sampling to the stage -- fused so
//                      the full-resolution [D,H,W] tensor is never written.
//   d3d_nchw_to_nhwc   [C,H,W] -> [H,W,C] so that one texel (all channels of one pixel) is contiguous.
#include "common.cuh"

namespace d3d {

struct SamplesParams {
    const float* __restrict__ cur;
    const float* __restrict__ spread;
    float* __restrict__ out;
    int mode, D, H, W, HW;
    int sh, sw, fh, fw;
    float half_span;        // (float)(D/2 * interval)
    float dmin, dmax;
    float up_h, up_w;       // sh/fh, sw/fw : bilinear up-sampling scales (align_corners=False)
    float dn_h, dn_w;       // fh/H,  fw/W  : tri-linear down-sampling scales
};

// ATen area_pixel_compute_source_index (align_corners=False): src = scale*(dst+0.5)-0.5, clamped at 0.
__device__ __forceinline__ void tap1d(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
    float s = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
    i0 = min((int)s, in_size - 1);
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    l1 = s - (float)i0;
    l0 = 1.f - l1;
}

__device__ __forceinline__ float upsampled_depth(const SamplesParams& p, int fy, int fx) {
    int y0, y1, x0, x1;
    float h0, h1, w0, w1;
    tap1d(p.up_h, fy, p.sh, y0, y1, h0, h1);
    tap1d(p.up_w, fx, p.sw, x0, x1, w0, w1);
    const float* c = p.cur;
    return h0 * (w0 * __ldg(c + y0 * p.sw + x0) + w1 * __ldg(c + y0 * p.sw + x1)) +
           h1 * (w0 * __ldg(c + y1 * p.sw + x0) + w1 * __ldg(c + y1 * p.sw + x1));
}

__global__ void __launch_bounds__(256) depth_samples_kernel(const SamplesParams p) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p.HW) return;
    const float dm1 = (float)(p.D - 1);
    if (p.mode == D3D_SAMPLES_RANGE || (p.mode == D3D_SAMPLES_CASCADE && p.cur == nullptr)) {
        // module.py:637-645; the trilinear resize of a per-plane constant is that constant
        float step = __fdiv_rn(__fsub_rn(p.dmax, p.dmin), dm1);
        for (int k = 0; k < p.D; ++k)
            p.out[(size_t)k * p.HW + pix] = __fadd_rn(p.dmin, __fmul_rn((float)k, step));
        return;
    }
    if (p.mode == D3D_SAMPLES_AROUND) {
        float c = __ldg(p.cur + pix);
        float lo = __fsub_rn(c, p.half_span), hi = __fadd_rn(c, p.half_span);
        float step = __fdiv_rn(__fsub_rn(hi, lo), dm1);
        for (int k = 0; k < p.D; ++k)
            p.out[(size_t)k * p.HW + pix] = __fadd_rn(lo, __fmul_rn((float)k, step));
        return;
    }
    if (p.mode == D3D_SAMPLES_SPREAD) {          // ucsnet.py:41-51: low + step * i + eps, every step rounded as torch rounds it
        const float c = __ldg(p.cur + pix), e = __ldg(p.spread + pix);
        const float lo = __fsub_rn(c, e), hi = __fadd_rn(c, e);
        const float step = __fdiv_rn(__fsub_rn(hi, lo), dm1);
        for (int k = 0; k < p.D; ++k)
            p.out[(size_t)k * p.HW + pix] = __fadd_rn(__fadd_rn(lo, __fmul_rn(step, (float)k)), 1e-12f);
        return;
    }
    // CASCADE: four full-resolution taps of the tri-linear down-sampling (the depth axis keeps its
    // size, so its lambda is exactly 0/1), each an up-sampled previous depth
    const int y = pix / p.W, x = pix - y * p.W;
    int y0, y1, x0, x1;
    float h0, h1, w0, w1;
    tap1d(p.dn_h, y, p.fh, y0, y1, h0, h1);
    tap1d(p.dn_w, x, p.fw, x0, x1, w0, w1);
    float lo[4], st[4];
    const int ys[4] = {y0, y0, y1, y1}, xs[4] = {x0, x1, x0, x1};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float c = upsampled_depth(p, ys[j], xs[j]);
        lo[j] = __fsub_rn(c, p.half_span);
        float hi = __fadd_rn(c, p.half_span);
        st[j] = __fdiv_rn(__fsub_rn(hi, lo[j]), dm1);
    }
    for (int k = 0; k < p.D; ++k) {
        float s[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] = __fadd_rn(lo[j], __fmul_rn((float)k, st[j]));
        p.out[(size_t)k * p.HW + pix] = h0 * (w0 * s[0] + w1 * s[1]) + h1 * (w0 * s[2] + w1 * s[3]);
    }
}

constexpr int kTilePix = 64;

__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                          int C, int HW) {
    extern __shared__ float tile[];   // [C][kTilePix+1]
    const long long p0 = (long long)blockIdx.x * kTilePix;
    const int n = C * kTilePix;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int c = i / kTilePix, px = i - c * kTilePix;
        long long p = p0 + px;
        tile[c * (kTilePix + 1) + px] = (p < HW) ? __ldg(in + (size_t)c * HW + p) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int px = i / C, c = i - px * C;
        long long p = p0 + px;
        if (p < HW) out[(size_t)p * C + c] = tile[c * (kTilePix + 1) + px];
    }
}

// Register transpose, no shared memory: a thread owns 4 consecutive pixels x 8 consecutive channels.  It reads
// eight 16-byte pieces (4 pixels of one channel each; a warp reads 512 contiguous bytes per channel) and writes
// four 32-byte pieces (8 channels of one pixel each: whole sectors; with C = 8 a warp writes 4 KB contiguous).
// Measured against the shared-memory tile version (kept for C % 8 != 0 or unaligned maps): 37 -> 25.5 us per
// 32-channel 688x464 map, 97 -> 82 us per 8-channel 2752x1856 map.
__global__ void __launch_bounds__(256) nchw_to_nhwc_vec_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                              int C, int HW) {
    const long long p4 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;      // first of this thread's 4 pixels
    if (p4 >= HW) return;
    const int c0 = blockIdx.y * 8;
    float4 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = ldg4(in + (size_t)(c0 + c) * HW + p4);
    float* o = out + (size_t)p4 * C + c0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float e[8] = {k == 0 ? v[0].x : k == 1 ? v[0].y : k == 2 ? v[0].z : v[0].w,
                            k == 0 ? v[1].x : k == 1 ? v[1].y : k == 2 ? v[1].z : v[1].w,
                            k == 0 ? v[2].x : k == 1 ? v[2].y : k == 2 ? v[2].z : v[2].w,
                            k == 0 ? v[3].x : k == 1 ? v[3].y : k == 2 ? v[3].z : v[3].w,
                            k == 0 ? v[4].x : k == 1 ? v[4].y : k == 2 ? v[4].z : v[4].w,
                            k == 0 ? v[5].x : k == 1 ? v[5].y : k == 2 ? v[5].z : v[5].w,
                            k == 0 ? v[6].x : k == 1 ? v[6].y : k == 2 ? v[6].z : v[6].w,
                            k == 0 ? v[7].x : k == 1 ? v[7].y : k == 2 ? v[7].z : v[7].w};
        *reinterpret_cast<float4*>(o + (size_t)k * C) = make_float4(e[0], e[1], e[2], e[3]);
        *reinterpret_cast<float4*>(o + (size_t)k * C + 4) = make_float4(e[4], e[5], e[6], e[7]);
    }
}

// Warp-tile version (round 2): a warp moves 32 pixels x C channels.  It reads C rows of 32 consecutive pixels (one
// 128-byte line per load instruction), parks them in a padded shared-memory tile and writes the 32*C output floats as C
// store instructions of 32 CONSECUTIVE floats each (= 32/C whole texels): every global access of the kernel is a full
// 128-byte line.  The register-transpose kernel above writes 32-byte segments at a 128-byte stride (2.3 TB/s class
// stores, r1_microbench.txt): 25.6 us per 32-channel 688x464 map = 3.2 TB/s.
template <int C>
__global__ void __launch_bounds__(256) nchw_to_nhwc_tile_kernel(const float* __restrict__ in, float* __restrict__ out, int HW) {
    __shared__ float tile[8][C][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long pix0 = ((long long)blockIdx.x * 8 + warp) * 32;
    if (pix0 >= HW) return;
    const int n = min(32, (int)(HW - pix0));               // pixels of this tile (ragged tail)
    float v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = lane < n ? __ldg(in + (size_t)c * HW + pix0 + lane) : 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) tile[warp][c][lane] = v[c];
    __syncwarp();
    float* o = out + (size_t)pix0 * C;
#pragma unroll
    for (int j = 0; j < C; ++j) {
        const int f = j * 32 + lane;                       // flat index inside the tile's output: pixel f / C, channel f % C
        if (f < n * C) o[f] = tile[warp][f % C][f / C];
    }
}

}  // namespace d3d

using namespace d3d;

extern "C" int d3d_depth_samples(const D3dSamplesArgs* a, void* cuda_stream) {
    if (!a) return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_samples: args is NULL");
    if (a->struct_size != sizeof(D3dSamplesArgs))
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_samples: struct_size %u != %zu", a->struct_size,
                    sizeof(D3dSamplesArgs));
    if (a->mode < D3D_SAMPLES_RANGE || a->mode > D3D_SAMPLES_SPREAD)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_samples: unknown mode %d", a->mode);
    if (a->num_depth < 2 || a->height <= 0 || a->width <= 0)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_samples: need D >= 2 and positive extent (D=%d H=%d W=%d)",
                    a->num_depth, a->height, a->width);
    if ((long long)a->height * a->width > INT32_MAX)
        return fail(D3D_ERR_UNSUPPORTED, "d3d_depth_samples: H*W exceeds 2^31-1");
    if (!a->out) return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_samples: out is NULL");
    if (a->mode == D3D_SAMPLES_AROUND && !a->cur)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_samples: AROUND needs cur");
    if (a->mode == D3D_SAMPLES_SPREAD && (!a->cur || !a->spread))
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_samples: SPREAD needs cur and spread");
    if (a->mode == D3D_SAMPLES_CASCADE && a->cur &&
        (a->src_height <= 0 || a->src_width <= 0 || a->full_height <= 0 || a->full_width <= 0))
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_depth_samples: CASCADE needs src and full extents");
    SamplesParams p;
    p.cur = a->cur; p.spread = a->spread; p.out = a->out; p.mode = a->mode; p.D = a->num_depth;
    p.H = a->height; p.W = a->width; p.HW = a->height * a->width;
    p.sh = a->src_height; p.sw = a->src_width; p.fh = a->full_height; p.fw = a->full_width;
    p.half_span = (float)((double)a->num_depth / 2.0 * a->interval);
    p.dmin = a->dmin; p.dmax = a->dmax;
    p.up_h = p.fh > 0 ? (float)p.sh / (float)p.fh : 1.f;
    p.up_w = p.fw > 0 ? (float)p.sw / (float)p.fw : 1.f;
    p.dn_h = (float)p.fh / (float)p.H;
    p.dn_w = (float)p.fw / (float)p.W;
    depth_samples_kernel<<<(p.HW + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(p);
    count_launch();
    return check_launch("depth_samples_kernel");
}

// rot @ [x,y,1] in the order the sweep kernels use when they form the rays themselves (sweep_quad.cuh, sweep_direct.cuh)
__global__ void __launch_bounds__(256) pixel_rays_kernel(const float* __restrict__ pose, int W, int HW, float* __restrict__ out) {
    const int pix = blockIdx.x * 256 + threadIdx.x;
    if (pix >= HW) return;
    const int py = pix / W, px = pix - py * W;
    const float* m = pose + blockIdx.y * 16;
    float* o = out + (size_t)blockIdx.y * 3 * HW + pix;
    o[0] = fmaf(m[2], 1.f, fmaf(m[1], (float)py, m[0] * (float)px));
    o[HW] = fmaf(m[6], 1.f, fmaf(m[5], (float)py, m[4] * (float)px));
    o[2 * (size_t)HW] = fmaf(m[10], 1.f, fmaf(m[9], (float)py, m[8] * (float)px));
}

// homo_warping_double (module.py:560-601): one thread per (pixel, plane); coordinates in fp64 exactly as the reference
// forms them (cuBLAS dgemm accumulates the 3-term product with FMAs: fma(r2, 1, fma(r1, y, r0*x))), the normalised
// grid cast to fp32, then ATen's fp32 unnormalisation and bilinear weights (GridSampler.cu: nw = (ix_se-ix)*(iy_se-iy)...).
__global__ void __launch_bounds__(256) warp_f64_kernel(const float* __restrict__ tex, const double* __restrict__ pose,
                                                       const float* __restrict__ hyps, int perpix, int C, int D, int H, int W,
                                                       float* __restrict__ out) {
    const int HW = H * W;
    const int pix = blockIdx.x * 256 + threadIdx.x;
    const int d = blockIdx.y;
    if (pix >= HW) return;
    const int py = pix / W, px = pix - py * W;
    const double x = (double)px, y = (double)py;
    const double depth = (double)__ldg(hyps + (perpix ? (size_t)d * HW + pix : (size_t)d));
    double q[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double ray = fma(pose[4 * r + 2], 1.0, fma(pose[4 * r + 1], y, __dmul_rn(pose[4 * r], x)));
        q[r] = __dadd_rn(__dmul_rn(ray, depth), pose[4 * r + 3]);
    }
    const double hw2 = (double)(W - 1) / 2.0, hh2 = (double)(H - 1) / 2.0;       // python floats: (width - 1) / 2
    const float gx = (float)__dsub_rn(__ddiv_rn(__ddiv_rn(q[0], q[2]), hw2), 1.0);
    const float gy = (float)__dsub_rn(__ddiv_rn(__ddiv_rn(q[1], q[2]), hh2), 1.0);
    float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(W - 1));  // GridSampler.h:27-32, align_corners
    float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(H - 1));
    ix = fminf(fmaxf(ix, -2.f), (float)(W - 1) + 2.f);                           // NaN / far outside: every corner out
    iy = fminf(fmaxf(iy, -2.f), (float)(H - 1) + 2.f);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float ax = __fsub_rn(__fadd_rn(fx0, 1.f), ix), bx = __fsub_rn(ix, fx0);
    const float ay = __fsub_rn(__fadd_rn(fy0, 1.f), iy), by = __fsub_rn(iy, fy0);
    const float w[4] = {__fmul_rn(ax, ay), __fmul_rn(bx, ay), __fmul_rn(ax, by), __fmul_rn(bx, by)};
    const float* t[4];
    bool in[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
        in[k] = (unsigned)xx < (unsigned)W && (unsigned)yy < (unsigned)H;
        t[k] = tex + ((size_t)(in[k] ? yy : 0) * W + (in[k] ? xx : 0)) * C;
    }
    float* o = out + (size_t)d * HW + pix;
    for (int c = 0; c < C; ++c) {
        float acc = 0.f;                                                       // ATen accumulates nw, ne, sw, se in order
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (in[k]) acc = __fadd_rn(acc, __fmul_rn(__ldg(t[k] + c), w[k]));
        o[(size_t)c * D * HW] = acc;
    }
}

extern "C" int d3d_homo_warp_f64(const float* texels, const double* pose64, const float* hyps, int32_t hyps_per_pixel,
                                 int32_t channels, int32_t num_depth, int32_t height, int32_t width, float* out,
                                 void* cuda_stream) {
    if (!texels || !pose64 || !hyps || !out) return fail(D3D_ERR_BAD_ARGUMENT, "d3d_homo_warp_f64: texels/pose64/hyps/out is NULL");
    if (channels <= 0 || num_depth <= 0 || height < 2 || width < 2 || num_depth > 65535)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_homo_warp_f64: bad extent C=%d D=%d H=%d W=%d", channels, num_depth, height, width);
    if ((long long)height * width > INT32_MAX / 2) return fail(D3D_ERR_UNSUPPORTED, "d3d_homo_warp_f64: H*W too large");
    const int hw = height * width;
    warp_f64_kernel<<<dim3((hw + 255) / 256, num_depth), 256, 0, (cudaStream_t)cuda_stream>>>(
        texels, pose64, hyps, hyps_per_pixel != 0, channels, num_depth, height, width, out);
    count_launch();
    return check_launch("warp_f64_kernel");
}

extern "C" int d3d_pixel_rays(const float* pose, int32_t num_src, int32_t height, int32_t width, float* out, void* cuda_stream) {
    if (!pose || !out) return fail(D3D_ERR_BAD_ARGUMENT, "d3d_pixel_rays: pose/out is NULL");
    if (num_src < 1 || num_src > 65535 || height < 1 || width < 1 || (long long)height * width > INT32_MAX / 2)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_pixel_rays: bad extent V-1=%d H=%d W=%d", num_src, height, width);
    const int hw = height * width;
    pixel_rays_kernel<<<dim3((hw + 255) / 256, num_src), 256, 0, (cudaStream_t)cuda_stream>>>(pose, width, hw, out);
    count_launch();
    return check_launch("pixel_rays_kernel");
}

// F.interpolate(x, [H, W], mode="bilinear", align_corners=False) for a stack of maps (the AdaMVS pair confidences between
// stages, adamvs.py:291-302): ATen's upsample_bilinear2d formula -- src = scale * (dst + 0.5) - 0.5 clamped at 0, i1 =
// (int)src, lambda1 = src - i1, out = h0 * (w0 * t00 + w1 * t01) + h1 * (w0 * t10 + w1 * t11) -- one thread per output
// pixel, every map of the stack from the same taps.
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int hi,
                                                              int wi, int ho, int wo, float sh, float sw) {
    const int pix = blockIdx.x * 256 + threadIdx.x;
    if (pix >= ho * wo) return;
    const int y = pix / wo, x = pix - y * wo;
    const float sy = fmaxf(sh * ((float)y + 0.5f) - 0.5f, 0.f), sx = fmaxf(sw * ((float)x + 0.5f) - 0.5f, 0.f);
    const int y1 = (int)sy, x1 = (int)sx;
    const int yp = (y1 < hi - 1) ? 1 : 0, xp = (x1 < wi - 1) ? 1 : 0;
    const float h1 = sy - (float)y1, h0 = 1.f - h1, w1 = sx - (float)x1, w0 = 1.f - w1;
    const float* a = src + (size_t)y1 * wi + x1;
    const float* b = a + (size_t)yp * wi;
    const size_t in_map = (size_t)hi * wi, out_map = (size_t)ho * wo;
    for (int c = 0; c < n; ++c, a += in_map, b += in_map)
        dst[(size_t)c * out_map + pix] = h0 * (w0 * __ldg(a) + w1 * __ldg(a + xp)) + h1 * (w0 * __ldg(b) + w1 * __ldg(b + xp));
}

extern "C" int d3d_resize_bilinear(const float* in, float* out, int32_t maps, int32_t in_height, int32_t in_width,
                                   int32_t out_height, int32_t out_width, void* cuda_stream) {
    if (!in || !out) return fail(D3D_ERR_BAD_ARGUMENT, "d3d_resize_bilinear: NULL pointer");
    if (maps < 1 || in_height < 1 || in_width < 1 || out_height < 1 || out_width < 1 ||
        (long long)out_height * out_width > INT32_MAX || (long long)in_height * in_width > INT32_MAX)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_resize_bilinear: bad extent N=%d %dx%d -> %dx%d", maps, in_height, in_width,
                    out_height, out_width);
    const int hw = out_height * out_width;
    resize_bilinear_kernel<<<(hw + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(
        in, out, maps, in_height, in_width, out_height, out_width, (float)in_height / (float)out_height,
        (float)in_width / (float)out_width);
    count_launch();
    return check_launch("resize_bilinear_kernel");
}

extern "C" int d3d_nchw_to_nhwc(const float* in, float* out, int32_t channels, int32_t height, int32_t width,
                                void* cuda_stream) {
    if (!in || !out) return fail(D3D_ERR_BAD_ARGUMENT, "d3d_nchw_to_nhwc: NULL pointer");
    if (channels <= 0 || height <= 0 || width <= 0)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_nchw_to_nhwc: non-positive extent C=%d H=%d W=%d", channels, height,
                    width);
    long long hw = (long long)height * width;
    if (hw > INT32_MAX) return fail(D3D_ERR_UNSUPPORTED, "d3d_nchw_to_nhwc: H*W exceeds 2^31-1");
    if (channels == 32 || channels == 16 || channels == 8) {       // the reference's feature widths: warp-tile version
        const unsigned blocks = (unsigned)((hw + 255) / 256);
        cudaStream_t st = (cudaStream_t)cuda_stream;
        if (channels == 32) nchw_to_nhwc_tile_kernel<32><<<blocks, 256, 0, st>>>(in, out, (int)hw);
        else if (channels == 16) nchw_to_nhwc_tile_kernel<16><<<blocks, 256, 0, st>>>(in, out, (int)hw);
        else nchw_to_nhwc_tile_kernel<8><<<blocks, 256, 0, st>>>(in, out, (int)hw);
        count_launch();
        return check_launch("nchw_to_nhwc_tile_kernel");
    }
    if (channels % 8 == 0 && hw % 4 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        const dim3 grid((unsigned)((hw / 4 + 255) / 256), (unsigned)(channels / 8));
        nchw_to_nhwc_vec_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(in, out, channels, (int)hw);
        count_launch();
        return check_launch("nchw_to_nhwc_vec_kernel");
    }
    size_t smem = (size_t)channels * (kTilePix + 1) * sizeof(float);
    if (smem > 48 * 1024) return fail(D3D_ERR_UNSUPPORTED, "d3d_nchw_to_nhwc: C=%d too large", channels);
    unsigned blocks = (unsigned)((hw + kTilePix - 1) / kTilePix);
    nchw_to_nhwc_kernel<<<blocks, 256, smem, (cudaStream_t)cuda_stream>>>(in, out, channels, (int)hw);
    count_launch();
    return check_launch("nchw_to_nhwc_kernel");
}

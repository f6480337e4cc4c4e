// Depth-map fusion: geometric consistency of one reference view against all its source views, one launch.
//
// Stands in for ConsistencyChecker.check_cupy (fuse/consistency_check_n.py:29-138: ~25 CuPy launches, a host-side
// meshgrid and 9 host->device + 5 device->host copies per SOURCE view) and for the accumulation loop around it
// (fuse/fusion_3d_normal.py:436-536).  One thread per reference pixel, consecutive lanes = consecutive x: the
// reference maps are read once, coalesced; neighbouring pixels project to neighbouring source pixels, so the
// gathers of a warp fall into a handful of 128-byte lines.  Points are fp64 and normals fp32 -- the types the
// reference's own promotion rules produce (int64 grid x fp32 depth -> fp64; fp32 rotation x fp32 normal -> fp32).
// Bound: HBM (20 B per reference pixel + 16 B gathered and <= 33 B written per pixel and source view).
#include <cmath>

#include "common.cuh"

namespace d3d {

constexpr int kFuseThreads = 256;
constexpr int kGeom = D3D_FUSE_GEOM_DOUBLES;

struct FuseParams {
    const float* __restrict__ depth_ref;
    const float* __restrict__ normal_ref;
    const float* __restrict__ prob_ref;
    const double* __restrict__ geometry;
    const float* depth_src[D3D_FUSE_MAX_SRC];
    const float* normal_src[D3D_FUSE_MAX_SRC];
    float* depth_src_out[D3D_FUSE_MAX_SRC];
    uint8_t* __restrict__ mask;
    float* __restrict__ depth_reprojected;
    float* __restrict__ xyz_world_src;
    float* __restrict__ angle_conf;
    int32_t* __restrict__ consistent_count;
    float* __restrict__ xyz_fused;
    uint8_t* __restrict__ final_mask;
    float* __restrict__ depth_ref_filtered;
    float* __restrict__ accum;
    int accumulate;
    double dist2_limit;               // see dist2_limit_of
    float depth_threshold, confidence_threshold, normal_threshold_cos;
    int S, H, W, Hs, Ws, min_consistent;
};

// Geometry block (D3D_FUSE_GEOM_DOUBLES = 64 doubles; rows padded to 4 so every row is two 16-byte loads):
//   [0..11] A 3x4   [12..23] B 3x4   [24..35] C 3x4   [36..51] D 4x4   [52..63] R 3x4 (staged as fp32)
constexpr int kA = 0, kB = 12, kC = 24, kD = 36, kR = 52;

__device__ __forceinline__ double row3(const double* r, double a, double b, double c) {
    const double2 u = *reinterpret_cast<const double2*>(r), v = *reinterpret_cast<const double2*>(r + 2);
    return u.x * a + u.y * b + v.x * c;                    // fma(c2, c, fma(c1, b, c0 * a)): dgemm's order
}
__device__ __forceinline__ double row4(const double* r, double a, double b, double c, double w) {
    const double2 u = *reinterpret_cast<const double2*>(r), v = *reinterpret_cast<const double2*>(r + 2);
    return u.x * a + u.y * b + v.x * c + v.y * w;
}
// fp32 rotation of an fp32 normal (np.matmul of two float32 arrays stays float32)
__device__ __forceinline__ void rot3f(const float* m, float a, float b, float c, float& x, float& y, float& z) {
    const float4 r0 = *reinterpret_cast<const float4*>(m), r1 = *reinterpret_cast<const float4*>(m + 4),
                 r2 = *reinterpret_cast<const float4*>(m + 8);
    x = r0.x * a + r0.y * b + r0.z * c;
    y = r1.x * a + r1.y * b + r1.z * c;
    z = r2.x * a + r2.y * b + r2.z * c;
}
// a1/b and a2/b, correctly rounded, from ONE reciprocal: q0 = a*r is within an ulp of a/b and the residual
// correction fma(fma(-b, q0, a), r, q0) rounds it correctly (Markstein), as the tail of IEEE division does
__device__ __forceinline__ void div_pair(double a1, double a2, double b, double& q1, double& q2) {
    const double r = __drcp_rn(b);
    double t = a1 * r;
    q1 = fma(fma(-b, t, a1), r, t);
    t = a2 * r;
    q2 = fma(fma(-b, t, a2), r, t);
}
// Python-style modulo: CuPy wraps out-of-bounds integer-array indices around the axis
__device__ __forceinline__ long long wrap(long long i, int n) {
    long long r = i % n;
    return r < 0 ? r + n : r;
}

// kPerSource: the per-source maps of ConsistencyChecker.check are wanted (depth_reprojected, xyz_world_src,
// angle_conf); kFused: the accumulated outputs of fuse_depths are wanted.  Compile-time so the common fused call
// carries no dead stores or pointer tests in its loop (the kernel is issue-bound).
template <bool kPerSource, bool kFused>
__global__ void __launch_bounds__(kFuseThreads) consistency_fuse_kernel(const FuseParams p) {
    extern __shared__ double geo[];                        // (1 + S) blocks of kGeom doubles, then the fp32 rotations
    float* rotf = reinterpret_cast<float*>(geo + (1 + p.S) * kGeom);
    for (int i = threadIdx.x; i < (1 + p.S) * kGeom; i += kFuseThreads) {
        const double v = __ldg(p.geometry + i);
        geo[i] = v;
        const int blk = i / kGeom, k = i % kGeom;
        if (k >= kR) rotf[blk * 12 + (k - kR)] = (float)v;
    }
    __syncthreads();

    const long long hw = (long long)p.H * p.W;
    const long long pix = (long long)blockIdx.x * kFuseThreads + threadIdx.x;
    if (pix >= hw) return;
    const int y = (int)(pix / p.W), x = (int)(pix - (long long)y * p.W);

    const float d = __ldg(p.depth_ref + pix);
    const float prob = __ldg(p.prob_ref + pix);
    const float n0 = __ldg(p.normal_ref + 3 * pix), n1 = __ldg(p.normal_ref + 3 * pix + 1),
                n2 = __ldg(p.normal_ref + 3 * pix + 2);
    const double dd = (double)d, xd = (double)x, yd = (double)y;
    const double* G0 = geo;
    // reference camera space (:51-53): Kinv_ref (x d, y d, d)
    const double vx = xd * dd, vy = yd * dd;
    const double px = row3(G0 + kA, vx, vy, dd), py = row3(G0 + kA + 4, vx, vy, dd), pz = row3(G0 + kA + 8, vx, vy, dd);
    float nrx, nry, nrz;                                   // reference normal in the world (:104-106)
    rot3f(rotf, n0, n1, n2, nrx, nry, nrz);
    // numpy rounds every product before it adds (np.sum(a*b), np.linalg.norm): no FMA contraction on these
    auto dot3 = [](float a0, float a1, float a2, float b0, float b1, float b2) {
        return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
    };
    const float nr_norm = sqrtf(dot3(nrx, nry, nrz, nrx, nry, nrz));

    // accumulators of fusion_3d_normal.py:449-455, 525-527: world point of the reference pixel, confidence 1
    float ax = (float)row4(G0 + kD, px, py, pz, 1.0), ay = (float)row4(G0 + kD + 4, px, py, pz, 1.0),
          az = (float)row4(G0 + kD + 8, px, py, pz, 1.0);
    float aconf = 1.f;
    int count = 1;
    if (kFused && p.accumulate) {                          // continue an earlier call over other source views
        ax = p.accum[pix]; ay = p.accum[hw + pix]; az = p.accum[2 * hw + pix]; aconf = p.accum[3 * hw + pix];
        count = p.consistent_count[pix];
    }
    const bool gate = prob > p.confidence_threshold && d > 0.f;

    for (int s = 0; s < p.S; ++s) {
        const double* G = geo + (1 + s) * kGeom;
        // source camera space and pixel (:56-72)
        const double qx = row4(G + kA, px, py, pz, 1.0), qy = row4(G + kA + 4, px, py, pz, 1.0),
                     qz = row4(G + kA + 8, px, py, pz, 1.0);
        const double kx = row3(G + kB, qx, qy, qz), ky = row3(G + kB + 4, qx, qy, qz), kz = row3(G + kB + 8, qx, qy, qz);
        double fxs, fys;
        div_pair(kx, ky, kz, fxs, fys);
        fxs += 0.5;
        fys += 0.5;
        // `at`: where the source maps are read; `at_used`: the source pixel a consistent reference pixel consumes,
        // (x_src[mask] + 0.5).astype(int) ON THE INTEGER x_src (:123-124) = x_src, or x_src + 1 where it is negative
        // (truncation towards zero), wrapped by the same index rule.  They differ only for wrapped (negative) coordinates.
        long long at, at_used;
        double xsd, ysd;
        if (fabs(fxs) < 2147483000.0 && fabs(fys) < 2147483000.0) {      // (NaN fails the test)
            int xi = __double2int_rz(fxs), yi = __double2int_rz(fys);
            xsd = (double)xi;
            ysd = (double)yi;
            if ((unsigned)xi < (unsigned)p.Ws && (unsigned)yi < (unsigned)p.Hs) {
                at = at_used = (long long)yi * p.Ws + xi;
            } else {                                                     // CuPy's wrap-around (32-bit: |xi|, |yi| < 2^31 - 648)
                auto wrap32 = [](int i, int n) { const int r = i % n; return r < 0 ? r + n : r; };
                const int xw = wrap32(xi, p.Ws), yw = wrap32(yi, p.Hs);
                at = (long long)yw * p.Ws + xw;
                // the consumed pixel differs only for negative coordinates: one step towards zero, which the wrapped
                // index follows (+1, and back to 0 past the last column / row)
                const int xu = xi < 0 ? (xw + 1 == p.Ws ? 0 : xw + 1) : xw;
                const int yu = yi < 0 ? (yw + 1 == p.Hs ? 0 : yw + 1) : yw;
                at_used = (long long)yu * p.Ws + xu;
            }
        } else {                                                         // astype(int) is int64 upstream
            const long long xs = __double2ll_rz(fxs), ys = __double2ll_rz(fys);
            xsd = (double)xs;
            ysd = (double)ys;
            at = wrap(ys, p.Hs) * p.Ws + wrap(xs, p.Ws);
            at_used = wrap(ys + (ys < 0), p.Hs) * p.Ws + wrap(xs + (xs < 0), p.Ws);
        }
        const float sd = __ldg(p.depth_src[s] + at);
        const float* np_ = p.normal_src[s] + 3 * at;
        const float m0 = __ldg(np_), m1 = __ldg(np_ + 1), m2 = __ldg(np_ + 2);
        // back to the source camera, the world, the reference camera (:76-92)
        const double sdd = (double)sd;
        const double bx = xsd * sdd, by = ysd * sdd;
        const double cx = row3(G + kC, bx, by, sdd), cy = row3(G + kC + 4, bx, by, sdd), cz = row3(G + kC + 8, bx, by, sdd);
        const double wx = row4(G + kD, cx, cy, cz, 1.0), wy = row4(G + kD + 4, cx, cy, cz, 1.0),
                     wz = row4(G + kD + 8, cx, cy, cz, 1.0), ww = row4(G + kD + 12, cx, cy, cz, 1.0);
        const double rx = row4(G0 + kB, wx, wy, wz, ww), ry = row4(G0 + kB + 4, wx, wy, wz, ww),
                     rz = row4(G0 + kB + 8, wx, wy, wz, ww);
        const float depth_rep = (float)rz;
        const double ux = row3(G0 + kC, rx, ry, rz), uy = row3(G0 + kC + 4, rx, ry, rz), uz = row3(G0 + kC + 8, rx, ry, rz);
        double xrd, yrd;
        div_pair(ux, uy, uz, xrd, yrd);
        const float xr = (float)xrd, yr = (float)yrd;
        // position (fp32 pixel minus int64 grid -> fp64; sqrt(s) < t restated as s < dist2_limit, exactly: abi),
        // depth (fp32), normal (fp32) tests (:95-123)
        const double ex = (double)xr - xd, ey = (double)yr - yd;
        const double dist2 = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
        const float rel = __fdiv_rn(fabsf(depth_rep - d), d);
        float nsx, nsy, nsz;
        rot3f(rotf + (1 + s) * 12, m0, m1, m2, nsx, nsy, nsz);
        const float cosv = __fdiv_rn(dot3(nrx, nry, nrz, nsx, nsy, nsz),
                                     __fmul_rn(nr_norm, sqrtf(dot3(nsx, nsy, nsz, nsx, nsy, nsz))));
        const bool ok = gate && dist2 < p.dist2_limit && rel < p.depth_threshold && cosv > p.normal_threshold_cos;

        const float conf = ok ? fmaxf(cosv, 0.f) : 0.f;
        const float fx = ok ? (float)wx : 0.f, fy = ok ? (float)wy : 0.f, fz = ok ? (float)wz : 0.f;
        if (ok) {
            ++count;
            ax = __fadd_rn(ax, __fmul_rn(conf, fx));                // (angle_conf * xyz).astype(float32), :526
            ay = __fadd_rn(ay, __fmul_rn(conf, fy));
            az = __fadd_rn(az, __fmul_rn(conf, fz));
            aconf += conf;
            if (p.depth_src_out[s]) p.depth_src_out[s][at_used] = 0.f;   // consumed by this reference view (:123-126)
        }
        const long long o = (long long)s * hw + pix;
        if (p.mask) p.mask[o] = ok;
        if (kPerSource) {
            if (p.depth_reprojected) p.depth_reprojected[o] = ok ? depth_rep : 0.f;
            if (p.angle_conf) p.angle_conf[o] = conf;
            if (p.xyz_world_src) {
                float* q = p.xyz_world_src + (long long)s * 3 * hw + pix;
                q[0] = fx; q[hw] = fy; q[2 * hw] = fz;
            }
        }
    }
    if (!kFused) return;
    if (p.accum) {
        p.accum[pix] = ax; p.accum[hw + pix] = ay; p.accum[2 * hw + pix] = az; p.accum[3 * hw + pix] = aconf;
    }
    const bool keep = count >= p.min_consistent;
    if (p.consistent_count) p.consistent_count[pix] = count;
    if (p.final_mask) p.final_mask[pix] = keep;
    if (p.depth_ref_filtered) p.depth_ref_filtered[pix] = keep ? d : 0.f;
    if (p.xyz_fused) {
        p.xyz_fused[pix] = __fdiv_rn(ax, aconf);
        p.xyz_fused[hw + pix] = __fdiv_rn(ay, aconf);
        p.xyz_fused[2 * hw + pix] = __fdiv_rn(az, aconf);
    }
}

// smallest s with sqrt(s) >= t: `sqrt(s) < t` (consistency_check_n.py:95, 116) is then exactly `s < limit`
// (IEEE sqrt is correctly rounded and monotonic, on the host as on the device)
static double dist2_limit_of(double t) {
    if (!(t > 0.0)) return 0.0;                            // nothing is closer than a non-positive threshold
    if (std::isinf(t)) return t;
    double s = t * t;
    while (s > 0.0 && std::sqrt(s) >= t) s = std::nextafter(s, 0.0);
    while (std::sqrt(std::nextafter(s, INFINITY)) < t) s = std::nextafter(s, INFINITY);
    return std::sqrt(s) < t ? std::nextafter(s, INFINITY) : s;
}

}  // namespace d3d

using namespace d3d;

extern "C" int d3d_consistency_fuse(const D3dFuseArgs* a, void* cuda_stream) {
    if (!a) return fail(D3D_ERR_BAD_ARGUMENT, "d3d_consistency_fuse: args is NULL");
    if (a->struct_size != sizeof(D3dFuseArgs))
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_consistency_fuse: struct_size %u != %zu", a->struct_size, sizeof(D3dFuseArgs));
    if (a->num_src < 1 || a->num_src > D3D_FUSE_MAX_SRC)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_consistency_fuse: num_src %d outside 1..%d", a->num_src, D3D_FUSE_MAX_SRC);
    if (a->height < 1 || a->width < 1 || a->src_height < 1 || a->src_width < 1)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_consistency_fuse: bad extent %dx%d / %dx%d", a->height, a->width,
                    a->src_height, a->src_width);
    if (!a->depth_ref || !a->normal_ref || !a->prob_ref || !a->geometry)
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_consistency_fuse: depth_ref/normal_ref/prob_ref/geometry is NULL");
    for (int s = 0; s < a->num_src; ++s) {
        if (!a->depth_src[s] || !a->normal_src[s])
            return fail(D3D_ERR_BAD_ARGUMENT, "d3d_consistency_fuse: depth_src[%d]/normal_src[%d] is NULL", s, s);
        if (a->depth_src_out[s] && a->depth_src_out[s] == a->depth_src[s])
            return fail(D3D_ERR_BAD_ARGUMENT, "d3d_consistency_fuse: depth_src_out[%d] aliases depth_src[%d] (the gathers "
                        "of one pixel would race with the zeroing of another)", s, s);
    }
    if (a->accumulate && (!a->accum || !a->consistent_count))
        return fail(D3D_ERR_BAD_ARGUMENT, "d3d_consistency_fuse: accumulate needs accum and consistent_count");
    const long long hw = (long long)a->height * a->width;
    if (hw > (1LL << 31) - kFuseThreads) return fail(D3D_ERR_UNSUPPORTED, "d3d_consistency_fuse: H*W too large");

    cudaStream_t stream = (cudaStream_t)cuda_stream;
    FuseParams p;
    p.depth_ref = a->depth_ref; p.normal_ref = a->normal_ref; p.prob_ref = a->prob_ref; p.geometry = a->geometry;
    const size_t src_bytes = (size_t)a->src_height * a->src_width * sizeof(float);
    for (int s = 0; s < D3D_FUSE_MAX_SRC; ++s) {
        const bool on = s < a->num_src;
        p.depth_src[s] = on ? a->depth_src[s] : nullptr;
        p.normal_src[s] = on ? a->normal_src[s] : nullptr;
        p.depth_src_out[s] = on ? a->depth_src_out[s] : nullptr;
        if (on && a->depth_src_out[s]) {
            cudaError_t e = cudaMemcpyAsync(a->depth_src_out[s], a->depth_src[s], src_bytes, cudaMemcpyDeviceToDevice, stream);
            if (e != cudaSuccess) return fail(D3D_ERR_CUDA, "d3d_consistency_fuse: copy of depth_src[%d]: %s", s, cudaGetErrorString(e));
        }
    }
    p.mask = a->mask; p.depth_reprojected = a->depth_reprojected; p.xyz_world_src = a->xyz_world_src;
    p.angle_conf = a->angle_conf; p.consistent_count = a->consistent_count; p.xyz_fused = a->xyz_fused;
    p.final_mask = a->final_mask; p.depth_ref_filtered = a->depth_ref_filtered;
    p.accum = a->accum; p.accumulate = a->accumulate != 0;
    p.dist2_limit = dist2_limit_of(a->position_threshold);
    p.depth_threshold = a->depth_threshold; p.confidence_threshold = a->confidence_threshold;
    p.normal_threshold_cos = a->normal_threshold_cos;
    p.S = a->num_src; p.H = a->height; p.W = a->width; p.Hs = a->src_height; p.Ws = a->src_width;
    p.min_consistent = a->min_consistent;

    const unsigned blocks = (unsigned)((hw + kFuseThreads - 1) / kFuseThreads);
    const size_t smem = (size_t)(1 + a->num_src) * (kGeom * sizeof(double) + 12 * sizeof(float));
    const bool per_source = p.depth_reprojected || p.angle_conf || p.xyz_world_src;
    const bool fused = p.consistent_count || p.final_mask || p.depth_ref_filtered || p.xyz_fused || p.accum;
    if (per_source && fused) consistency_fuse_kernel<true, true><<<blocks, kFuseThreads, smem, stream>>>(p);
    else if (per_source) consistency_fuse_kernel<true, false><<<blocks, kFuseThreads, smem, stream>>>(p);
    else consistency_fuse_kernel<false, true><<<blocks, kFuseThreads, smem, stream>>>(p);
    count_launch();
    return check_launch("consistency_fuse_kernel");
}

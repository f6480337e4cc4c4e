// Shared device helpers of the sweep kernels: packed-float shorthands, the floor-by-magic-constant constants, the shared
// division (one reciprocal, one Newton step, the residual correction IEEE division ends with) and the scalar projection
// chain in the reference's fp32 operation order (sweep_direct.cuh runs it as is; sweep_quad.cuh / sweep_lean.cuh restate
// it packed / with their own key).  (Round 1's first production kernel, sweep_fast.cuh, lived here; it was superseded by
// sweep_lean.cuh and then sweep_quad.cuh and is retired -- git history has it.)
#pragma once
#include "common.cuh"
#include "sweep_refetch.cuh"

namespace d3d {

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 splat(float a) { return make_float2(a, a); }

__device__ __forceinline__ float2 fadd2_rd(float2 a, float2 b) {       // packed round-down add (FADD2.RM)
    float2 r;
    asm("add.rm.f32x2 %0, %1, %2;"
        : "=l"(*reinterpret_cast<unsigned long long*>(&r))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return r;
}
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

// key = [27:14] y0+4 | [13:0] x0+4  (floor corner of the 2x2 footprint)
constexpr float kMagic = 12582912.f;           // 1.5 * 2^23: float(kMagic + n) has bits 0x4B400000 + n
constexpr int kMagicBits = 0x4B400000;

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

// The same chain for one pixel and view at TWO depths, packed (what sweep_quad.cuh's projecting lanes run; restated here for
// sweep_acc.cuh): every packed operation is the reference's IEEE operation on each half.  The floor is exact without the
// step-down: a round-DOWN add against 1.5*2^23 leaves floor() in the low mantissa bits.  Entries: (fx, fy, fx*fy, key) with
// key = [31:16] y0 | [15:0] x0 as 16-bit two's complement (coordinates are clamped to [-2, size+1] first).
template <bool kIeeeDiv>
__device__ __forceinline__ void project_pair(float rx, float ry, float rz, float tx, float ty, float tz, float2 d,
                                             const SweepParams& p, float4& ea, float4& eb) {
    // (nvcc contracts __fmul2_rn + __fadd2_rn into one FFMA2, which would round once where the reference rounds twice:
    // every add that follows a multiply is a scalar __fadd_rn)
    const float2 Xm = __fmul2_rn(splat(rx), d), Ym = __fmul2_rn(splat(ry), d), Zm = __fmul2_rn(splat(rz), d);
    const float2 X = f2(__fadd_rn(Xm.x, tx), __fadd_rn(Xm.y, tx));
    const float2 Y = f2(__fadd_rn(Ym.x, ty), __fadd_rn(Ym.y, ty));
    const float2 Z = f2(__fadd_rn(Zm.x, tz), __fadd_rn(Zm.y, tz));
    float2 u, v;
    if (kIeeeDiv) {
        u = f2(__fdiv_rn(X.x, Z.x), __fdiv_rn(X.y, Z.y));
        v = f2(__fdiv_rn(Y.x, Z.x), __fdiv_rn(Y.y, Z.y));
    } else {                                               // div2(), both depths at once
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
    ix = __fmul2_rn(__fmul2_rn(__fadd2_rn(ix, splat(1.f)), splat(0.5f)), splat(p.wm1));
    iy = __fmul2_rn(__fmul2_rn(__fadd2_rn(iy, splat(1.f)), splat(0.5f)), splat(p.hm1));
    const float xhi = p.wm1 + 2.f, yhi = p.hm1 + 2.f;
    ix = f2(fminf(fmaxf(ix.x, -2.f), xhi), fminf(fmaxf(ix.y, -2.f), xhi));   // NaN -> -2: out of bounds
    iy = f2(fminf(fmaxf(iy.x, -2.f), yhi), fminf(fmaxf(iy.y, -2.f), yhi));
    const float2 mx = fadd2_rd(ix, splat(kMagic)), my = fadd2_rd(iy, splat(kMagic));
    const float2 fx = __fadd2_rn(ix, neg2(__fadd2_rn(mx, splat(-kMagic))));
    const float2 fy = __fadd2_rn(iy, neg2(__fadd2_rn(my, splat(-kMagic))));
    const float2 fxy = __fmul2_rn(fx, fy);
    const unsigned ka = __byte_perm(__float_as_uint(mx.x), __float_as_uint(my.x), 0x5410);
    const unsigned kb = __byte_perm(__float_as_uint(mx.y), __float_as_uint(my.y), 0x5410);
    ea = make_float4(fx.x, fy.x, fxy.x, __uint_as_float(ka));
    eb = make_float4(fx.y, fy.y, fxy.y, __uint_as_float(kb));
}

}  // namespace d3d

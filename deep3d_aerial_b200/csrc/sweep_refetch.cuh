// In-place footprint re-fetch for the production sweep kernel (generated text, see the comment below).
#pragma once
#include "common.cuh"

namespace d3d {

// Re-fetch one view's 2x2 footprint (8 channels per corner) when its key moved, and turn the corners
// a,b,c,d (out-of-bounds ones zeroed by the key's mask; their addresses are clamped into the image)
// into the interpolation coefficients A=a, B=b-a, C=c-a, D=a-b-c+d -- all IN PLACE, in one asm block
// with its own branch.  Written in PTX on purpose: any C++-level conditional update of the 128-register
// footprint cache made NVVM/ptxas copy the whole cache on every plane's common path (24-32 MOVs per
// view per plane, measured in SASS); an opaque block with read-write operands keeps it in fixed registers.
//   t[k][j]: corner k (nw, ne, sw, se), channel pair j of this lane.
__device__ __forceinline__ void refetch_footprint(float2 (&t)[4][4], unsigned key, unsigned old_key,
                                                  const float* base, int wm1, int hm1, int row_bytes,
                                                  int texel_bytes) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .s32 x0, y0, xa, xb, ya, yb, t;\n\t"
        ".reg .s64 r0, r1, oa, ob, pa, pb, pc, pd;\n\t"
        ".reg .f32 ma, mb, mc, md;\n\t"
        "setp.eq.u32 p, %32, %33;\n\t"
        "@p bra SAME;\n\t"
        "and.b32 x0, %32, 0x3fff;\n\t"
        "sub.s32 x0, x0, 4;\n\t"
        "shr.u32 y0, %32, 14;\n\t"
        "and.b32 y0, y0, 0x3fff;\n\t"
        "sub.s32 y0, y0, 4;\n\t"
        "max.s32 xa, x0, 0;\n\t"
        "min.s32 xa, xa, %35;\n\t"
        "add.s32 xb, x0, 1;\n\t"
        "max.s32 xb, xb, 0;\n\t"
        "min.s32 xb, xb, %35;\n\t"
        "max.s32 ya, y0, 0;\n\t"
        "min.s32 ya, ya, %36;\n\t"
        "add.s32 yb, y0, 1;\n\t"
        "max.s32 yb, yb, 0;\n\t"
        "min.s32 yb, yb, %36;\n\t"
        "mul.wide.s32 r0, ya, %37;\n\t"
        "mul.wide.s32 r1, yb, %37;\n\t"
        "mul.wide.s32 oa, xa, %38;\n\t"
        "mul.wide.s32 ob, xb, %38;\n\t"
        "add.s64 r0, r0, %34;\n\t"
        "add.s64 r1, r1, %34;\n\t"
        "add.s64 pa, r0, oa;\n\t"
        "add.s64 pb, r0, ob;\n\t"
        "add.s64 pc, r1, oa;\n\t"
        "add.s64 pd, r1, ob;\n\t"
        "ld.global.nc.v4.f32 {%0, %1, %2, %3}, [pa];\n\t"
        "ld.global.nc.v4.f32 {%4, %5, %6, %7}, [pa+16];\n\t"
        "ld.global.nc.v4.f32 {%8, %9, %10, %11}, [pb];\n\t"
        "ld.global.nc.v4.f32 {%12, %13, %14, %15}, [pb+16];\n\t"
        "ld.global.nc.v4.f32 {%16, %17, %18, %19}, [pc];\n\t"
        "ld.global.nc.v4.f32 {%20, %21, %22, %23}, [pc+16];\n\t"
        "ld.global.nc.v4.f32 {%24, %25, %26, %27}, [pd];\n\t"
        "ld.global.nc.v4.f32 {%28, %29, %30, %31}, [pd+16];\n\t"
        "shr.u32 t, %32, 28;\n\t"
        "and.b32 t, t, 1;\n\t"
        "cvt.rn.f32.s32 ma, t;\n\t"
        "shr.u32 t, %32, 29;\n\t"
        "and.b32 t, t, 1;\n\t"
        "cvt.rn.f32.s32 mb, t;\n\t"
        "shr.u32 t, %32, 30;\n\t"
        "and.b32 t, t, 1;\n\t"
        "cvt.rn.f32.s32 mc, t;\n\t"
        "shr.u32 t, %32, 31;\n\t"
        "and.b32 t, t, 1;\n\t"
        "cvt.rn.f32.s32 md, t;\n\t"
        "mul.f32 %0, %0, ma;\n\t"
        "mul.f32 %8, %8, mb;\n\t"
        "mul.f32 %16, %16, mc;\n\t"
        "mul.f32 %24, %24, md;\n\t"
        "sub.f32 %8, %8, %0;\n\t"
        "sub.f32 %24, %24, %16;\n\t"
        "sub.f32 %16, %16, %0;\n\t"
        "sub.f32 %24, %24, %8;\n\t"
        "mul.f32 %1, %1, ma;\n\t"
        "mul.f32 %9, %9, mb;\n\t"
        "mul.f32 %17, %17, mc;\n\t"
        "mul.f32 %25, %25, md;\n\t"
        "sub.f32 %9, %9, %1;\n\t"
        "sub.f32 %25, %25, %17;\n\t"
        "sub.f32 %17, %17, %1;\n\t"
        "sub.f32 %25, %25, %9;\n\t"
        "mul.f32 %2, %2, ma;\n\t"
        "mul.f32 %10, %10, mb;\n\t"
        "mul.f32 %18, %18, mc;\n\t"
        "mul.f32 %26, %26, md;\n\t"
        "sub.f32 %10, %10, %2;\n\t"
        "sub.f32 %26, %26, %18;\n\t"
        "sub.f32 %18, %18, %2;\n\t"
        "sub.f32 %26, %26, %10;\n\t"
        "mul.f32 %3, %3, ma;\n\t"
        "mul.f32 %11, %11, mb;\n\t"
        "mul.f32 %19, %19, mc;\n\t"
        "mul.f32 %27, %27, md;\n\t"
        "sub.f32 %11, %11, %3;\n\t"
        "sub.f32 %27, %27, %19;\n\t"
        "sub.f32 %19, %19, %3;\n\t"
        "sub.f32 %27, %27, %11;\n\t"
        "mul.f32 %4, %4, ma;\n\t"
        "mul.f32 %12, %12, mb;\n\t"
        "mul.f32 %20, %20, mc;\n\t"
        "mul.f32 %28, %28, md;\n\t"
        "sub.f32 %12, %12, %4;\n\t"
        "sub.f32 %28, %28, %20;\n\t"
        "sub.f32 %20, %20, %4;\n\t"
        "sub.f32 %28, %28, %12;\n\t"
        "mul.f32 %5, %5, ma;\n\t"
        "mul.f32 %13, %13, mb;\n\t"
        "mul.f32 %21, %21, mc;\n\t"
        "mul.f32 %29, %29, md;\n\t"
        "sub.f32 %13, %13, %5;\n\t"
        "sub.f32 %29, %29, %21;\n\t"
        "sub.f32 %21, %21, %5;\n\t"
        "sub.f32 %29, %29, %13;\n\t"
        "mul.f32 %6, %6, ma;\n\t"
        "mul.f32 %14, %14, mb;\n\t"
        "mul.f32 %22, %22, mc;\n\t"
        "mul.f32 %30, %30, md;\n\t"
        "sub.f32 %14, %14, %6;\n\t"
        "sub.f32 %30, %30, %22;\n\t"
        "sub.f32 %22, %22, %6;\n\t"
        "sub.f32 %30, %30, %14;\n\t"
        "mul.f32 %7, %7, ma;\n\t"
        "mul.f32 %15, %15, mb;\n\t"
        "mul.f32 %23, %23, mc;\n\t"
        "mul.f32 %31, %31, md;\n\t"
        "sub.f32 %15, %15, %7;\n\t"
        "sub.f32 %31, %31, %23;\n\t"
        "sub.f32 %23, %23, %7;\n\t"
        "sub.f32 %31, %31, %15;\n\t"
        "SAME:\n\t"
        "}"
        : "+f"(t[0][0].x), "+f"(t[0][0].y), "+f"(t[0][1].x), "+f"(t[0][1].y),
          "+f"(t[0][2].x), "+f"(t[0][2].y), "+f"(t[0][3].x), "+f"(t[0][3].y),
          "+f"(t[1][0].x), "+f"(t[1][0].y), "+f"(t[1][1].x), "+f"(t[1][1].y),
          "+f"(t[1][2].x), "+f"(t[1][2].y), "+f"(t[1][3].x), "+f"(t[1][3].y),
          "+f"(t[2][0].x), "+f"(t[2][0].y), "+f"(t[2][1].x), "+f"(t[2][1].y),
          "+f"(t[2][2].x), "+f"(t[2][2].y), "+f"(t[2][3].x), "+f"(t[2][3].y),
          "+f"(t[3][0].x), "+f"(t[3][0].y), "+f"(t[3][1].x), "+f"(t[3][1].y),
          "+f"(t[3][2].x), "+f"(t[3][2].y), "+f"(t[3][3].x), "+f"(t[3][3].y)
        : "r"(key), "r"(old_key), "l"(base), "r"(wm1), "r"(hm1), "r"(row_bytes), "r"(texel_bytes));
}


}  // namespace d3d

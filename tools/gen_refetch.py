"""Generates deep3d_aerial_b200/csrc/sweep_refetch.cuh: the in-place PTX footprint re-fetch blocks of the
production sweep kernel, one per channels-per-lane width (4 and 8).

    python tools/gen_refetch.py

The block is written in PTX because every C++-level conditional update of the register-resident footprint
cache made NVVM/ptxas copy the cache on each plane's common path (24-32 MOVs per view per plane, seen in
SASS); an opaque asm block with read-write operands and its own branch keeps the cache in fixed registers.
"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "deep3d_aerial_b200", "csrc", "sweep_refetch.cuh")


def block(cpt):
    n = 4 * cpt                       # float operands: corner k, channel c -> index k*cpt + c
    key, old, base, wm1, hm1, rowb, texb = (n + i for i in range(7))
    L = []
    A = L.append
    A("{")
    A(".reg .pred p, q, px0, px1, py0, py1;")
    A(".reg .s32 x0, y0, xa, xb, ya, yb, t;")
    A(".reg .s64 r0, r1, oa, ob, pa, pb, pc, pd;")
    A(".reg .f32 ma, mb, mc, md;")
    A("setp.eq.u32 p, %%%d, %%%d;" % (key, old))
    A("@p bra SAME;")
    A("and.b32 x0, %%%d, 0x3fff;" % key)
    A("sub.s32 x0, x0, 4;")
    A("shr.u32 y0, %%%d, 14;" % key)
    A("and.b32 y0, y0, 0x3fff;")
    A("sub.s32 y0, y0, 4;")
    A("max.s32 xa, x0, 0;")
    A("min.s32 xa, xa, %%%d;" % wm1)
    A("add.s32 xb, x0, 1;")
    A("max.s32 xb, xb, 0;")
    A("min.s32 xb, xb, %%%d;" % wm1)
    A("max.s32 ya, y0, 0;")
    A("min.s32 ya, ya, %%%d;" % hm1)
    A("add.s32 yb, y0, 1;")
    A("max.s32 yb, yb, 0;")
    A("min.s32 yb, yb, %%%d;" % hm1)
    A("mul.wide.s32 r0, ya, %%%d;" % rowb)
    A("mul.wide.s32 r1, yb, %%%d;" % rowb)
    A("mul.wide.s32 oa, xa, %%%d;" % texb)
    A("mul.wide.s32 ob, xb, %%%d;" % texb)
    A("add.s64 r0, r0, %%%d;" % base)
    A("add.s64 r1, r1, %%%d;" % base)
    A("add.s64 pa, r0, oa;")
    A("add.s64 pb, r0, ob;")
    A("add.s64 pc, r1, oa;")
    A("add.s64 pd, r1, ob;")
    for k, ptr in enumerate(["pa", "pb", "pc", "pd"]):
        for c in range(0, cpt, 4):
            b = k * cpt + c
            off = "" if c == 0 else "+%d" % (4 * c)
            A("ld.global.nc.v4.f32 {%%%d, %%%d, %%%d, %%%d}, [%s%s];" % (b, b + 1, b + 2, b + 3, ptr, off))
    # in-bounds masks of the four corners (zeros padding): x0 == xa <=> 0 <= x0 <= W-1, etc.
    A("setp.eq.s32 px0, x0, xa;")
    A("add.s32 t, x0, 1;")
    A("setp.eq.s32 px1, t, xb;")
    A("setp.eq.s32 py0, y0, ya;")
    A("add.s32 t, y0, 1;")
    A("setp.eq.s32 py1, t, yb;")
    A("and.pred q, px0, py0;")
    A("selp.f32 ma, 0f3F800000, 0f00000000, q;")
    A("and.pred q, px1, py0;")
    A("selp.f32 mb, 0f3F800000, 0f00000000, q;")
    A("and.pred q, px0, py1;")
    A("selp.f32 mc, 0f3F800000, 0f00000000, q;")
    A("and.pred q, px1, py1;")
    A("selp.f32 md, 0f3F800000, 0f00000000, q;")
    for c in range(cpt):
        a, b, cc, d = c, cpt + c, 2 * cpt + c, 3 * cpt + c
        A("mul.f32 %%%d, %%%d, ma;" % (a, a))
        A("mul.f32 %%%d, %%%d, mb;" % (b, b))
        A("mul.f32 %%%d, %%%d, mc;" % (cc, cc))
        A("mul.f32 %%%d, %%%d, md;" % (d, d))
        A("sub.f32 %%%d, %%%d, %%%d;" % (b, b, a))
        A("sub.f32 %%%d, %%%d, %%%d;" % (d, d, cc))
        A("sub.f32 %%%d, %%%d, %%%d;" % (cc, cc, a))
        A("sub.f32 %%%d, %%%d, %%%d;" % (d, d, b))
    A("SAME:")
    A("}")
    body = "\n        ".join('"%s\\n\\t"' % x for x in L)
    ops = []
    for k in range(4):
        for j in range(cpt // 2):
            ops.append('"+f"(t[%d][%d].x)' % (k, j))
            ops.append('"+f"(t[%d][%d].y)' % (k, j))
    outs = ",\n          ".join(", ".join(ops[i:i + 4]) for i in range(0, len(ops), 4))
    return '''__device__ __forceinline__ void refetch_footprint(float2 (&t)[4][%d], unsigned key, unsigned old_key,
                                                  const float* base, int wm1, int hm1, int row_bytes,
                                                  int texel_bytes) {
    asm volatile(
        %s
        : %s
        : "r"(key), "r"(old_key), "l"(base), "r"(wm1), "r"(hm1), "r"(row_bytes), "r"(texel_bytes));
}
''' % (cpt // 2, body, outs)


def tex_operands(cpt):
    ops = []
    for k in range(4):
        for j in range(cpt // 2):
            ops.append('"+f"(t[%d][%d].x)' % (k, j))
            ops.append('"+f"(t[%d][%d].y)' % (k, j))
    return ops


def rebuild_lines(A, cpt, ref0=None):
    # A = a, B = b - a, C = c - a, D = (d - c) - (b - a), packed over channel pairs; with ref0 (operand index of the
    # lane's reference texel) A = a - ref: the variance volume is shift invariant, so the reference view drops out
    for c in range(0, cpt, 2):
        a, b, cc, d = c, cpt + c, 2 * cpt + c, 3 * cpt + c
        for reg, i in (("u0", a), ("u1", b), ("u2", cc), ("u3", d)):
            A("mov.b64 %s, {%%%d, %%%d};" % (reg, i, i + 1))
        A("sub.rn.f32x2 u1, u1, u0;")
        A("sub.rn.f32x2 u3, u3, u2;")
        A("sub.rn.f32x2 u2, u2, u0;")
        A("sub.rn.f32x2 u3, u3, u1;")
        for reg, i in (("u1", b), ("u2", cc), ("u3", d)):
            A("mov.b64 {%%%d, %%%d}, %s;" % (i, i + 1, reg))
        if ref0 is not None:
            A("mov.b64 u1, {%%%d, %%%d};" % (ref0 + c, ref0 + c + 1))
            A("sub.rn.f32x2 u0, u0, u1;")
            A("mov.b64 {%%%d, %%%d}, u0;" % (a, a + 1))


def block_off(cpt, split=False):
    """Offset-keyed re-fetch: the projection producer already turned the footprint into a byte offset, so the
    common (interior) case is two address adds, four loads and the packed rebuild."""
    n = 4 * cpt                       # float operands: corner k, channel c -> index k*cpt + c
    if split:
        old, mv, key, base, rowb, hwc4, wid, hei, vplus1, texb, texb16 = (n + i for i in range(11))
    else:
        old, key, base, rowb, hwc4, wid, hei, vplus1, texb, texb16 = (n + i for i in range(10))
    L = []
    A = L.append
    A("{")
    A(".reg .pred p, q, px0, px1, py0, py1;")
    A(".reg .b32 x0, y0, x1, y1, t, vb, o00, o01, o10, o11;")
    A(".reg .b64 w, pa, pb, pc, pd, u0, u1, u2, u3;")
    A("setp.eq.u32 p, %%%d, %%%d;" % (key, old))
    A("@p bra SAME;")
    A("mov.u32 %%%d, %%%d;" % (old, key))
    if split:
        A("shl.b32 t, 1, %%%d;" % vplus1)
        A("or.b32 %%%d, %%%d, t;" % (mv, mv))
    A("and.b32 t, %%%d, 1;" % key)
    A("setp.ne.u32 q, t, 0;")
    A("@q bra SPECIAL;")
    # interior: key = byte offset (view base included) of the north-west texel; all four corners valid
    A("cvt.u64.u32 w, %%%d;" % key)
    A("add.s64 pa, %%%d, w;" % base)
    A("cvt.u64.u32 w, %%%d;" % rowb)
    A("add.s64 pc, pa, w;")
    for k, (ptr, off) in enumerate([("pa", 0), ("pa", 1), ("pc", 0), ("pc", 1)]):
        for c in range(0, cpt, 4):
            b = k * cpt + c
            assert c in (0, 4)
            if off:
                o = "+%%%d" % (texb16 if c else texb)
            else:
                o = "+16" if c else ""
            A("ld.global.nc.v4.f32 {%%%d, %%%d, %%%d, %%%d}, [%s%s];" % (b, b + 1, b + 2, b + 3, ptr, o))
    A("bra SAME;" if split else "bra REBUILD;")
    A("SPECIAL:")
    # border / out of bounds: key = 1 | (x0+4) << 4 | (y0+4) << 18 (key == 1: no corner inside the image);
    # zeros padding = corners outside the image stay 0 (loads predicated off, so no address is clamped)
    A("shr.u32 x0, %%%d, 4;" % key)
    A("and.b32 x0, x0, 0x3fff;")
    A("sub.s32 x0, x0, 4;")
    A("shr.u32 y0, %%%d, 18;" % key)
    A("sub.s32 y0, y0, 4;")
    A("add.s32 x1, x0, 1;")
    A("add.s32 y1, y0, 1;")
    A("setp.lt.u32 px0, x0, %%%d;" % wid)
    A("setp.lt.u32 px1, x1, %%%d;" % wid)
    A("setp.lt.u32 py0, y0, %%%d;" % hei)
    A("setp.lt.u32 py1, y1, %%%d;" % hei)
    A("mul.lo.u32 vb, %%%d, %%%d;" % (hwc4, vplus1))
    A("mad.lo.s32 t, y0, %%%d, x0;" % wid)
    A("mad.lo.s32 o00, t, %%%d, vb;" % texb)
    A("add.s32 o01, o00, %%%d;" % texb)
    A("add.s32 o10, o00, %%%d;" % rowb)
    A("add.s32 o11, o10, %%%d;" % texb)
    for o, ptr in (("o00", "pa"), ("o01", "pb"), ("o10", "pc"), ("o11", "pd")):
        A("cvt.u64.u32 w, %s;" % o)
        A("add.s64 %s, %%%d, w;" % (ptr, base))
    for i in range(n):
        A("mov.f32 %%%d, 0f00000000;" % i)
    for k, (ptr, pxn, pyn) in enumerate([("pa", "px0", "py0"), ("pb", "px1", "py0"), ("pc", "px0", "py1"), ("pd", "px1", "py1")]):
        A("and.pred q, %s, %s;" % (pxn, pyn))
        for c in range(0, cpt, 4):
            b = k * cpt + c
            o = "" if c == 0 else "+%d" % (4 * c)
            A("@q ld.global.nc.v4.f32 {%%%d, %%%d, %%%d, %%%d}, [%s%s];" % (b, b + 1, b + 2, b + 3, ptr, o))
    if not split:
        A("REBUILD:")
        rebuild_lines(A, cpt)
    A("SAME:")
    A("}")
    body = "\n        ".join('"%s\\n\\t"' % x for x in L)
    ops = tex_operands(cpt)
    ops.append('"+r"(cur_key)')
    if split:
        ops.append('"+r"(moved)')
        outs = ",\n          ".join(", ".join(ops[i:i + 4]) for i in range(0, len(ops), 4))
        return """// Split form, first half: when `key` moved, start the loads of the new footprint into the cache registers
// (raw corners a, b, c, d; corners outside the image = 0), update `cur_key` and set bit VPLUS1 of `moved`.
// rebuild_off() turns the corners into A, B, C, D once they are needed, so the load latency overlaps whatever
// the caller puts between the two.
template <int VPLUS1, int TEXEL_BYTES>
__device__ __forceinline__ void issue_off(float2 (&t)[4][%d], unsigned& cur_key, unsigned& moved, unsigned key,
                                          const float* base, unsigned row_bytes, unsigned view_bytes, int width,
                                          int height) {
    asm volatile(
        %s
        : %s
        : "r"(key), "l"(base), "r"(row_bytes), "r"(view_bytes), "r"(width), "r"(height), "n"(VPLUS1),
          "n"(TEXEL_BYTES), "n"(TEXEL_BYTES + 16));
}
""" % (cpt // 2, body, outs)
    outs = ",\n          ".join(", ".join(ops[i:i + 4]) for i in range(0, len(ops), 4))
    return """// Offset-keyed variant (sweep_lean.cuh): `key` comes from project_off() -- for a footprint whose four corners
// are inside the image it IS the byte offset of the north-west texel from `base` (view offset included), so
// the common case is two address adds, four loads and the packed rebuild; border / outside footprints carry
// their corner instead (bit 0 set) and take the predicated path.  `cur_key` is updated in place.
template <int VPLUS1, int TEXEL_BYTES>
__device__ __forceinline__ void refetch_off(float2 (&t)[4][%d], unsigned& cur_key, unsigned key, const float* base,
                                            unsigned row_bytes, unsigned view_bytes, int width, int height) {
    asm volatile(
        %s
        : %s
        : "r"(key), "l"(base), "r"(row_bytes), "r"(view_bytes), "r"(width), "r"(height), "n"(VPLUS1),
          "n"(TEXEL_BYTES), "n"(TEXEL_BYTES + 16));
}
""" % (cpt // 2, body, outs)


def block_tok(cpt, shift=False):
    """Token-keyed re-fetch (sweep_quad.cuh): the projecting lane publishes only the packed floor corner
    (low 16 bits of the two round-down-add results); address and border handling live here, on the rare path.
    shift: A = a - ref (the variance volume of the long sweeps, see rebuild_lines)."""
    n = 4 * cpt
    old, key, base, rowb, vtex, wid, hei, wm1, hm1, vplus1, texb, texb16 = (n + i for i in range(12))
    ref0 = n + 12
    L = []
    A = L.append
    A("{")
    A(".reg .pred p, q, r, px0, px1, py0, py1;")
    A(".reg .b32 x0, y0, x1, y1, t;")
    A(".reg .b64 w, pa, pc, u0, u1, u2, u3;")
    A("setp.eq.u32 p, %%%d, %%%d;" % (key, old))
    A("@p bra SAME;")
    A("mov.u32 %%%d, %%%d;" % (old, key))
    A("bfe.s32 x0, %%%d, 0, 16;" % key)
    A("bfe.s32 y0, %%%d, 16, 16;" % key)
    A("mad.lo.s32 t, y0, %%%d, x0;" % wid)
    A("add.s32 t, t, %%%d;" % vtex)                 # + the texel index at which this view starts in `base` (SweepParams.view_tex)
    A("mad.wide.s32 pa, t, %%%d, %%%d;" % (texb, base))
    A("cvt.u64.u32 w, %%%d;" % rowb)
    A("add.s64 pc, pa, w;")
    A("setp.lt.u32 q, x0, %%%d;" % wm1)
    A("setp.lt.u32 r, y0, %%%d;" % hm1)
    A("and.pred q, q, r;")
    A("@!q bra SPECIAL;")
    for k, (ptr, off) in enumerate([("pa", 0), ("pa", 1), ("pc", 0), ("pc", 1)]):
        for c in range(0, cpt, 4):
            b = k * cpt + c
            assert c in (0, 4)
            if off:
                o = "+%%%d" % (texb16 if c else texb)
            else:
                o = "+16" if c else ""
            A("ld.global.nc.v4.f32 {%%%d, %%%d, %%%d, %%%d}, [%s%s];" % (b, b + 1, b + 2, b + 3, ptr, o))
    A("bra REBUILD;")
    A("SPECIAL:")
    # zeros padding: corners outside the image stay 0 (loads predicated off, nothing is clamped)
    A("add.s32 x1, x0, 1;")
    A("add.s32 y1, y0, 1;")
    A("setp.lt.u32 px0, x0, %%%d;" % wid)
    A("setp.lt.u32 px1, x1, %%%d;" % wid)
    A("setp.lt.u32 py0, y0, %%%d;" % hei)
    A("setp.lt.u32 py1, y1, %%%d;" % hei)
    for i in range(n):
        A("mov.f32 %%%d, 0f00000000;" % i)
    for k, (ptr, off, pxn, pyn) in enumerate([("pa", 0, "px0", "py0"), ("pa", 1, "px1", "py0"), ("pc", 0, "px0", "py1"), ("pc", 1, "px1", "py1")]):
        A("and.pred q, %s, %s;" % (pxn, pyn))
        for c in range(0, cpt, 4):
            b = k * cpt + c
            if off:
                o = "+%%%d" % (texb16 if c else texb)
            else:
                o = "+16" if c else ""
            A("@q ld.global.nc.v4.f32 {%%%d, %%%d, %%%d, %%%d}, [%s%s];" % (b, b + 1, b + 2, b + 3, ptr, o))
    A("REBUILD:")
    rebuild_lines(A, cpt, ref0 if shift else None)
    A("SAME:")
    A("}")
    body = "\n        ".join('"%s\\n\\t"' % x for x in L)
    ops = tex_operands(cpt)
    ops.append('"+r"(cur_key)')
    outs = ",\n          ".join(", ".join(ops[i:i + 4]) for i in range(0, len(ops), 4))
    if shift:
        refs = ", ".join('"f"(ref[%d].%s)' % (c // 2, "xy"[c % 2]) for c in range(cpt))
        return """// Token-keyed re-fetch with the reference texel subtracted from A (variance volume, long sweeps): see block_tok.
template <int VPLUS1, int TEXEL_BYTES>
__device__ __forceinline__ void refetch_tok_shift(float2 (&t)[4][%d], unsigned& cur_key, unsigned key, const float* base,
                                                  unsigned row_bytes, int view_tex, int width, int height, const float2 (&ref)[%d]) {
    asm volatile(
        %s
        : %s
        : "r"(key), "l"(base), "r"(row_bytes), "r"(view_tex), "r"(width), "r"(height), "r"(width - 1), "r"(height - 1),
          "n"(VPLUS1), "n"(TEXEL_BYTES), "n"(TEXEL_BYTES + 16), %s);
}
""" % (cpt // 2, cpt // 2, body, outs, refs)
    return """// Token-keyed variant (sweep_quad.cuh): `key` = (y0 & 0xffff) << 16 | (x0 & 0xffff), the floor corner of the
// footprint as it falls out of the round-down adds of the projection.  Everything else -- the texel address
// (64-bit: no size limit on `feats`), the interior test, zeros padding at the border (loads predicated off) and
// the rebuild of A, B, C, D -- happens here, on the rare path.  `cur_key` is updated in place.
template <int VPLUS1, int TEXEL_BYTES>
__device__ __forceinline__ void refetch_tok(float2 (&t)[4][%d], unsigned& cur_key, unsigned key, const float* base,
                                            unsigned row_bytes, int view_tex, int width, int height) {
    asm volatile(
        %s
        : %s
        : "r"(key), "l"(base), "r"(row_bytes), "r"(view_tex), "r"(width), "r"(height), "r"(width - 1), "r"(height - 1),
          "n"(VPLUS1), "n"(TEXEL_BYTES), "n"(TEXEL_BYTES + 16));
}
""" % (cpt // 2, body, outs)


def block_tok_split(cpt):
    """The token-keyed re-fetch in two halves: issue_tok() starts the loads of a moved footprint into the cache
    registers (raw corners) and records the view in `moved`; rebuild_tok() turns the corners into A, B, C, D.  A
    caller that issues every view before it rebuilds any has the loads of all views in flight at once -- what the
    short sweeps of the cascade need, where a footprint lasts 1-3 planes and the re-fetch latency, paid once per
    view in the one-block form, is half of all stall samples (profiles/ncu_r1_cfg3.txt)."""
    n = 4 * cpt
    old, mv, key, base, rowb, vtex, wid, hei, wm1, hm1, vplus1, texb, texb16 = (n + i for i in range(13))
    L = []
    A = L.append
    A("{")
    A(".reg .pred p, q, r, px0, px1, py0, py1;")
    A(".reg .b32 x0, y0, x1, y1, t;")
    A(".reg .b64 w, pa, pc;")
    A("setp.eq.u32 p, %%%d, %%%d;" % (key, old))
    A("@p bra SAME;")
    A("mov.u32 %%%d, %%%d;" % (old, key))
    A("or.b32 %%%d, %%%d, 1 << %%%d;" % (mv, mv, vplus1))
    A("bfe.s32 x0, %%%d, 0, 16;" % key)
    A("bfe.s32 y0, %%%d, 16, 16;" % key)
    A("mad.lo.s32 t, y0, %%%d, x0;" % wid)
    A("add.s32 t, t, %%%d;" % vtex)                 # + the texel index at which this view starts in `base` (SweepParams.view_tex)
    A("mad.wide.s32 pa, t, %%%d, %%%d;" % (texb, base))
    A("cvt.u64.u32 w, %%%d;" % rowb)
    A("add.s64 pc, pa, w;")
    A("setp.lt.u32 q, x0, %%%d;" % wm1)
    A("setp.lt.u32 r, y0, %%%d;" % hm1)
    A("and.pred q, q, r;")
    A("@!q bra SPECIAL;")
    for k, (ptr, off) in enumerate([("pa", 0), ("pa", 1), ("pc", 0), ("pc", 1)]):
        for c in range(0, cpt, 4):
            b = k * cpt + c
            if off:
                o = "+%%%d" % (texb16 if c else texb)
            else:
                o = "+16" if c else ""
            A("ld.global.nc.v4.f32 {%%%d, %%%d, %%%d, %%%d}, [%s%s];" % (b, b + 1, b + 2, b + 3, ptr, o))
    A("bra SAME;")
    A("SPECIAL:")
    A("add.s32 x1, x0, 1;")
    A("add.s32 y1, y0, 1;")
    A("setp.lt.u32 px0, x0, %%%d;" % wid)
    A("setp.lt.u32 px1, x1, %%%d;" % wid)
    A("setp.lt.u32 py0, y0, %%%d;" % hei)
    A("setp.lt.u32 py1, y1, %%%d;" % hei)
    for i in range(n):
        A("mov.f32 %%%d, 0f00000000;" % i)
    for k, (ptr, off, pxn, pyn) in enumerate([("pa", 0, "px0", "py0"), ("pa", 1, "px1", "py0"), ("pc", 0, "px0", "py1"), ("pc", 1, "px1", "py1")]):
        A("and.pred q, %s, %s;" % (pxn, pyn))
        for c in range(0, cpt, 4):
            b = k * cpt + c
            if off:
                o = "+%%%d" % (texb16 if c else texb)
            else:
                o = "+16" if c else ""
            A("@q ld.global.nc.v4.f32 {%%%d, %%%d, %%%d, %%%d}, [%s%s];" % (b, b + 1, b + 2, b + 3, ptr, o))
    A("SAME:")
    A("}")
    body = "\n        ".join('"%s\\n\\t"' % x for x in L)
    ops = tex_operands(cpt)
    ops.append('"+r"(cur_key)')
    ops.append('"+r"(moved)')
    outs = ",\n          ".join(", ".join(ops[i:i + 4]) for i in range(0, len(ops), 4))
    issue = """// Split token-keyed re-fetch, first half (see block_tok_split in tools/gen_refetch.py): when `key` moved, start the
// loads of the new footprint into the cache registers (raw corners a, b, c, d; corners outside the image = 0), update
// `cur_key` and set bit VPLUS1 of `moved`.
template <int VPLUS1, int TEXEL_BYTES>
__device__ __forceinline__ void issue_tok(float2 (&t)[4][%d], unsigned& cur_key, unsigned& moved, unsigned key,
                                          const float* base, unsigned row_bytes, int view_tex, int width, int height) {
    asm volatile(
        %s
        : %s
        : "r"(key), "l"(base), "r"(row_bytes), "r"(view_tex), "r"(width), "r"(height), "r"(width - 1), "r"(height - 1),
          "n"(VPLUS1), "n"(TEXEL_BYTES), "n"(TEXEL_BYTES + 16));
}
""" % (cpt // 2, body, outs)
    L = []
    A = L.append
    A("{")
    A(".reg .pred p;")
    A(".reg .b32 t;")
    A(".reg .b64 u0, u1, u2, u3;")
    A("and.b32 t, %%%d, 1 << %%%d;" % (n, n + 1))
    A("setp.eq.u32 p, t, 0;")
    A("@p bra DONE;")
    rebuild_lines(A, cpt)
    A("DONE:")
    A("}")
    body = "\n        ".join('"%s\\n\\t"' % x for x in L)
    outs = ",\n          ".join(", ".join(tex_operands(cpt)[i:i + 4]) for i in range(0, 4 * cpt, 4))
    rebuild = """// Second half: corners a, b, c, d -> A = a, B = b - a, C = c - a, D = (d - c) - (b - a) for the views issue_tok() marked.
template <int VPLUS1>
__device__ __forceinline__ void rebuild_tok(float2 (&t)[4][%d], unsigned moved) {
    asm volatile(
        %s
        : %s
        : "r"(moved), "n"(VPLUS1));
}
""" % (cpt // 2, body, outs)
    return issue + "\n" + rebuild


def block_ws(cpt, shift):
    """Consumer side of the warp-specialised sweep (sweep_ws.cuh).  `info` is what the producer warp published for this
    (pixel, view, plane): 0 = the footprint did not move; bit 0 set = the four corners wait in a shared-memory slot
    (info & ~1 = its address: [corner][32 channels], copied there by the producer with cp.async, corners outside the
    image already zero); otherwise 2 | (x0 + 8) << 2 | (y0 + 8) << 17 = the floor corner, to be fetched from global
    memory here (the slot region was full).  Either way the corners become A, B, C, D in place; with `shift` the
    reference texel is subtracted from A (variance is shift invariant: the reference view then contributes 0)."""
    n = 4 * cpt
    info, laneoff, base, rowb, wid, hei, texb, texb16 = (n + i for i in range(8))
    ref0 = n + 8
    L = []
    A = L.append
    A("{")
    A(".reg .pred p, q, px0, px1, py0, py1;")
    A(".reg .b32 x0, y0, x1, y1, t, sa;")
    A(".reg .b64 w, pa, pc, u0, u1, u2, u3;")
    A("setp.eq.u32 p, %%%d, 0;" % info)
    A("@p bra DONE;")
    A("and.b32 t, %%%d, 1;" % info)
    A("setp.eq.u32 q, t, 0;")
    A("@q bra DIRECT;")
    A("add.u32 sa, %%%d, %%%d;" % (info, laneoff))          # laneoff = lane's channel offset in a texel - 1
    for k in range(4):
        for c in range(0, cpt, 4):
            b = k * cpt + c
            A("ld.shared.v4.f32 {%%%d, %%%d, %%%d, %%%d}, [sa+%d];" % (b, b + 1, b + 2, b + 3, 128 * k + 4 * c))
    A("bra REBUILD;")
    A("DIRECT:")
    A("shr.u32 x0, %%%d, 2;" % info)
    A("and.b32 x0, x0, 0x7fff;")
    A("sub.s32 x0, x0, 8;")
    A("shr.u32 y0, %%%d, 17;" % info)
    A("sub.s32 y0, y0, 8;")
    A("mad.lo.s32 t, y0, %%%d, x0;" % wid)
    A("mad.wide.s32 pa, t, %%%d, %%%d;" % (texb, base))
    A("cvt.u64.u32 w, %%%d;" % rowb)
    A("add.s64 pc, pa, w;")
    A("add.s32 x1, x0, 1;")
    A("add.s32 y1, y0, 1;")
    A("setp.lt.u32 px0, x0, %%%d;" % wid)
    A("setp.lt.u32 px1, x1, %%%d;" % wid)
    A("setp.lt.u32 py0, y0, %%%d;" % hei)
    A("setp.lt.u32 py1, y1, %%%d;" % hei)
    for i in range(n):
        A("mov.f32 %%%d, 0f00000000;" % i)
    for k, (ptr, off, pxn, pyn) in enumerate([("pa", 0, "px0", "py0"), ("pa", 1, "px1", "py0"), ("pc", 0, "px0", "py1"), ("pc", 1, "px1", "py1")]):
        A("and.pred q, %s, %s;" % (pxn, pyn))
        for c in range(0, cpt, 4):
            b = k * cpt + c
            if off:
                o = "+%%%d" % (texb16 if c else texb)
            else:
                o = "+16" if c else ""
            A("@q ld.global.nc.v4.f32 {%%%d, %%%d, %%%d, %%%d}, [%s%s];" % (b, b + 1, b + 2, b + 3, ptr, o))
    A("REBUILD:")
    for c in range(0, cpt, 2):
        a, b, cc, d = c, cpt + c, 2 * cpt + c, 3 * cpt + c
        for reg, i in (("u0", a), ("u1", b), ("u2", cc), ("u3", d)):
            A("mov.b64 %s, {%%%d, %%%d};" % (reg, i, i + 1))
        A("sub.rn.f32x2 u1, u1, u0;")
        A("sub.rn.f32x2 u3, u3, u2;")
        A("sub.rn.f32x2 u2, u2, u0;")
        A("sub.rn.f32x2 u3, u3, u1;")
        for reg, i in (("u1", b), ("u2", cc), ("u3", d)):
            A("mov.b64 {%%%d, %%%d}, %s;" % (i, i + 1, reg))
        if shift:
            A("mov.b64 u1, {%%%d, %%%d};" % (ref0 + c, ref0 + c + 1))
            A("sub.rn.f32x2 u0, u0, u1;")
            A("mov.b64 {%%%d, %%%d}, u0;" % (a, a + 1))
    A("DONE:")
    A("}")
    body = "\n        ".join('"%s\\n\\t"' % x for x in L)
    ops = tex_operands(cpt)
    outs = ",\n          ".join(", ".join(ops[i:i + 4]) for i in range(0, len(ops), 4))
    refs = ", ".join('"f"(ref[%d].%s)' % (c // 2, "xy"[c % 2]) for c in range(cpt)) if shift else ""
    name = "consume_ws_shift" if shift else "consume_ws"
    sig_ref = ", const float2 (&ref)[%d]" % (cpt // 2) if shift else ""
    return """// %s: see block_ws in tools/gen_refetch.py.
template <int TEXEL_BYTES>
__device__ __forceinline__ void %s(float2 (&t)[4][%d], unsigned info, unsigned lane_off, const float* base,
                                   unsigned row_bytes, int width, int height%s) {
    asm volatile(
        %s
        : %s
        : "r"(info), "r"(lane_off), "l"(base), "r"(row_bytes), "r"(width), "r"(height), "n"(TEXEL_BYTES),
          "n"(TEXEL_BYTES + 16)%s);
}
""" % (name, name, cpt // 2, sig_ref, body, outs, (",\n          " + refs) if shift else "")


def block_dot(cpt):
    """Token-keyed re-fetch for the correlation volumes (group-wise correlation, AdaMVS pair volumes): what those
    modes need of a footprint is sum_c ref_c * (A_c + fx B_c + fy C_c + fxy D_c) over the lane's channels, i.e. the
    four DOT PRODUCTS PA = sum ref_c A_c, PB, PC, PD -- 4 registers per view instead of 16, and 3 FMAs per view and
    plane instead of 3 per channel.  The corners live in block-local registers only while the dots are formed."""
    n = 4                                   # outputs: PA, PB, PC, PD
    old, key, base, rowb, vtex, wid, hei, wm1, hm1, vplus1, texb, texb16 = (n + i for i in range(12))
    ref0 = n + 12
    L = []
    A = L.append
    A("{")
    A(".reg .pred p, q, r, px0, px1, py0, py1;")
    A(".reg .b32 x0, y0, x1, y1, t;")
    A(".reg .b64 w, pa, pc, u0, u1, u2, u3;")
    A(".reg .f32 ca<%d>, cb<%d>, cc<%d>, cd<%d>;" % (cpt, cpt, cpt, cpt))
    A("setp.eq.u32 p, %%%d, %%%d;" % (key, old))
    A("@p bra SAME;")
    A("mov.u32 %%%d, %%%d;" % (old, key))
    A("bfe.s32 x0, %%%d, 0, 16;" % key)
    A("bfe.s32 y0, %%%d, 16, 16;" % key)
    A("mad.lo.s32 t, y0, %%%d, x0;" % wid)
    A("add.s32 t, t, %%%d;" % vtex)                 # + the texel index at which this view starts in `base` (SweepParams.view_tex)
    A("mad.wide.s32 pa, t, %%%d, %%%d;" % (texb, base))
    A("cvt.u64.u32 w, %%%d;" % rowb)
    A("add.s64 pc, pa, w;")
    A("setp.lt.u32 q, x0, %%%d;" % wm1)
    A("setp.lt.u32 r, y0, %%%d;" % hm1)
    A("and.pred q, q, r;")
    A("@!q bra SPECIAL;")
    names = ["ca", "cb", "cc", "cd"]
    for k, (ptr, off) in enumerate([("pa", 0), ("pa", 1), ("pc", 0), ("pc", 1)]):
        o = ("+%%%d" % texb) if off else ""
        A("ld.global.nc.v4.f32 {%s0, %s1, %s2, %s3}, [%s%s];" % (names[k], names[k], names[k], names[k], ptr, o))
    A("bra REBUILD;")
    A("SPECIAL:")
    A("add.s32 x1, x0, 1;")
    A("add.s32 y1, y0, 1;")
    A("setp.lt.u32 px0, x0, %%%d;" % wid)
    A("setp.lt.u32 px1, x1, %%%d;" % wid)
    A("setp.lt.u32 py0, y0, %%%d;" % hei)
    A("setp.lt.u32 py1, y1, %%%d;" % hei)
    for nm in names:
        for c in range(cpt):
            A("mov.f32 %s%d, 0f00000000;" % (nm, c))
    for k, (ptr, off, pxn, pyn) in enumerate([("pa", 0, "px0", "py0"), ("pa", 1, "px1", "py0"), ("pc", 0, "px0", "py1"), ("pc", 1, "px1", "py1")]):
        A("and.pred q, %s, %s;" % (pxn, pyn))
        o = ("+%%%d" % texb) if off else ""
        A("@q ld.global.nc.v4.f32 {%s0, %s1, %s2, %s3}, [%s%s];" % (names[k], names[k], names[k], names[k], ptr, o))
    A("REBUILD:")
    for c in range(cpt):
        A("sub.rn.f32 cb%d, cb%d, ca%d;" % (c, c, c))       # B = b - a
        A("sub.rn.f32 cd%d, cd%d, cc%d;" % (c, c, c))       # d - c
        A("sub.rn.f32 cc%d, cc%d, ca%d;" % (c, c, c))       # C = c - a
        A("sub.rn.f32 cd%d, cd%d, cb%d;" % (c, c, c))       # D = (d - c) - (b - a)
    for out, nm in enumerate(names):
        A("mul.rn.f32 %%%d, %%%d, %s0;" % (out, ref0, nm))
        for c in range(1, cpt):
            A("fma.rn.f32 %%%d, %%%d, %s%d, %%%d;" % (out, ref0 + c, nm, c, out))
    A("SAME:")
    A("}")
    body = "\n        ".join('"%s\\n\\t"' % x for x in L)
    return """// Dot-product variant (sweep_quad.cuh, correlation volumes): see block_dot in tools/gen_refetch.py.  `dot` = {PA, PB, PC,
// PD} of this lane's %d channels against the reference texel `ref`; `cur_key` is updated in place.
template <int VPLUS1, int TEXEL_BYTES>
__device__ __forceinline__ void refetch_dot(float (&dot)[4], unsigned& cur_key, unsigned key, const float* base,
                                            unsigned row_bytes, int view_tex, int width, int height, const float (&ref)[%d]) {
    asm volatile(
        %s
        : "+f"(dot[0]), "+f"(dot[1]), "+f"(dot[2]), "+f"(dot[3]), "+r"(cur_key)
        : "r"(key), "l"(base), "r"(row_bytes), "r"(view_tex), "r"(width), "r"(height), "r"(width - 1), "r"(height - 1),
          "n"(VPLUS1), "n"(TEXEL_BYTES), "n"(TEXEL_BYTES + 16), %s);
}
""" % (cpt, cpt, body, ", ".join('"f"(ref[%d])' % c for c in range(cpt)))


HEADER = '''// GENERATED by tools/gen_refetch.py -- do not edit by hand.
//
// refetch_footprint(t, key, old_key, base, W-1, H-1, row_bytes, texel_bytes)
//   Re-fetches one source view's 2x2 texel footprint (this lane's channels of each corner) when its key
//   moved, and turns the corners a,b,c,d -- out-of-bounds ones zeroed (zeros padding), their addresses
//   clamped into the image -- into the interpolation coefficients A=a, B=b-a, C=c-a, D=a-b-c+d, all IN
//   PLACE, in one PTX block with its own branch (a no-op when key == old_key).
//     key  = [27:14] y0+4 | [13:0] x0+4   (floor corner of the footprint; bits 31:28 are ignored)
//     t[k][j]: corner k (nw, ne, sw, se), channel pair j of this lane;  base = view + lane's channel offset
//   Why PTX: any C++-level conditional update of the register-resident footprint cache made NVVM/ptxas
//   copy the whole cache on every plane's common path (24-32 MOVs per view per plane, measured in SASS);
//   an opaque block with read-write operands keeps the cache in fixed registers.
#pragma once
#include "common.cuh"

namespace d3d {

'''

if __name__ == "__main__":
    with open(OUT, "w") as f:
        f.write(HEADER + block(8) + "\n" + block(4) + "\n" + block_off(8) + "\n" + block_off(4) + "\n" + "\n" + block_tok(4) + "\n" + block_tok(4, True) + "\n" + block_tok_split(4) + "\n" + block_ws(4, True) + "\n" + block_ws(4, False) + "\n" + block_dot(4) + "\n}  // namespace d3d\n")
    print("wrote", OUT)

// Packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2) and the two-pixels-at-a-time versions of the per-pixel
// functions of be_math.cuh.
//
// sm_100a issues `fma.rn.f32x2` as ONE warp instruction that occupies the FMA pipe for two cycles: the same FLOP rate as
// two FFMAs in half the issue slots (tools/microbench/ffma2.cu, measured on B200: FFMA 0.96, FFMA2 0.49 warp-inst/clk/SMSP
// at the same 71-73 TFLOP/s; a 1:2 FFMA2:LOP3 mix runs 1.27x faster than the equivalent 2:2 FFMA:LOP3 mix).  The render
// kernels are issue bound and every thread owns two pixel slots that see the same patch parameters, so the two slots are
// computed as the two halves of a packed register pair.  ptxas folds broadcast scalars, immediates and negations into the
// FFMA2 operands (no MOVs), and each half is an IEEE fp32 rn operation: results are bit-identical to the scalar functions,
// which remain the specification (tests/test_hostmath_vs_oracle.py checks f2 against them on the host).
#pragma once
#include "be_math.cuh"

#if defined(__CUDA_ARCH__)
struct f2 { unsigned long long v; };
BE_HD f2 mk2(float a, float b) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
BE_HD float lo(f2 r) { float a; asm("{.reg .f32 t; mov.b64 {%0,t}, %1;}" : "=f"(a) : "l"(r.v)); return a; }
BE_HD float hi(f2 r) { float b; asm("{.reg .f32 t; mov.b64 {t,%0}, %1;}" : "=f"(b) : "l"(r.v)); return b; }
BE_HD f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return d; }
BE_HD f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v)); return d; }
BE_HD f2 add2(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v)); return d; }
BE_HD f2 sub2(f2 a, f2 b) { f2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v)); return d; }
#else
struct f2 { float x, y; };
BE_HD f2 mk2(float a, float b) { f2 r; r.x = a; r.y = b; return r; }
BE_HD float lo(f2 r) { return r.x; }
BE_HD float hi(f2 r) { return r.y; }
BE_HD f2 fma2(f2 a, f2 b, f2 c) { return mk2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
BE_HD f2 mul2(f2 a, f2 b) { return mk2(a.x * b.x, a.y * b.y); }
BE_HD f2 add2(f2 a, f2 b) { return mk2(a.x + b.x, a.y + b.y); }
BE_HD f2 sub2(f2 a, f2 b) { return mk2(a.x - b.x, a.y - b.y); }
#endif
BE_HD f2 bc2(float a) { return mk2(a, a); }

// be_edge for two pixels (be_math.cuh:be_edge)
BE_HD f2 be_edge2(f2 dx, f2 dy, float sn, float cs, float w) {
    const f2 d = fma2(bc2(cs), dy, mul2(bc2(-sn), dx));
    const f2 a = fma2(bc2(cs), dx, mul2(bc2(sn), dy));
    const f2 aw = mul2(a, bc2(w));
    const f2 c2 = fma2(d, d, mul2(aw, aw));
    const float c0 = be_sqrt(lo(c2)), c1 = be_sqrt(hi(c2));
    const float d0 = lo(d), d1 = hi(d);
    return mk2((lo(a) < 0.0f) ? ((d0 < 0.0f) ? -c0 : c0) : d0, (hi(a) < 0.0f) ? ((d1 < 0.0f) ? -c1 : c1) : d1);
}

BE_HD float be_wedge_dist(float DA, float DB, float f, bool strict) {
    const bool in = strict ? ((f * DA > 0.0f) && (f * DB < 0.0f)) : ((f * DA >= 0.0f) && (f * DB <= 0.0f));
    return fminf(fabsf(DA), fabsf(DB)) * (in ? f : -f);
}

// be_pixel_dists for two pixels
BE_HD void be_pixel_dists2(const BePatch& P, f2 X, f2 Y, float w, f2* d1, f2* d2) {
    {
        const f2 dx = sub2(X, bc2(P.vx[0])), dy = sub2(Y, bc2(P.vy[0]));
        const f2 DA = be_edge2(dx, dy, P.sn[0], P.cs[0], w), DB = be_edge2(dx, dy, P.sn[1], P.cs[1], w);
        *d1 = mk2(be_wedge_dist(lo(DA), lo(DB), P.flip[0], true), be_wedge_dist(hi(DA), hi(DB), P.flip[0], true));
    }
    {
        const f2 dx = sub2(X, bc2(P.vx[1])), dy = sub2(Y, bc2(P.vy[1]));
        const f2 DA = be_edge2(dx, dy, P.sn[2], P.cs[2], w), DB = be_edge2(dx, dy, P.sn[3], P.cs[3], w);
        *d2 = mk2(be_wedge_dist(lo(DA), lo(DB), P.flip[1], false), be_wedge_dist(hi(DA), hi(DB), P.flip[1], false));
    }
}

// be_h for two pixels
BE_HD f2 be_h2(f2 dist, float inv_eta) {
    const f2 t = mul2(dist, bc2(inv_eta));
    const f2 a = mk2(fminf(fabsf(lo(t)), 4.0f), fminf(fabsf(hi(t)), 4.0f));
    f2 p = bc2(-4.5357578e-05f);
    p = fma2(p, a, bc2(4.4549927e-04f));
    p = fma2(p, a, bc2(-1.4894147e-03f));
    p = fma2(p, a, bc2(-7.7467301e-04f));
    p = fma2(p, a, bc2(2.8253718e-02f));
    p = fma2(p, a, bc2(-1.4848163e-01f));
    p = fma2(p, a, bc2(-9.1841639e-01f));
    p = fma2(p, a, bc2(-1.6279086e+00f));
    p = fma2(p, a, bc2(-1.0f));
    const f2 r = sub2(bc2(0.5f), mk2(be_exp2(lo(p)), be_exp2(hi(p))));
    return add2(bc2(0.5f), mk2(copysignf(lo(r), lo(t)), copysignf(hi(r), hi(t))));
}

// be_boundary for two pixels
BE_HD f2 be_boundary2(f2 d1, f2 d2) {
    const float a0 = fminf(fabsf(lo(d1)), fabsf(lo(d2))), a1 = fminf(fabsf(hi(d1)), fabsf(hi(d2)));
    const f2 dB = mk2((lo(d2) >= 0.0f) ? lo(d2) : a0, (hi(d2) >= 0.0f) ? hi(d2) : a1);
    const f2 e = mul2(mul2(dB, dB), bc2(-(1.44269504f / (BE_DELTA * BE_DELTA))));
    return mk2(be_exp2(lo(e)), be_exp2(hi(e)));
}

// be_mask as two float weights: m1 = [mask == 1], m2 = [mask == 2]
BE_HD void be_mask_weights(float d1, float d2, bool densify_w, float* m1, float* m2) {
    const int mk = be_mask(d1, d2, densify_w);
    *m1 = (mk == 1) ? 1.0f : 0.0f;
    *m2 = (mk == 2) ? 1.0f : 0.0f;
}

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

// ------------------------------------------------------------------------------------------
// backward pieces, two pixels at a time (be_math.cuh: be_h_grad, be_wedges_backward, be_boundary_backward, be_wedge_backward)
BE_HD f2 neg2(f2 a) { return mk2(-lo(a), -hi(a)); }
BE_HD f2 sel2(bool c0, bool c1, f2 a, f2 b) { return mk2(c0 ? lo(a) : lo(b), c1 ? hi(a) : hi(b)); }
BE_HD float be_sign(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

BE_HD void be_h_grad2(f2 dist, float inv_eta, f2* dh_dd, f2* dh_deta) {
    const f2 t = mul2(dist, bc2(inv_eta));
    const f2 x = mul2(mul2(t, t), bc2(-1.44269504f));
    const f2 e = mul2(bc2(BE_INV_SQRT_PI), mk2(be_exp2(lo(x)), be_exp2(hi(x))));
    *dh_dd = mul2(e, bc2(inv_eta));
    *dh_deta = mul2(mul2(neg2(e), t), bc2(BE_SQRT2_F * inv_eta));
}

BE_HD void be_boundary_backward1(float d1, float d2, float lb, float glb, float* gd1, float* gd2) {
    float a = 0.0f, b = 0.0f;
    be_boundary_backward(d1, d2, lb, glb, &a, &b);
    *gd1 = a; *gd2 = b;
}

// be_wedge_backward for two pixels; acc[0..3] are packed partial sums (vertex x, vertex y, angle A, angle B)
BE_HD void be_wedge_backward2(const BePatch& P, int k, f2 X, f2 Y, float w, f2 g, f2* acc) {
    const f2 dx = sub2(X, bc2(P.vx[k])), dy = sub2(Y, bc2(P.vy[k]));
    const float f = P.flip[k];
    const float snA = P.sn[2 * k], csA = P.cs[2 * k], snB = P.sn[2 * k + 1], csB = P.cs[2 * k + 1];
    const f2 dA = fma2(bc2(csA), dy, mul2(bc2(-snA), dx)), aA = fma2(bc2(csA), dx, mul2(bc2(snA), dy));
    const f2 dB = fma2(bc2(csB), dy, mul2(bc2(-snB), dx)), aB = fma2(bc2(csB), dx, mul2(bc2(snB), dy));
    const f2 awA = mul2(aA, bc2(w)), awB = mul2(aB, bc2(w));
    const f2 cA2 = fma2(dA, dA, mul2(awA, awA)), cB2 = fma2(dB, dB, mul2(awB, awB));
    // Per edge: D = a<0 ? sign(d) r : d.  d|D|/dd = d/|D| on both branches (sign(d) = d/|d| where |D| = |d|), d|D|/da = w^2 a/|D| on the
    // cap branch only; 1/|D| from MUFU.RCP, 0 at |D| = 0 (abs'(0) = 0).
    float sA[2], sB[2], irA[2], irB[2], mA[2], mB[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const float dAs = s ? hi(dA) : lo(dA), aAs = s ? hi(aA) : lo(aA), dBs = s ? hi(dB) : lo(dB), aBs = s ? hi(aB) : lo(aB);
        const float capA = be_sqrt(s ? hi(cA2) : lo(cA2)), capB = be_sqrt(s ? hi(cB2) : lo(cB2));
        const float DA = (aAs < 0.0f) ? ((dAs < 0.0f) ? -capA : capA) : dAs;
        const float DB = (aBs < 0.0f) ? ((dBs < 0.0f) ? -capB : capB) : dBs;
        const bool in = (k == 0) ? ((f * DA > 0.0f) && (f * DB < 0.0f)) : ((f * DA >= 0.0f) && (f * DB <= 0.0f));
        const float sg = (s ? hi(g) : lo(g)) * (in ? f : -f);
        const float absA = fabsf(DA), absB = fabsf(DB);
        const float wA = (absA < absB) ? 1.0f : ((absA == absB) ? 0.5f : 0.0f);      // min(): ties split 1/2
        sA[s] = sg * wA; sB[s] = sg * (1.0f - wA);
        irA[s] = (absA > 0.0f) ? be_rcp(absA) : 0.0f; irB[s] = (absB > 0.0f) ? be_rcp(absB) : 0.0f;
        mA[s] = (aAs < 0.0f) ? w * w : 0.0f; mB[s] = (aBs < 0.0f) ? w * w : 0.0f;
    }
    const f2 tA = mul2(mk2(sA[0], sA[1]), mk2(irA[0], irA[1])), tB = mul2(mk2(sB[0], sB[1]), mk2(irB[0], irB[1]));
    const f2 GdA = mul2(tA, dA), GaA = mul2(mul2(tA, mk2(mA[0], mA[1])), aA);
    const f2 GdB = mul2(tB, dB), GaB = mul2(mul2(tB, mk2(mB[0], mB[1])), aB);
    // acc[0] += (gdA snA - gaA csA) + (gdB snB - gaB csB)
    acc[0] = add2(acc[0], add2(fma2(GdA, bc2(snA), mul2(GaA, bc2(-csA))), fma2(GdB, bc2(snB), mul2(GaB, bc2(-csB)))));
    // acc[1] += (-gdA csA - gaA snA) + (-gdB csB - gaB snB)
    acc[1] = add2(acc[1], add2(fma2(GdA, bc2(-csA), mul2(GaA, bc2(-snA))), fma2(GdB, bc2(-csB), mul2(GaB, bc2(-snB)))));
    acc[2] = add2(acc[2], fma2(GaA, dA, mul2(neg2(GdA), aA)));     // -gdA aA + gaA dA
    acc[3] = add2(acc[3], fma2(GaB, dB, mul2(neg2(GdB), aB)));
}

// Per-patch / per-pixel arithmetic of the Blurry-Edges render -> depth path, written once as
// __host__ __device__ inline functions so that the very same source is (a) inlined into the
// sm_100a kernels of be_kernels.cu and (b) compiled for the host by oracle/be_hostmath.cpp, where
// the CPU test-suite checks it against the oracle without a GPU.
//
// Reference lines restated (paths relative to /root/reference):
//   utils/postprocessing_loss.py:15-17 (pixel grid), :26-86 (params2dists), :88-89 (params2etas),
//   :91-95 (dists2indicators), :97-98 (normalized_gaussian); utils/depth_etas.py:4-37;
//   blurry_edges_test.py:19-79, global_training.py:62-91,141-145 (script-level glue).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BE_HD __host__ __device__ __forceinline__
#else
#define BE_HD inline
#endif

#define BE_PI_F 3.14159274101257324f      // fp32(pi): what torch uses when a python float meets an fp32 tensor
#define BE_2PI_F 6.28318548202514648f     // fp32(2*pi)
#define BE_SQRT2_F 1.41421353816986084f   // torch.sqrt(torch.tensor(2)) is an fp32 value (postprocessing_loss.py:92)
#define BE_DELTA 0.07f                    // normalized_gaussian default (:97)
#define BE_ETA_SHARP 1e-4f                // blurry_edges_test.py:63

// how the 8 geometric + eta parameters of a patch arrive (same values as include/blurry_edges_b200.h)
#ifndef BE_PARAMS_RESTORED12
#define BE_PARAMS_RESTORED12 0   // [.,12] xy, wrapped angles, eta COEFFICIENTS (blurry_edges_test.py:135-138 already applied)
#define BE_PARAMS_RAW12 1        // [.,12] raw GlobalStage output; restore = global_training.py:141-145
#define BE_PARAMS_LOCAL10 2      // [.,10] xy, wrapped angles, 2 eta coefficients (pass A / pre-cal)
#define BE_PARAMS_LOCALRAW10 3   // [.,10] raw LocalStage output; angles wrapped here (local_training.py:33)
#endif

struct BeCam {   // utils/depth_etas.py:4-21 (host fills this, see be_host.cpp)
    float numerator, k_fac, k_const, k_root, intercept;
    float sin_w, cos_w, sin_m, cos_m;   // sin/cos of theta_wng = fp32(pi/4), theta_mid = fp32(3pi/4)
    float s, rho_prime;
};

struct BePatch {   // everything phase 1/2 need that is constant over the 21x21 pixels of one patch
    float sn[4], cs[4];      // edge e: 0 = wedge1 first (theta1), 1 = wedge1 second (theta1+phi1), 2,3 = wedge 2
    float vx[2], vy[2];      // wedge vertices
    float flip[2];           // +1 if (phi mod 2pi) < pi else -1            (:46-47)
    float eta[4];            // (w1 img1, w2 img1, w1 img2, w2 img2)         (blurry_edges_test.py:36-37,44-45)
    float inv_eta[4];        // 1 / (sqrt2_f32 * eta)
    float z[2];              // analytic depth of wedge 1 / wedge 2
};

// ------------------------------------------------------------------------------------------
BE_HD float be_axis(int j, int R) {
    // torch.linspace(-1,1,R) on CPU, fp32: fma(step, j, -1) below the midpoint, fma(-step, R-1-j, 1) above
    // (bit-checked against torch 2.11 for R = 5..33).
    const float step = 2.0f / (float)(R - 1);
    return (j < R / 2) ? fmaf(step, (float)j, -1.0f) : fmaf(-step, (float)(R - 1 - j), 1.0f);
}

BE_HD float be_wrap_2pi(float a) {   // torch.remainder(a, 2*pi) for fp32 (exact fmod + sign fix)
    float m = fmodf(a, BE_2PI_F);
    if (m != 0.0f && m < 0.0f) m += BE_2PI_F;
    return m;
}

BE_HD void be_sincos(float a, float* s, float* c) {
#if defined(__CUDA_ARCH__)
    sincosf(a, s, c);
#else
    *s = sinf(a); *c = cosf(a);
#endif
}

BE_HD float be_exp10(float x) {
#if defined(__CUDA_ARCH__)
    return exp10f(x);
#else
    return (float)pow(10.0, (double)x);
#endif
}

// 2^x and sqrt on the SFU (MUFU.EX2 / MUFU.SQRT, <= 2 ulp); libm on the host
BE_HD float be_exp2(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return exp2f(x);
#endif
}

BE_HD float be_sqrt(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return sqrtf(x);
#endif
}

BE_HD float be_rsqrt(float x) {   // MUFU.RSQ (<= 2 ulp)
#if defined(__CUDA_ARCH__)
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / sqrtf(x);
#endif
}

BE_HD float be_rcp(float x) {     // MUFU.RCP (<= 1 ulp)
#if defined(__CUDA_ARCH__)
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / x;
#endif
}

BE_HD float be_eta(float coef) { return be_exp10(erff(coef) * 2.0f - 2.0f); }   // :88-89

// sin/cos of (a + b) where a, b are fp32 angles: evaluate at s = fl(a+b) and correct with the exact
// rounding error e of the sum (TwoSum), so the result is the sine of the real-number sum.
BE_HD void be_sincos_sum(float a, float b, float* sn, float* cs) {
    const float s = a + b;
    const float bb = s - a;
    const float e = (a - (s - bb)) + (b - bb);
    float s0, c0;
    be_sincos(s, &s0, &c0);
    *sn = fmaf(e, c0, s0);
    *cs = fmaf(-e, s0, c0);
}

BE_HD float be_depth(const BeCam& cam, float e1, float e2) {   // utils/depth_etas.py:23-34
    const float c = cam.intercept;
    const float c1 = -cam.sin_w * e1 + cam.cos_w * (e2 - c);
    const float c2 = -cam.sin_m * (e1 - c) + cam.cos_m * e2;
    const float c3 = -cam.sin_w * (e1 - c) + cam.cos_w * e2;
    const float half = (e1 + e2 - c) * 0.5f;
    float a, b;
    if (c1 > 0.0f) { a = half; b = c + half; }
    else if (c2 > 0.0f) { a = c + (e1 - e2 - c) * 0.5f; b = (e2 - e1 + c) * 0.5f; }
    else if (c3 < 0.0f) { a = c + half; b = half; }
    else { a = e1; b = e2; }
    return cam.numerator / (cam.k_fac * (a * a - b * b) + cam.k_const);
}

BE_HD float be_refocus_sigma(const BeCam& cam, float z) {   // utils/depth_etas.py:36-37
    return fabsf((1.0f / z - cam.rho_prime) * cam.s + 1.0f) / cam.k_root;
}

// Fill a BePatch from the parameter vector of one patch.  `p` points at 12 (or 10) floats.
BE_HD void be_patch_setup(const float* p, int mode, const BeCam& cam, BePatch& P) {
    float xy[4], th[2], ph[2], coef[4];
    const bool raw12 = (mode == BE_PARAMS_RAW12);
    for (int k = 0; k < 4; ++k) xy[k] = raw12 ? p[k] * 3.0f : p[k];
    for (int k = 0; k < 2; ++k) {
        float a = p[4 + 2 * k], b = p[5 + 2 * k];
        if (raw12) { a = be_wrap_2pi((a + 1.0f) * BE_PI_F); b = be_wrap_2pi((b + 1.0f) * BE_PI_F); }
        else if (mode == BE_PARAMS_LOCALRAW10) { a = be_wrap_2pi(a); b = be_wrap_2pi(b); }
        th[k] = a; ph[k] = b;
    }
    const int neta = (mode == BE_PARAMS_LOCAL10 || mode == BE_PARAMS_LOCALRAW10) ? 2 : 4;
    for (int k = 0; k < 4; ++k) coef[k] = (k < neta) ? (raw12 ? p[8 + k] + 0.5f : p[8 + k]) : 0.0f;
    for (int k = 0; k < 2; ++k) {
        P.vx[k] = xy[2 * k]; P.vy[k] = xy[2 * k + 1];
        be_sincos(th[k], &P.sn[2 * k], &P.cs[2 * k]);
        be_sincos_sum(th[k], ph[k], &P.sn[2 * k + 1], &P.cs[2 * k + 1]);
        P.flip[k] = (be_wrap_2pi(ph[k]) < BE_PI_F) ? 1.0f : -1.0f;
    }
    for (int k = 0; k < 4; ++k) {
        P.eta[k] = (k < neta) ? be_eta(coef[k]) : 1.0f;
        P.inv_eta[k] = 1.0f / (BE_SQRT2_F * P.eta[k]);
    }
    if (neta == 4) {
        P.z[0] = be_depth(cam, P.eta[0], P.eta[2]);
        P.z[1] = be_depth(cam, P.eta[1], P.eta[3]);
    } else { P.z[0] = P.z[1] = 0.0f; }
}

// One half-line edge (:26-30,:57-78): returns D = a<0 ? sign(d)*sqrt(d^2+(a w)^2) : d
BE_HD float be_edge(float dx, float dy, float sn, float cs, float w) {
    const float d = fmaf(cs, dy, -sn * dx);
    const float a = fmaf(cs, dx, sn * dy);
    const float aw = a * w;
    const float cap = be_sqrt(fmaf(d, d, aw * aw));
    return (a < 0.0f) ? ((d < 0.0f) ? -cap : cap) : d;
}

// Signed distances of one pixel to the two wedges (:80-84).
BE_HD void be_pixel_dists(const BePatch& P, float X, float Y, float w, float* d1, float* d2) {
    {
        const float dx = X - P.vx[0], dy = Y - P.vy[0];
        const float DA = be_edge(dx, dy, P.sn[0], P.cs[0], w), DB = be_edge(dx, dy, P.sn[1], P.cs[1], w);
        const float f = P.flip[0];
        const bool in = (f * DA > 0.0f) && (f * DB < 0.0f);               // strict
        *d1 = fminf(fabsf(DA), fabsf(DB)) * (in ? f : -f);
    }
    {
        const float dx = X - P.vx[1], dy = Y - P.vy[1];
        const float DA = be_edge(dx, dy, P.sn[2], P.cs[2], w), DB = be_edge(dx, dy, P.sn[3], P.cs[3], w);
        const float f = P.flip[1];
        const bool in = (f * DA >= 0.0f) && (f * DB <= 0.0f);             // non-strict
        *d2 = fminf(fabsf(DA), fabsf(DB)) * (in ? f : -f);
    }
}

// h = 0.5 * (1 + erf(dist * inv_eta))   (:92).  Single-branch erf: log2(erfc(a)) ~ a*poly8(a) on [0,4] (weighted
// minimax fit, |erf error| < 1e-7 in fp32 including rounding; erf(a>=4) rounds to 1 in fp32), one MUFU.EX2.
BE_HD float be_h(float dist, float inv_eta) {
    const float t = dist * inv_eta;
    const float a = fminf(fabsf(t), 4.0f);
    float p = -4.5357578e-05f;
    p = fmaf(p, a, 4.4549927e-04f);
    p = fmaf(p, a, -1.4894147e-03f);
    p = fmaf(p, a, -7.7467301e-04f);
    p = fmaf(p, a, 2.8253718e-02f);
    p = fmaf(p, a, -1.4848163e-01f);
    p = fmaf(p, a, -9.1841639e-01f);
    p = fmaf(p, a, -1.6279086e+00f);
    const float r = 0.5f - be_exp2(fmaf(p, a, -1.0f));   // 0.5 * erf(a)
    return 0.5f + copysignf(r, t);
}

// (u0,u1,u2) from (h1,h2)   (:93-95)
BE_HD void be_wedges(float h1, float h2, float* u) {
    const float g = 1.0f - h2;
    u[0] = (1.0f - h1) * g; u[1] = h1 * g; u[2] = h2;
}

// exp(-x^2/delta^2) (:97-98) as 2^(-x^2 * log2(e)/delta^2)
BE_HD float be_bump(float x) { return be_exp2(-(x * x) * (1.44269504f / (BE_DELTA * BE_DELTA))); }

// boundary map value (blurry_edges_test.py:59-61)
BE_HD float be_boundary(float d1, float d2) {
    const float a1 = fabsf(d1), a2 = fabsf(d2);
    const float dB = (d2 >= 0.0f) ? d2 : ((a1 < a2) ? a1 : a2);
    return be_bump(dB);
}

// depth mask in {0,1,2} (blurry_edges_test.py:47-54); densify_w = the `--densify w` rule
BE_HD int be_mask(float d1, float d2, bool densify_w) {
    if (densify_w) return (d2 > 0.0f) ? 2 : ((d1 > 0.0f) ? 1 : 0);
    // normalized_gaussian(d) > 0.5  <=>  d^2 < delta^2 ln 2
    const float T = BE_DELTA * BE_DELTA * 0.693147181f;
    const int m = (d1 * d1 < T) ? 1 : 0;
    const int t = (d2 * d2 < T) ? 2 : 0;
    return (t == 2 || d2 >= 0.0f) ? t : m;
}

// ------------------------------------------------------------------------------------------
// Ridge regression (A^T A + lam I) C = A^T y for 3 wedges x 3 channels, solved in fp64 from fp32 sums.
// S[0..5] = sum u_i u_j for (00,01,02,11,12,22); S[6 + 3*w + c] = sum u_w y_c.
// Minv (6, same packing) is returned for the backward pass; C[3*w + c].
BE_HD void be_solve_colors(const float* S, float lam, double* Minv, float* C) {
    const double a = (double)S[0] + (double)lam, b = S[1], c = S[2];
    const double d = (double)S[3] + (double)lam, e = S[4], f = (double)S[5] + (double)lam;
    const double A00 = d * f - e * e, A01 = c * e - b * f, A02 = b * e - c * d;
    const double A11 = a * f - c * c, A12 = b * c - a * e, A22 = a * d - b * b;
    const double det = a * A00 + b * A01 + c * A02;     // >= lam^3 > 0
#if defined(__CUDA_ARCH__)
    // 1/det: MUFU.RCP seed + two Newton steps in fp64 (2^-23 -> 2^-46 -> 2^-92) instead of the ~30-instruction IEEE division;
    // this chain is on the solver warp's critical path, which gates the render warps
    double idet = (double)be_rcp((float)det);
    idet = idet * (2.0 - det * idet);
    idet = idet * (2.0 - det * idet);
#else
    const double idet = 1.0 / det;
#endif
    Minv[0] = A00 * idet; Minv[1] = A01 * idet; Minv[2] = A02 * idet;
    Minv[3] = A11 * idet; Minv[4] = A12 * idet; Minv[5] = A22 * idet;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const double b0 = S[6 + ch], b1 = S[9 + ch], b2 = S[12 + ch];
        C[0 + ch] = (float)(Minv[0] * b0 + Minv[1] * b1 + Minv[2] * b2);
        C[3 + ch] = (float)(Minv[1] * b0 + Minv[3] * b1 + Minv[4] * b2);
        C[6 + ch] = (float)(Minv[2] * b0 + Minv[4] * b1 + Minv[5] * b2);
    }
}

// ==========================================================================================
// Backward pieces (closed forms of SURVEY.md section 7; every folded target of the loss is detached in the reference,
// global_training.py:95,101,106, so the backward pass is purely per patch).
// ==========================================================================================
#define BE_INV_SQRT_PI 0.564189584f
#define BE_LN10 2.30258509f

struct BePatchGrad {      // per-patch chain-rule scalars
    float deta_dcoef[4];  // d eta_q / d coef_q   (:88-89)
    float dz_deta[4];     // dz1/deta0, dz1/deta2, dz2/deta1, dz2/deta3   (utils/depth_etas.py:23-34)
    float xy_scale, ang_scale;   // d(restored)/d(raw): 3 and pi for RAW12 (global_training.py:142-143), 1 otherwise
};

// d h / d dist and d h / d eta for h = 0.5 (1 + erf(dist * inv_eta)), inv_eta = 1/(sqrt2 eta)
BE_HD void be_h_grad(float dist, float inv_eta, float* dh_dd, float* dh_deta) {
    const float t = dist * inv_eta;
    const float e = BE_INV_SQRT_PI * be_exp2(-(t * t) * 1.44269504f);
    *dh_dd = e * inv_eta;
    *dh_deta = -e * t * (BE_SQRT2_F * inv_eta);     // 1/eta = sqrt2 * inv_eta
}

BE_HD void be_depth_grad(const BeCam& cam, float e1, float e2, float* dz1, float* dz2) {
    const float c = cam.intercept;
    const float c1 = -cam.sin_w * e1 + cam.cos_w * (e2 - c);
    const float c2 = -cam.sin_m * (e1 - c) + cam.cos_m * e2;
    const float c3 = -cam.sin_w * (e1 - c) + cam.cos_w * e2;
    const float half = (e1 + e2 - c) * 0.5f;
    float a, b, a1, a2, b1, b2;     // a,b and their partials wrt e1,e2
    if (c1 > 0.0f) { a = half; b = c + half; a1 = a2 = b1 = b2 = 0.5f; }
    else if (c2 > 0.0f) { a = c + (e1 - e2 - c) * 0.5f; b = (e2 - e1 + c) * 0.5f; a1 = 0.5f; a2 = -0.5f; b1 = -0.5f; b2 = 0.5f; }
    else if (c3 < 0.0f) { a = c + half; b = half; a1 = a2 = b1 = b2 = 0.5f; }
    else { a = e1; b = e2; a1 = 1.0f; a2 = 0.0f; b1 = 0.0f; b2 = 1.0f; }
    const float den = cam.k_fac * (a * a - b * b) + cam.k_const;
    const float f = -cam.numerator * cam.k_fac * 2.0f / (den * den);     // dz/d(a^2-b^2) * 2
    *dz1 = f * (a * a1 - b * b1);
    *dz2 = f * (a * a2 - b * b2);
}

BE_HD void be_patch_grad_setup(const float* p, int mode, const BeCam& cam, const BePatch& P, BePatchGrad& G) {
    const bool raw12 = (mode == BE_PARAMS_RAW12);
    const int neta = (mode == BE_PARAMS_LOCAL10 || mode == BE_PARAMS_LOCALRAW10) ? 2 : 4;
    for (int k = 0; k < 4; ++k) {
        if (k < neta) {
            const float e = raw12 ? p[8 + k] + 0.5f : p[8 + k];
            G.deta_dcoef[k] = P.eta[k] * (BE_LN10 * 4.0f * BE_INV_SQRT_PI) * be_exp2(-(e * e) * 1.44269504f);
        } else G.deta_dcoef[k] = 0.0f;
    }
    if (neta == 4) {
        be_depth_grad(cam, P.eta[0], P.eta[2], &G.dz_deta[0], &G.dz_deta[1]);
        be_depth_grad(cam, P.eta[1], P.eta[3], &G.dz_deta[2], &G.dz_deta[3]);
    } else { G.dz_deta[0] = G.dz_deta[1] = G.dz_deta[2] = G.dz_deta[3] = 0.0f; }
    G.xy_scale = raw12 ? 3.0f : 1.0f;
    G.ang_scale = raw12 ? BE_PI_F : 1.0f;
}

// Gradient of the signed distance of wedge k at one pixel w.r.t. (vertex x, vertex y, angle of edge A, angle of edge B),
// scaled by g = dL/d dist_k and ADDED into acc[0..3].  Mirrors autograd through :52-84: min() routes to the smaller
// |D| (ties split 1/2), abs() has sign(0) = 0, the cap branch is d/dd = d/r, d/da = w^2 a/r.
BE_HD void be_wedge_backward(const BePatch& P, int k, float X, float Y, float w, float g, float* acc) {
    const float dx = X - P.vx[k], dy = Y - P.vy[k];
    const float f = P.flip[k];
    const float snA = P.sn[2 * k], csA = P.cs[2 * k], snB = P.sn[2 * k + 1], csB = P.cs[2 * k + 1];
    const float dA = fmaf(csA, dy, -snA * dx), aA = fmaf(csA, dx, snA * dy);
    const float dB = fmaf(csB, dy, -snB * dx), aB = fmaf(csB, dx, snB * dy);
    const float capA = be_sqrt(fmaf(dA, dA, (aA * w) * (aA * w))), capB = be_sqrt(fmaf(dB, dB, (aB * w) * (aB * w)));
    const float DA = (aA < 0.0f) ? ((dA < 0.0f) ? -capA : capA) : dA;
    const float DB = (aB < 0.0f) ? ((dB < 0.0f) ? -capB : capB) : dB;
    const bool in = (k == 0) ? ((f * DA > 0.0f) && (f * DB < 0.0f)) : ((f * DA >= 0.0f) && (f * DB <= 0.0f));
    const float sg = g * (in ? f : -f);
    const float absA = fabsf(DA), absB = fabsf(DB);
    const float wA = (absA < absB) ? 1.0f : ((absA == absB) ? 0.5f : 0.0f);
    const float wB = 1.0f - wA;
    // edge A
    float gdA, gaA, gdB, gaB;
    if (aA < 0.0f) {
        const float ir = (absA > 0.0f) ? 1.0f / absA : 0.0f;             // |D| = r on the cap branch
        gdA = sg * wA * dA * ir; gaA = sg * wA * (w * w) * aA * ir;
    } else {
        gdA = sg * wA * ((dA > 0.0f) ? 1.0f : ((dA < 0.0f) ? -1.0f : 0.0f)); gaA = 0.0f;
    }
    if (aB < 0.0f) {
        const float ir = (absB > 0.0f) ? 1.0f / absB : 0.0f;
        gdB = sg * wB * dB * ir; gaB = sg * wB * (w * w) * aB * ir;
    } else {
        gdB = sg * wB * ((dB > 0.0f) ? 1.0f : ((dB < 0.0f) ? -1.0f : 0.0f)); gaB = 0.0f;
    }
    acc[0] += (gdA * snA - gaA * csA) + (gdB * snB - gaB * csB);
    acc[1] += (-gdA * csA - gaA * snA) + (-gdB * csB - gaB * snB);
    acc[2] += -gdA * aA + gaA * dA;
    acc[3] += -gdB * aB + gaB * dB;
}

// boundary map backward (blurry_edges_test.py:59-61): glb = dL/d lb -> adds to gd1, gd2
BE_HD void be_boundary_backward(float d1, float d2, float lb, float glb, float* gd1, float* gd2) {
    const float a1 = fabsf(d1), a2 = fabsf(d2);
    const float k = -2.0f / (BE_DELTA * BE_DELTA);
    if (d2 >= 0.0f) { *gd2 += glb * lb * k * d2; }
    else if (a1 < a2) { *gd1 += glb * lb * k * a1 * ((d1 > 0.0f) ? 1.0f : ((d1 < 0.0f) ? -1.0f : 0.0f)); }
    else { *gd2 += glb * lb * k * a2 * ((d2 > 0.0f) ? 1.0f : ((d2 < 0.0f) ? -1.0f : 0.0f)); }
}

// wedge values backward: gu = dL/du (3) -> dL/dh1, dL/dh2  (:93-95)
BE_HD void be_wedges_backward(float h1, float h2, const float* gu, float* gh1, float* gh2) {
    *gh1 = (1.0f - h2) * (gu[1] - gu[0]);
    *gh2 = gu[2] - (1.0f - h1) * gu[0] - h1 * gu[1];
}

// Second solve of the ridge backward: V = Minv (A^T G), Ssym = V C^T + C V^T (packed 00,01,02,11,12,22).
// AtG[3*w + c], C[3*w + c], V[3*w + c].
BE_HD void be_backsolve(const double* Minv, const float* AtG, const float* C, float* V, float* Ssym) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double b0 = AtG[c], b1 = AtG[3 + c], b2 = AtG[6 + c];
        V[0 + c] = (float)(Minv[0] * b0 + Minv[1] * b1 + Minv[2] * b2);
        V[3 + c] = (float)(Minv[1] * b0 + Minv[3] * b1 + Minv[4] * b2);
        V[6 + c] = (float)(Minv[2] * b0 + Minv[4] * b1 + Minv[5] * b2);
    }
    const int pi[6] = {0, 0, 0, 1, 1, 2}, pj[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) s += V[3 * pi[q] + c] * C[3 * pj[q] + c] + C[3 * pi[q] + c] * V[3 * pj[q] + c];
        Ssym[q] = s;
    }
}

// dL/du for one pixel of one image: G (3 channels) = dL/dP, y (3) = regression pixel, u (3) = wedges.
BE_HD void be_ridge_backward_pixel(const float* G, const float* y, const float* u, const float* C, const float* V,
                                   const float* Ssym, float* gu) {
    const float S[9] = {Ssym[0], Ssym[1], Ssym[2], Ssym[1], Ssym[3], Ssym[4], Ssym[2], Ssym[4], Ssym[5]};
#pragma unroll
    for (int w = 0; w < 3; ++w) {
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) s += G[c] * C[3 * w + c] + y[c] * V[3 * w + c];
        gu[w] = s - (S[3 * w] * u[0] + S[3 * w + 1] * u[1] + S[3 * w + 2] * u[2]);
    }
}

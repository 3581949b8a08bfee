// TEST INFRASTRUCTURE (oracle side): the device arithmetic of blurry_edges_b200/csrc/be_math.cuh compiled for the
// HOST and driven by plain loops, so that the CPU test-suite can check the exact fp32 formulas the kernels use
// against oracle/be_oracle.py without a GPU, and so that bench.py has a multi-threaded C port of the path to time as
// a CPU baseline ("kind": "port").  Nothing in blurry_edges_b200/ links or loads this file.
//
// Restates: blurry_edges_test.py:19-100 (pass A / pass B), utils/postprocessing_loss.py:151-173 (folds).
// Build: oracle/build_oracle.py  ->  oracle/_build/libbe_hostmath.so
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../blurry_edges_b200/csrc/be_math.cuh"
#include "../blurry_edges_b200/csrc/be_pack.cuh"

namespace {

struct Img {
    const float* p;
    int64_t sb, sm, sc, sy, sx;
    float at(int b, int m, int c, int y, int x) const { return p[b * sb + m * sm + c * sc + y * sy + x * sx]; }
};

int cover_1d(int y, int R, int s, int np) {
    const int hi = (y / s < np - 1) ? y / s : np - 1;
    const int lo = (y - R + 1 <= 0) ? 0 : (y - R + s) / s;
    return hi - lo + 1 > 0 ? hi - lo + 1 : 0;
}

}  // namespace

extern "C" {

// cam9 = numerator, k_fac, k_const, k_root, intercept, s, rho_prime  (sin/cos of the fp32 angles derived here)
static BeCam make_cam(const float* c) {
    BeCam cam;
    cam.numerator = c[0]; cam.k_fac = c[1]; cam.k_const = c[2]; cam.k_root = c[3]; cam.intercept = c[4];
    cam.s = c[5]; cam.rho_prime = c[6];
    const float tw = (float)(M_PI / 4.0), tm = (float)(3.0 * M_PI / 4.0);
    cam.sin_w = sinf(tw); cam.cos_w = cosf(tw); cam.sin_m = sinf(tm); cam.cos_m = cosf(tm);
    return cam;
}

// Pass A: est [M,L,10] -> colours [M,3(c),3(w),Hp,Wp]
int behm_colors(const float* est, int param_mode, const float* img, const int64_t* strides, int M, int H, int W, int R,
                int stride, float w, float lam, const float* cam7, float* colors) {
    const int Hp = (H - R) / stride + 1, Wp = (W - R) / stride + 1;
    const Img im{img, strides[0], strides[1], strides[2], strides[3], strides[4]};
    const BeCam cam = make_cam(cam7);
#pragma omp parallel for schedule(static)
    for (int n = 0; n < M * Hp * Wp; ++n) {
        const int b = n / (Hp * Wp), py = (n / Wp) % Hp, px = n % Wp;
        BePatch P;
        be_patch_setup(est + (size_t)n * 10, param_mode, cam, P);
        float S[16] = {0};
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < R; ++j) {
                float d1, d2, u[3];
                be_pixel_dists(P, be_axis(j, R), be_axis(i, R), w, &d1, &d2);
                be_wedges(be_h(d1, P.inv_eta[0]), be_h(d2, P.inv_eta[1]), u);
                S[0] += u[0] * u[0]; S[1] += u[0] * u[1]; S[2] += u[0] * u[2];
                S[3] += u[1] * u[1]; S[4] += u[1] * u[2]; S[5] += u[2] * u[2];
                for (int wd = 0; wd < 3; ++wd)
                    for (int c = 0; c < 3; ++c) S[6 + 3 * wd + c] += u[wd] * im.at(b, 0, c, py * stride + i, px * stride + j);
            }
        double Minv[6];
        float C[9];
        be_solve_colors(S, lam, Minv, C);
        for (int wd = 0; wd < 3; ++wd)
            for (int c = 0; c < 3; ++c) colors[(((size_t)b * 3 + c) * 3 + wd) * Hp * Wp + (size_t)py * Wp + px] = C[3 * wd + c];
    }
    return 0;
}

// Pass B for B pairs: outputs image [B,2,3,H,W], sharp [B,3,H,W], refoc [B,3,H,W], bndry [B,1,H,W], depth, conf [B,H,W]
int behm_render_fold(const float* est, int param_mode, const float* img, const int64_t* strides, int B, int H, int W, int R,
                     int stride, float w, float lam, const float* cam7, int densify_w, float* image, float* sharp,
                     float* refoc, float* bndry, float* depth, float* conf) {
    const int Hp = (H - R) / stride + 1, Wp = (W - R) / stride + 1, RR = R * R;
    const size_t HW = (size_t)H * W;
    const Img im{img, strides[0], strides[1], strides[2], strides[3], strides[4]};
    const BeCam cam = make_cam(cam7);
    const float inv_sharp = 1.0f / (BE_SQRT2_F * BE_ETA_SHARP);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        std::vector<float> acc(HW * 15, 0.0f), d1(RR), d2(RR), hh(RR * 4);
        for (int py = 0; py < Hp; ++py)
            for (int px = 0; px < Wp; ++px) {
                const size_t n = ((size_t)b * Hp + py) * Wp + px;
                BePatch P;
                be_patch_setup(est + n * 12, param_mode, cam, P);
                float S[16] = {0};
                int cnt1 = 0, cnt2 = 0;
                for (int i = 0; i < R; ++i)
                    for (int j = 0; j < R; ++j) {
                        const int q = i * R + j;
                        be_pixel_dists(P, be_axis(j, R), be_axis(i, R), w, &d1[q], &d2[q]);
                        for (int m = 0; m < 2; ++m) {
                            float u[3];
                            const float h1 = be_h(d1[q], P.inv_eta[2 * m]), h2 = be_h(d2[q], P.inv_eta[2 * m + 1]);
                            hh[q * 4 + 2 * m] = h1; hh[q * 4 + 2 * m + 1] = h2;
                            be_wedges(h1, h2, u);
                            S[0] += u[0] * u[0]; S[1] += u[0] * u[1]; S[2] += u[0] * u[2];
                            S[3] += u[1] * u[1]; S[4] += u[1] * u[2]; S[5] += u[2] * u[2];
                            for (int wd = 0; wd < 3; ++wd)
                                for (int c = 0; c < 3; ++c)
                                    S[6 + 3 * wd + c] += u[wd] * im.at(b, m, c, py * stride + i, px * stride + j);
                        }
                        const int mk = be_mask(d1[q], d2[q], densify_w != 0);
                        cnt1 += (mk == 1); cnt2 += (mk == 2);
                    }
                double Minv[6];
                float C[9];
                be_solve_colors(S, lam, Minv, C);
                const float ir1 = 1.0f / (BE_SQRT2_F * (cnt1 > 0 ? be_refocus_sigma(cam, P.z[0]) : BE_ETA_SHARP));
                const float ir2 = 1.0f / (BE_SQRT2_F * (cnt2 > 0 ? be_refocus_sigma(cam, P.z[1]) : BE_ETA_SHARP));
                for (int i = 0; i < R; ++i)
                    for (int j = 0; j < R; ++j) {
                        const int q = i * R + j;
                        float* a = &acc[((size_t)(py * stride + i) * W + px * stride + j) * 15];
                        float u[3];
                        for (int m = 0; m < 2; ++m) {
                            be_wedges(hh[q * 4 + 2 * m], hh[q * 4 + 2 * m + 1], u);
                            for (int c = 0; c < 3; ++c) a[3 * m + c] += fmaf(u[0], C[c], fmaf(u[1], C[3 + c], u[2] * C[6 + c]));
                        }
                        be_wedges(be_h(d1[q], inv_sharp), be_h(d2[q], inv_sharp), u);
                        for (int c = 0; c < 3; ++c) a[6 + c] += fmaf(u[0], C[c], fmaf(u[1], C[3 + c], u[2] * C[6 + c]));
                        be_wedges(be_h(d1[q], ir1), be_h(d2[q], ir2), u);
                        for (int c = 0; c < 3; ++c) a[9 + c] += fmaf(u[0], C[c], fmaf(u[1], C[3 + c], u[2] * C[6 + c]));
                        a[12] += be_boundary(d1[q], d2[q]);
                        const int mk = be_mask(d1[q], d2[q], densify_w != 0);
                        a[13] += (mk == 1) ? P.z[0] : ((mk == 2) ? P.z[1] : 0.0f);
                        a[14] += (mk > 0) ? 1.0f : 0.0f;
                    }
            }
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const size_t p = (size_t)y * W + x;
                const float* a = &acc[p * 15];
                const float n = (float)(cover_1d(y, R, stride, Hp) * cover_1d(x, R, stride, Wp));
                for (int c = 0; c < 6; ++c) image[((size_t)b * 6 + c) * HW + p] = a[c] / n;
                for (int c = 0; c < 3; ++c) sharp[((size_t)b * 3 + c) * HW + p] = a[6 + c] / n;
                for (int c = 0; c < 3; ++c) refoc[((size_t)b * 3 + c) * HW + p] = a[9 + c] / n;
                bndry[(size_t)b * HW + p] = a[12] / n;
                depth[(size_t)b * HW + p] = a[13] / (a[14] > 0.0f ? a[14] : 1.0f);
                conf[(size_t)b * HW + p] = a[14] / n;
            }
    }
    return 0;
}


// ------------------------------------------------------------------------------------------------------------------
// Global-stage loss, forward + analytic backward (global_training.py:62-157), the algorithm of the CUDA loss kernel
// written as plain loops.  Images are dataset-native channels-last [B,2,H,W,3]; deri [B,2,H-2,W-2,3].
// terms7 = the seven UNWEIGHTED terms (means / masked mean), loss = sum gamma_k terms_k, grad = dloss/draw [B,L,12].
// mode_local != 0: the local-stage loss (local_training.py:32-52): raw [B,10], images [B,R,R,3], one patch per sample,
// terms = (colour, boundary localisation, smoothness), gammas = (1, beta_loc, beta_smth).
// ------------------------------------------------------------------------------------------------------------------
struct PatchWork {
    std::vector<float> d1, d2, h, P, lb, G, gx, gy;
    std::vector<int> mk;
    explicit PatchWork(int RR) : d1(RR), d2(RR), h(RR * 4), P(RR * 6), lb(RR), G(RR * 6), gx(RR * 6), gy(RR * 6), mk(RR) {}
};

static void sobel_at(const float* p, int stride_row, float* sx, float* sy) {   // p points at the centre pixel
    const float a = p[-stride_row - 1], b = p[-stride_row], c = p[-stride_row + 1];
    const float d = p[-1], f = p[1];
    const float g = p[stride_row - 1], h = p[stride_row], i = p[stride_row + 1];
    *sx = (c - a) + 2.0f * (f - d) + (i - g);          // sobel_x = [[-1,0,1],[-2,0,2],[-1,0,1]]  (postprocessing_loss.py:19)
    *sy = (a + 2.0f * b + c) - (g + 2.0f * h + i);     // sobel_y = [[1,2,1],[0,0,0],[-1,-2,-1]]   (:20)
}

int behm_global_loss(const float* raw, const float* img_ny, const float* img_gt, const float* bndry_dist, const float* deri,
                     const float* bndry_depth, const float* gammas, int B, int H, int W, int R, int stride, float w, float lam,
                     const float* cam7, int mode_local, float* terms, float* loss, float* grad, float* gimg, float* gbnd) {
    const int Hp = (H - R) / stride + 1, Wp = (W - R) / stride + 1, RR = R * R, L = Hp * Wp, Ri = R - 2;
    const size_t HW = (size_t)H * W;
    const BeCam cam = make_cam(cam7);
    const int nimg = mode_local ? 1 : 2;
    const int np = mode_local ? 10 : 12;
    const int pmode = mode_local ? BE_PARAMS_LOCALRAW10 : BE_PARAMS_RAW12;
    auto ny = [&](int b, int m, int y, int x, int c) { return img_ny[((((size_t)b * nimg + m) * H + y) * W + x) * 3 + c]; };
    auto gt = [&](int b, int m, int y, int x, int c) { return img_gt[((((size_t)b * nimg + m) * H + y) * W + x) * 3 + c]; };

    // ---- pass 1: render + fold -> global image / boundary, mask count -------------------------------------------
    std::vector<double> acc(mode_local ? 0 : (size_t)B * HW * 7, 0.0);
    double msum = 0.0;
    if (!mode_local) {
        for (int b = 0; b < B; ++b)
            for (int py = 0; py < Hp; ++py)
                for (int px = 0; px < Wp; ++px) {
                    BePatch P;
                    be_patch_setup(raw + (((size_t)b * Hp + py) * Wp + px) * 12, pmode, cam, P);
                    PatchWork wk(RR);
                    float S[16] = {0};
                    for (int q = 0; q < RR; ++q) {
                        const int i = q / R, j = q % R;
                        be_pixel_dists(P, be_axis(j, R), be_axis(i, R), w, &wk.d1[q], &wk.d2[q]);
                        for (int m = 0; m < 2; ++m) {
                            float u[3];
                            wk.h[q * 4 + 2 * m] = be_h(wk.d1[q], P.inv_eta[2 * m]);
                            wk.h[q * 4 + 2 * m + 1] = be_h(wk.d2[q], P.inv_eta[2 * m + 1]);
                            be_wedges(wk.h[q * 4 + 2 * m], wk.h[q * 4 + 2 * m + 1], u);
                            S[0] += u[0] * u[0]; S[1] += u[0] * u[1]; S[2] += u[0] * u[2];
                            S[3] += u[1] * u[1]; S[4] += u[1] * u[2]; S[5] += u[2] * u[2];
                            for (int wd = 0; wd < 3; ++wd)
                                for (int c = 0; c < 3; ++c) S[6 + 3 * wd + c] += u[wd] * ny(b, m, py * stride + i, px * stride + j, c);
                        }
                    }
                    double Minv[6];
                    float C[9];
                    be_solve_colors(S, lam, Minv, C);
                    for (int q = 0; q < RR; ++q) {
                        const int y = py * stride + q / R, x = px * stride + q % R;
                        double* a = &acc[((size_t)b * HW + (size_t)y * W + x) * 7];
                        for (int m = 0; m < 2; ++m) {
                            float u[3];
                            be_wedges(wk.h[q * 4 + 2 * m], wk.h[q * 4 + 2 * m + 1], u);
                            for (int c = 0; c < 3; ++c) a[3 * m + c] += fmaf(u[0], C[c], fmaf(u[1], C[3 + c], u[2] * C[6 + c]));
                        }
                        a[6] += be_boundary(wk.d1[q], wk.d2[q]);
                        const int mk = be_mask(wk.d1[q], wk.d2[q], false);
                        if (mk != 0 && bndry_depth[(size_t)b * HW + (size_t)y * W + x] != 0.0f) msum += 1.0;
                    }
                }
        for (int b = 0; b < B; ++b)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const size_t p = (size_t)y * W + x;
                    const float n = (float)(cover_1d(y, R, stride, Hp) * cover_1d(x, R, stride, Wp));
                    for (int c = 0; c < 6; ++c) gimg[((size_t)b * 6 + c) * HW + p] = (float)acc[((size_t)b * HW + p) * 7 + c] / n;
                    gbnd[(size_t)b * HW + p] = (float)acc[((size_t)b * HW + p) * 7 + 6] / n;
                }
    }

    // ---- pass 2: per patch loss terms + backward ------------------------------------------------------------------
    const double Np = (double)B * L;
    double T[7] = {0, 0, 0, 0, 0, 0, 0};
    float kc, kcc, kbc, ks, ksc, kbl, kd;
    if (mode_local) {   // loss = colour + beta_loc * loc + beta_smth * smth   (local_training.py:47-52)
        kc = (float)(gammas[0] / (RR * Np)); kcc = 0; kbc = 0; ks = (float)(gammas[2] / (Ri * Ri * Np)); ksc = 0;
        kbl = (float)(gammas[1] / (RR * Np)); kd = 0;
    } else {
        kc = (float)(gammas[0] / (2.0 * RR * Np)); kcc = (float)(gammas[1] / (2.0 * RR * Np)); kbc = (float)(gammas[2] / (RR * Np));
        ks = (float)(gammas[3] / (2.0 * Ri * Ri * Np)); ksc = (float)(gammas[4] / (2.0 * Ri * Ri * Np));
        kbl = (float)(gammas[5] / (RR * Np)); kd = (float)(gammas[6] / msum);
    }
    for (int n = 0; n < B * L; ++n) {
        const int b = n / L, py = (n / Wp) % Hp, px = n % Wp;
        const int y0 = py * stride, x0 = px * stride;
        BePatch P;
        BePatchGrad PG;
        be_patch_setup(raw + (size_t)n * np, pmode, cam, P);
        be_patch_grad_setup(raw + (size_t)n * np, pmode, cam, P, PG);
        PatchWork wk(RR);
        float S[16] = {0};
        for (int q = 0; q < RR; ++q) {
            const int i = q / R, j = q % R;
            be_pixel_dists(P, be_axis(j, R), be_axis(i, R), w, &wk.d1[q], &wk.d2[q]);
            for (int m = 0; m < nimg; ++m) {
                float u[3];
                wk.h[q * 4 + 2 * m] = be_h(wk.d1[q], P.inv_eta[2 * m]);
                wk.h[q * 4 + 2 * m + 1] = be_h(wk.d2[q], P.inv_eta[2 * m + 1]);
                be_wedges(wk.h[q * 4 + 2 * m], wk.h[q * 4 + 2 * m + 1], u);
                S[0] += u[0] * u[0]; S[1] += u[0] * u[1]; S[2] += u[0] * u[2];
                S[3] += u[1] * u[1]; S[4] += u[1] * u[2]; S[5] += u[2] * u[2];
                for (int wd = 0; wd < 3; ++wd)
                    for (int c = 0; c < 3; ++c) S[6 + 3 * wd + c] += u[wd] * ny(b, m, y0 + i, x0 + j, c);
            }
        }
        double Minv[6];
        float C[9];
        be_solve_colors(S, lam, Minv, C);
        // render, direct gradient G = dL/dP
        for (int q = 0; q < RR; ++q) {
            const int y = y0 + q / R, x = x0 + q % R;
            for (int m = 0; m < nimg; ++m) {
                float u[3];
                be_wedges(wk.h[q * 4 + 2 * m], wk.h[q * 4 + 2 * m + 1], u);
                for (int c = 0; c < 3; ++c) {
                    const float Pv = fmaf(u[0], C[c], fmaf(u[1], C[3 + c], u[2] * C[6 + c]));
                    wk.P[(3 * m + c) * RR + q] = Pv;
                    const float e1 = Pv - gt(b, m, y, x, c);
                    T[0] += (double)e1 * e1;
                    float gval = 2.0f * kc * e1;
                    if (!mode_local) {
                        const float e2 = Pv - gimg[((size_t)b * 6 + 3 * m + c) * HW + (size_t)y * W + x];
                        T[1] += (double)e2 * e2;
                        gval += 2.0f * kcc * e2;
                    }
                    wk.G[(3 * m + c) * RR + q] = gval;
                }
            }
            wk.lb[q] = be_boundary(wk.d1[q], wk.d2[q]);
            wk.mk[q] = be_mask(wk.d1[q], wk.d2[q], false);
        }
        // smoothness: Sobel magnitude of the rendered patch vs targets, and its adjoint into G
        std::fill(wk.gx.begin(), wk.gx.end(), 0.0f);
        std::fill(wk.gy.begin(), wk.gy.end(), 0.0f);
        for (int mc = 0; mc < 3 * nimg; ++mc)
            for (int i = 1; i < R - 1; ++i)
                for (int j = 1; j < R - 1; ++j) {
                    float sx, sy;
                    sobel_at(&wk.P[mc * RR + i * R + j], R, &sx, &sy);
                    const float mag = sqrtf(sx * sx + sy * sy + 1e-8f);
                    const int m = mc / 3, c = mc % 3;
                    const float tg = deri[((((size_t)b * nimg + m) * (H - 2) + (y0 + i - 1)) * (W - 2) + (x0 + j - 1)) * 3 + c];
                    const float e1 = mag - tg;
                    T[3] += (double)e1 * e1;
                    float gm = 2.0f * ks * e1;
                    if (!mode_local) {
                        float gsx, gsy;
                        sobel_at(&gimg[((size_t)b * 6 + mc) * HW + (size_t)(y0 + i) * W + (x0 + j)], W, &gsx, &gsy);
                        const float e2 = mag - sqrtf(gsx * gsx + gsy * gsy + 1e-8f);
                        T[4] += (double)e2 * e2;
                        gm += 2.0f * ksc * e2;
                    }
                    wk.gx[mc * RR + i * R + j] = gm * sx / mag;
                    wk.gy[mc * RR + i * R + j] = gm * sy / mag;
                }
        static const float KX[3][3] = {{-1, 0, 1}, {-2, 0, 2}, {-1, 0, 1}}, KY[3][3] = {{1, 2, 1}, {0, 0, 0}, {-1, -2, -1}};
        for (int mc = 0; mc < 3 * nimg; ++mc)
            for (int i = 0; i < R; ++i)
                for (int j = 0; j < R; ++j) {
                    float s = 0.0f;   // adjoint: output (i',j') used input (i'+a-1, j'+b-1) with weight K[a][b]
                    for (int a = 0; a < 3; ++a)
                        for (int bb = 0; bb < 3; ++bb) {
                            const int io = i - a + 1, jo = j - bb + 1;
                            if (io < 1 || io > R - 2 || jo < 1 || jo > R - 2) continue;
                            s += wk.gx[mc * RR + io * R + jo] * KX[a][bb] + wk.gy[mc * RR + io * R + jo] * KY[a][bb];
                        }
                    wk.G[mc * RR + i * R + j] += s;
                }
        // A^T G, second solve
        float AtG[9] = {0}, V[9], Ssym[6];
        for (int q = 0; q < RR; ++q)
            for (int m = 0; m < nimg; ++m) {
                float u[3];
                be_wedges(wk.h[q * 4 + 2 * m], wk.h[q * 4 + 2 * m + 1], u);
                for (int wd = 0; wd < 3; ++wd)
                    for (int c = 0; c < 3; ++c) AtG[3 * wd + c] += u[wd] * wk.G[(3 * m + c) * RR + q];
            }
        be_backsolve(Minv, AtG, C, V, Ssym);
        // per pixel backward
        float geo[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}}, geta[4] = {0, 0, 0, 0}, gz[2] = {0, 0};
        for (int q = 0; q < RR; ++q) {
            const int i = q / R, j = q % R, y = y0 + i, x = x0 + j;
            float gd1 = 0.0f, gd2 = 0.0f;
            for (int m = 0; m < nimg; ++m) {
                float u[3], gu[3], Gp[3], yv[3], gh1, gh2, a, bq;
                const float h1 = wk.h[q * 4 + 2 * m], h2 = wk.h[q * 4 + 2 * m + 1];
                be_wedges(h1, h2, u);
                for (int c = 0; c < 3; ++c) { Gp[c] = wk.G[(3 * m + c) * RR + q]; yv[c] = ny(b, m, y, x, c); }
                be_ridge_backward_pixel(Gp, yv, u, C, V, Ssym, gu);
                be_wedges_backward(h1, h2, gu, &gh1, &gh2);
                be_h_grad(wk.d1[q], P.inv_eta[2 * m], &a, &bq);
                gd1 += gh1 * a; geta[2 * m] += gh1 * bq;
                be_h_grad(wk.d2[q], P.inv_eta[2 * m + 1], &a, &bq);
                gd2 += gh2 * a; geta[2 * m + 1] += gh2 * bq;
            }
            // boundary terms
            const float lb = wk.lb[q];
            float glb;
            if (mode_local) {
                const float bd = bndry_dist[(size_t)b * HW + (size_t)y * W + x];
                T[5] += (double)(bd * lb) * (bd * lb);
                glb = 2.0f * kbl * bd * bd * lb;
            } else {
                const float bd = log2f(bndry_dist[(size_t)b * HW + (size_t)y * W + x] + 1.0f);
                const float e = lb - gbnd[(size_t)b * HW + (size_t)y * W + x];
                T[2] += (double)e * e;
                T[5] += (double)(bd * lb) * (bd * lb);
                glb = 2.0f * kbc * e + 2.0f * kbl * bd * bd * lb;
                // depth term
                const float zg = bndry_depth[(size_t)b * HW + (size_t)y * W + x];
                const int mk = wk.mk[q];
                if (zg != 0.0f && mk != 0) {
                    const float e2 = P.z[mk - 1] - zg;
                    T[6] += (double)e2 * e2;
                    gz[mk - 1] += 2.0f * kd * e2;
                }
            }
            be_boundary_backward(wk.d1[q], wk.d2[q], lb, glb, &gd1, &gd2);
            be_wedge_backward(P, 0, be_axis(j, R), be_axis(i, R), w, gd1, geo[0]);
            be_wedge_backward(P, 1, be_axis(j, R), be_axis(i, R), w, gd2, geo[1]);
        }
        float* g = grad + (size_t)n * np;
        for (int k = 0; k < 2; ++k) {
            g[2 * k] = PG.xy_scale * geo[k][0];
            g[2 * k + 1] = PG.xy_scale * geo[k][1];
            g[4 + 2 * k] = PG.ang_scale * (geo[k][2] + geo[k][3]);
            g[5 + 2 * k] = PG.ang_scale * geo[k][3];
        }
        if (mode_local) {
            g[8] = geta[0] * PG.deta_dcoef[0];
            g[9] = geta[1] * PG.deta_dcoef[1];
        } else {
            const float ge[4] = {geta[0] + gz[0] * PG.dz_deta[0], geta[1] + gz[1] * PG.dz_deta[2],
                                 geta[2] + gz[0] * PG.dz_deta[1], geta[3] + gz[1] * PG.dz_deta[3]};
            for (int k = 0; k < 4; ++k) g[8 + k] = ge[k] * PG.deta_dcoef[k];
        }
    }
    if (mode_local) {
        terms[0] = (float)(T[0] / (RR * Np)); terms[1] = (float)(T[5] / (RR * Np)); terms[2] = (float)(T[3] / (Ri * Ri * Np));
        *loss = (float)(gammas[0] * terms[0] + gammas[1] * terms[1] + gammas[2] * terms[2]);
    } else {
        terms[0] = (float)(T[0] / (2.0 * RR * Np)); terms[1] = (float)(T[1] / (2.0 * RR * Np)); terms[2] = (float)(T[2] / (RR * Np));
        terms[3] = (float)(T[3] / (2.0 * Ri * Ri * Np)); terms[4] = (float)(T[4] / (2.0 * Ri * Ri * Np));
        terms[5] = (float)(T[5] / (RR * Np)); terms[6] = (float)(T[6] / msum);
        double l = 0;
        for (int k = 0; k < 7; ++k) l += (double)gammas[k] * terms[k];
        *loss = (float)l;
    }
    return 0;
}

// Packed (two-pixel) functions of be_pack.cuh against the scalar specification of be_math.cuh, on n random pixels of one
// patch.  out[0..3]: max |difference| of (d1, d2), h, boundary, mask weights (bit-identical arithmetic: expected 0);
// out[4]: max relative difference of the wedge backward sums (algebraically rewritten: d/|D| instead of sign(d)).
void hm_pack_selfcheck(const float* p12, const float* cam7, const float* xy, int n, float w, float* out5) {
    const BeCam cam = make_cam(cam7);
    BePatch P;
    be_patch_setup(p12, BE_PARAMS_RESTORED12, cam, P);
    for (int k = 0; k < 5; ++k) out5[k] = 0.0f;
    float acc_s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    f2 acc_p[8];
    for (int k = 0; k < 8; ++k) acc_p[k] = bc2(0.0f);
    for (int i = 0; i + 1 < n; i += 2) {
        const float X0 = xy[2 * i], Y0 = xy[2 * i + 1], X1 = xy[2 * i + 2], Y1 = xy[2 * i + 3];
        float a1, a2, b1, b2;
        be_pixel_dists(P, X0, Y0, w, &a1, &a2);
        be_pixel_dists(P, X1, Y1, w, &b1, &b2);
        f2 d1, d2;
        be_pixel_dists2(P, mk2(X0, X1), mk2(Y0, Y1), w, &d1, &d2);
        out5[0] = fmaxf(out5[0], fmaxf(fmaxf(fabsf(lo(d1) - a1), fabsf(hi(d1) - b1)), fmaxf(fabsf(lo(d2) - a2), fabsf(hi(d2) - b2))));
        for (int q = 0; q < 4; ++q) {
            const f2 h = be_h2((q & 1) ? d2 : d1, P.inv_eta[q]);
            out5[1] = fmaxf(out5[1], fmaxf(fabsf(lo(h) - be_h((q & 1) ? a2 : a1, P.inv_eta[q])), fabsf(hi(h) - be_h((q & 1) ? b2 : b1, P.inv_eta[q]))));
        }
        const f2 lb = be_boundary2(d1, d2);
        out5[2] = fmaxf(out5[2], fmaxf(fabsf(lo(lb) - be_boundary(a1, a2)), fabsf(hi(lb) - be_boundary(b1, b2))));
        for (int dw = 0; dw < 2; ++dw) {
            float m1, m2;
            be_mask_weights(a1, a2, dw != 0, &m1, &m2);
            const int mk = be_mask(a1, a2, dw != 0);
            out5[3] = fmaxf(out5[3], fabsf(m1 - (mk == 1 ? 1.0f : 0.0f)) + fabsf(m2 - (mk == 2 ? 1.0f : 0.0f)));
        }
        const float g0 = 0.3f + 0.01f * (float)(i % 7), g1 = -0.2f + 0.02f * (float)(i % 5);
        for (int k = 0; k < 2; ++k) {
            be_wedge_backward(P, k, X0, Y0, w, g0, &acc_s[4 * k]);
            be_wedge_backward(P, k, X1, Y1, w, g1, &acc_s[4 * k]);
            be_wedge_backward2(P, k, mk2(X0, X1), mk2(Y0, Y1), w, mk2(g0, g1), &acc_p[4 * k]);
        }
    }
    float num = 0.0f, den = 1e-30f;
    for (int k = 0; k < 8; ++k) { num = fmaxf(num, fabsf(lo(acc_p[k]) + hi(acc_p[k]) - acc_s[k])); den = fmaxf(den, fabsf(acc_s[k])); }
    out5[4] = num / den;
}

}  // extern "C"

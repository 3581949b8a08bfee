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

}  // extern "C"

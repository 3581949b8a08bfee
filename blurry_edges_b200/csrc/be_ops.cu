// Method-granularity kernels behind the reference's helper-class METHODS: the tensors keep the
// reference's own (unfolded) layouts, so these kernels are plain HBM-bound elementwise / gather / per-patch-reduce
// passes.  They exist so that subclasses written against PostProcess*Base / DepthEtas (the reference scripts) run on
// CUDA kernels of this library with autograd support; the fused classes never use them.
//
// Layout convention: `Lsp` = Hp*Wp for the global layout ([B,K,Hp,Wp], [B,2,R,R,Hp,Wp], ...) and 1 for the local
// layout ([B,K], [B,2,R,R], ...); the patch index l is always the fastest dimension.
#include "be_internal.h"

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ void load_geo(const float* __restrict__ params, int K, size_t b, size_t l, size_t Lsp, float* p8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) p8[k] = __ldg(params + (b * K + k) * Lsp + l);
}

// geometry-only patch setup (angles are used as given: utils/postprocessing_loss.py:43-55)
__device__ __forceinline__ void setup_geo(const float* p8, BePatch& P) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        P.vx[k] = p8[2 * k]; P.vy[k] = p8[2 * k + 1];
        const float th = p8[4 + 2 * k], ph = p8[5 + 2 * k];
        be_sincos(th, &P.sn[2 * k], &P.cs[2 * k]);
        be_sincos_sum(th, ph, &P.sn[2 * k + 1], &P.cs[2 * k + 1]);
        P.flip[k] = (be_wrap_2pi(ph) < BE_PI_F) ? 1.0f : -1.0f;
    }
}

// ---- params2dists (:43-86) ---------------------------------------------------------------------
// one thread per patch: the 4 sincos are evaluated once and the R*R pixel loop writes coalesced along l
__global__ void __launch_bounds__(TPB) k_params2dists(const float* __restrict__ params, int K, int B, size_t Lsp, int R, float w,
                                                      float* __restrict__ dists) {
    const size_t n = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (n >= (size_t)B * Lsp) return;
    const size_t b = n / Lsp, l = n % Lsp;
    float p8[8];
    load_geo(params, K, b, l, Lsp, p8);
    BePatch P;
    setup_geo(p8, P);
    for (int i = 0; i < R; ++i) {
        const float Y = be_axis(i, R);
        for (int j = 0; j < R; ++j) {
            float d1, d2;
            be_pixel_dists(P, be_axis(j, R), Y, w, &d1, &d2);
            dists[(((b * 2 + 0) * R + i) * R + j) * Lsp + l] = d1;
            dists[(((b * 2 + 1) * R + i) * R + j) * Lsp + l] = d2;
        }
    }
}

__global__ void __launch_bounds__(TPB) k_params2dists_bwd(const float* __restrict__ params, int K, const float* __restrict__ gd, int B,
                                                          size_t Lsp, int R, float w, float* __restrict__ gp) {
    const size_t n = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (n >= (size_t)B * Lsp) return;
    const size_t b = n / Lsp, l = n % Lsp;
    float p8[8];
    load_geo(params, K, b, l, Lsp, p8);
    BePatch P;
    setup_geo(p8, P);
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < R; ++i) {
        const float Y = be_axis(i, R);
        for (int j = 0; j < R; ++j) {
            const float X = be_axis(j, R);
            be_wedge_backward(P, 0, X, Y, w, __ldg(gd + (((b * 2 + 0) * R + i) * R + j) * Lsp + l), a0);
            be_wedge_backward(P, 1, X, Y, w, __ldg(gd + (((b * 2 + 1) * R + i) * R + j) * Lsp + l), a1);
        }
    }
    float* o = gp + b * 8 * Lsp + l;     // (x0,y0,x1,y1,theta1,phi1,theta2,phi2)
    o[0 * Lsp] = a0[0]; o[1 * Lsp] = a0[1]; o[2 * Lsp] = a1[0]; o[3 * Lsp] = a1[1];
    o[4 * Lsp] = a0[2] + a0[3]; o[5 * Lsp] = a0[3]; o[6 * Lsp] = a1[2] + a1[3]; o[7 * Lsp] = a1[3];
}

// ---- dists2indicators (:91-95) -----------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_indicators(const float* __restrict__ dists, const float* __restrict__ etas, int B, size_t Lsp,
                                                    int RR, float* __restrict__ wedges) {
    const size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;      // over (b, pixel, l)
    if (idx >= (size_t)B * RR * Lsp) return;
    const size_t l = idx % Lsp, q = (idx / Lsp) % RR, b = idx / (Lsp * RR);
    const float d1 = __ldg(dists + ((b * 2 + 0) * RR + q) * Lsp + l), d2 = __ldg(dists + ((b * 2 + 1) * RR + q) * Lsp + l);
    const float e1 = __ldg(etas + (b * 2 + 0) * Lsp + l), e2 = __ldg(etas + (b * 2 + 1) * Lsp + l);
    float u[3];
    be_wedges(be_h(d1, 1.0f / (BE_SQRT2_F * e1)), be_h(d2, 1.0f / (BE_SQRT2_F * e2)), u);
#pragma unroll
    for (int k = 0; k < 3; ++k) wedges[((b * 3 + k) * RR + q) * Lsp + l] = u[k];
}

// one thread per patch: grad wrt dists (elementwise) and wrt the 2 etas (sum over the patch)
__global__ void __launch_bounds__(TPB) k_indicators_bwd(const float* __restrict__ dists, const float* __restrict__ etas,
                                                        const float* __restrict__ gw, int B, size_t Lsp, int RR,
                                                        float* __restrict__ gd, float* __restrict__ ge) {
    const size_t n = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (n >= (size_t)B * Lsp) return;
    const size_t b = n / Lsp, l = n % Lsp;
    const float ie1 = 1.0f / (BE_SQRT2_F * __ldg(etas + (b * 2 + 0) * Lsp + l)), ie2 = 1.0f / (BE_SQRT2_F * __ldg(etas + (b * 2 + 1) * Lsp + l));
    float s1 = 0.0f, s2 = 0.0f;
    for (int q = 0; q < RR; ++q) {
        const float d1 = __ldg(dists + ((b * 2 + 0) * RR + q) * Lsp + l), d2 = __ldg(dists + ((b * 2 + 1) * RR + q) * Lsp + l);
        const float h1 = be_h(d1, ie1), h2 = be_h(d2, ie2);
        float gu[3], gh1, gh2, a, c;
#pragma unroll
        for (int k = 0; k < 3; ++k) gu[k] = __ldg(gw + ((b * 3 + k) * RR + q) * Lsp + l);
        be_wedges_backward(h1, h2, gu, &gh1, &gh2);
        be_h_grad(d1, ie1, &a, &c);
        gd[((b * 2 + 0) * RR + q) * Lsp + l] = gh1 * a; s1 = fmaf(gh1, c, s1);
        be_h_grad(d2, ie2, &a, &c);
        gd[((b * 2 + 1) * RR + q) * Lsp + l] = gh2 * a; s2 = fmaf(gh2, c, s2);
    }
    ge[(b * 2 + 0) * Lsp + l] = s1;
    ge[(b * 2 + 1) * Lsp + l] = s2;
}

// ---- elementwise family ------------------------------------------------------------------------
// op 0: params2etas (:88-89)   op 1: normalized_gaussian(x, delta=p0) (:97-98)   op 2: depth2sigma(depth, rho_prime=p0)
__global__ void __launch_bounds__(TPB) k_unary(int op, const float* __restrict__ x, float p0, BeCam cam, size_t n, float* __restrict__ y) {
    const size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    const float v = __ldg(x + i);
    float r;
    if (op == 0) r = be_eta(v);
    else if (op == 1) r = be_exp2(-(v * v) * (1.44269504f / (p0 * p0)));
    else r = fabsf((1.0f / v - p0) * cam.s + 1.0f) / cam.k_root;
    y[i] = r;
}

__global__ void __launch_bounds__(TPB) k_unary_bwd(int op, const float* __restrict__ x, const float* __restrict__ gy, float p0, BeCam cam,
                                                   size_t n, float* __restrict__ gx) {
    const size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    const float v = __ldg(x + i), g = __ldg(gy + i);
    float r;
    if (op == 0) r = g * be_eta(v) * (BE_LN10 * 4.0f * BE_INV_SQRT_PI) * be_exp2(-(v * v) * 1.44269504f);
    else if (op == 1) r = g * be_exp2(-(v * v) * (1.44269504f / (p0 * p0))) * (-2.0f * v / (p0 * p0));
    else {
        const float t = (1.0f / v - p0) * cam.s + 1.0f;
        const float sg = (t > 0.0f) ? 1.0f : ((t < 0.0f) ? -1.0f : 0.0f);
        r = g * sg * (-cam.s / (v * v)) / cam.k_root;
    }
    gx[i] = r;
}

// ---- Smish activation of LocalStage (models/local_stage.py:4-6): y = x tanh(log(1 + sigmoid(x))) ------------------------------
// With t = 1 + sigmoid(x): tanh(log t) = (t^2 - 1)/(t^2 + 1).  HBM bound (8 B/element forward, 12 B/element backward): 16-byte
// loads and stores, streaming cache hints, grid-stride over a grid sized to the SM count.
__device__ __forceinline__ float smish_f(float x, float* dfdx) {
    const float s = be_rcp(1.0f + be_exp2(-x * 1.44269504f));
    const float t = 1.0f + s, t2 = t * t;
    const float r = be_rcp(t2 + 1.0f);
    if (dfdx) *dfdx = 4.0f * t * r * r * s * (1.0f - s);
    return (t2 - 1.0f) * r;
}

__global__ void __launch_bounds__(TPB) k_smish(const float* __restrict__ x, size_t n, float* __restrict__ y, int vec4) {
    const size_t stride = (size_t)gridDim.x * TPB;
    if (vec4) {
        for (size_t i = (size_t)blockIdx.x * TPB + threadIdx.x; i < n / 4; i += stride) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i);
            float4 r;
            r.x = v.x * smish_f(v.x, nullptr); r.y = v.y * smish_f(v.y, nullptr);
            r.z = v.z * smish_f(v.z, nullptr); r.w = v.w * smish_f(v.w, nullptr);
            __stcs(reinterpret_cast<float4*>(y) + i, r);
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * TPB + threadIdx.x; i < n; i += stride) y[i] = x[i] * smish_f(x[i], nullptr);
    }
}

__global__ void __launch_bounds__(TPB) k_smish_bwd(const float* __restrict__ x, const float* __restrict__ gy, size_t n,
                                                   float* __restrict__ gx, int vec4) {
    const size_t stride = (size_t)gridDim.x * TPB;
    auto one = [](float v, float g) { float d; const float f = smish_f(v, &d); return g * fmaf(v, d, f); };
    if (vec4) {
        for (size_t i = (size_t)blockIdx.x * TPB + threadIdx.x; i < n / 4; i += stride) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i), g = __ldcs(reinterpret_cast<const float4*>(gy) + i);
            __stcs(reinterpret_cast<float4*>(gx) + i, make_float4(one(v.x, g.x), one(v.y, g.y), one(v.z, g.z), one(v.w, g.w)));
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * TPB + threadIdx.x; i < n; i += stride) gx[i] = one(x[i], gy[i]);
    }
}

__global__ void __launch_bounds__(TPB) k_depth(const float* __restrict__ e1, const float* __restrict__ e2, BeCam cam, size_t n,
                                               float* __restrict__ z) {
    const size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (i < n) z[i] = be_depth(cam, __ldg(e1 + i), __ldg(e2 + i));
}

__global__ void __launch_bounds__(TPB) k_depth_bwd(const float* __restrict__ e1, const float* __restrict__ e2, const float* __restrict__ gz,
                                                   BeCam cam, size_t n, float* __restrict__ g1, float* __restrict__ g2) {
    const size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    float d1, d2;
    be_depth_grad(cam, __ldg(e1 + i), __ldg(e2 + i), &d1, &d2);
    const float g = __ldg(gz + i);
    g1[i] = g * d1; g2[i] = g * d2;
}

// ---- inverse_3by3 (:104-112) --------------------------------------------------------------------
__device__ __forceinline__ void inv3(const double* m, double* r) {
    const double A = m[4] * m[8] - m[5] * m[7], Bc = -(m[3] * m[8] - m[5] * m[6]), Cc = m[3] * m[7] - m[4] * m[6];
    const double idet = 1.0 / (m[0] * A + m[1] * Bc + m[2] * Cc);
    r[0] = A * idet; r[1] = -(m[1] * m[8] - m[2] * m[7]) * idet; r[2] = (m[1] * m[5] - m[2] * m[4]) * idet;
    r[3] = Bc * idet; r[4] = (m[0] * m[8] - m[2] * m[6]) * idet; r[5] = -(m[0] * m[5] - m[2] * m[3]) * idet;
    r[6] = Cc * idet; r[7] = -(m[0] * m[7] - m[1] * m[6]) * idet; r[8] = (m[0] * m[4] - m[1] * m[3]) * idet;
}

__global__ void __launch_bounds__(TPB) k_inverse3(const float* __restrict__ A, size_t n, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    double m[9], r[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = (double)__ldg(A + i * 9 + k);
    inv3(m, r);
#pragma unroll
    for (int k = 0; k < 9; ++k) out[i * 9 + k] = (float)r[k];
}

// grad_A = -Inv^T G Inv^T
__global__ void __launch_bounds__(TPB) k_inverse3_bwd(const float* __restrict__ inv, const float* __restrict__ g, size_t n,
                                                      float* __restrict__ gA) {
    const size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    float I[9], G[9], T[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { I[k] = __ldg(inv + i * 9 + k); G[k] = __ldg(g + i * 9 + k); }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) T[3 * r + c] = I[r] * G[c] + I[3 + r] * G[3 + c] + I[6 + r] * G[6 + c];       // (Inv^T G)[r][c]
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) gA[i * 9 + 3 * r + c] = -(T[3 * r] * I[3 * c] + T[3 * r + 1] * I[3 * c + 1] + T[3 * r + 2] * I[3 * c + 2]);
}

// ---- get_image_derivative (:114-117): [N,H,W] planes -> [N,H-2,W-2] -----------------------------
__global__ void __launch_bounds__(TPB) k_sobel(const float* __restrict__ img, size_t N, int H, int W, float* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
    const int Ho = H - 2, Wo = W - 2;
    if (idx >= N * Ho * Wo) return;
    const int x = (int)(idx % Wo), y = (int)((idx / Wo) % Ho);
    const size_t pl = idx / ((size_t)Wo * Ho);
    const float* p = img + (pl * H + y + 1) * W + x + 1;
    const float a = p[-W - 1], b = p[-W], c = p[-W + 1], d = p[-1], f = p[1], g = p[W - 1], h = p[W], i = p[W + 1];
    const float sx = (c - a) + 2.0f * (f - d) + (i - g), sy = (a + 2.0f * b + c) - (g + 2.0f * h + i);
    out[idx] = sqrtf(sx * sx + sy * sy + 1e-8f);
}

__global__ void __launch_bounds__(TPB) k_sobel_bwd(const float* __restrict__ img, const float* __restrict__ gout, size_t N, int H, int W,
                                                   float* __restrict__ gimg) {
    const size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (idx >= N * H * W) return;
    const int x = (int)(idx % W), y = (int)((idx / W) % H);
    const size_t pl = idx / ((size_t)W * H);
    const int Ho = H - 2, Wo = W - 2;
    float s = 0.0f;
    for (int di = -1; di <= 1; ++di)
        for (int dj = -1; dj <= 1; ++dj) {
            if (di == 0 && dj == 0) continue;
            const int yo = y + di - 1, xo = x + dj - 1;          // output index whose centre is (y+di, x+dj)
            if (yo < 0 || yo >= Ho || xo < 0 || xo >= Wo) continue;
            const float* p = img + (pl * H + yo + 1) * W + xo + 1;
            const float a = p[-W - 1], b = p[-W], c = p[-W + 1], d = p[-1], f = p[1], g = p[W - 1], h = p[W], i = p[W + 1];
            const float sx = (c - a) + 2.0f * (f - d) + (i - g), sy = (a + 2.0f * b + c) - (g + 2.0f * h + i);
            const float gm = __ldg(gout + (pl * Ho + yo) * Wo + xo) * rsqrtf(sx * sx + sy * sy + 1e-8f);
            const float wx = (float)(-dj * ((di == 0) ? 2 : 1)), wy = (float)(di * ((dj == 0) ? 2 : 1));
            s += gm * (sx * wx + sy * wy);
        }
    gimg[idx] = s;
}

// ---- folds (:151-173) ---------------------------------------------------------------------------
__device__ __forceinline__ void cover(int y, int R, int s, int np, int* lo, int* hi) {
    *hi = min(y / s, np - 1);
    *lo = (y - R + 1 <= 0) ? 0 : (y - R + s) / s;
}

// patches [P,R,R,Hp,Wp] -> out [P,H,W] = overlap sum / num_patches (mode 0) or plain sum (mode 1)
__global__ void __launch_bounds__(TPB) k_fold(const float* __restrict__ patches, size_t P, BeGeom g, int mode, float* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (idx >= P * g.H * g.W) return;
    const int x = (int)(idx % g.W), y = (int)((idx / g.W) % g.H);
    const size_t pl = idx / ((size_t)g.W * g.H);
    int ylo, yhi, xlo, xhi;
    cover(y, g.R, g.stride, g.Hp, &ylo, &yhi);
    cover(x, g.R, g.stride, g.Wp, &xlo, &xhi);
    float s = 0.0f;
    for (int py = ylo; py <= yhi; ++py)
        for (int px = xlo; px <= xhi; ++px)
            s += __ldg(patches + (((pl * g.R + (y - py * g.stride)) * g.R + (x - px * g.stride)) * g.Hp + py) * g.Wp + px);
    const int n = max(yhi - ylo + 1, 0) * max(xhi - xlo + 1, 0);
    out[idx] = (mode == 0) ? s / (float)n : s;
}

// local2global_depth (:166-173): depth_map fp32 + depth_mask int32, both [B,R,R,Hp,Wp] -> depth, confidence [B,H,W]
__global__ void __launch_bounds__(TPB) k_fold_depth(const float* __restrict__ dmap, const int* __restrict__ dmask, size_t B, BeGeom g,
                                                    float* __restrict__ depth, float* __restrict__ conf) {
    const size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (idx >= B * g.H * g.W) return;
    const int x = (int)(idx % g.W), y = (int)((idx / g.W) % g.H);
    const size_t pl = idx / ((size_t)g.W * g.H);
    int ylo, yhi, xlo, xhi;
    cover(y, g.R, g.stride, g.Hp, &ylo, &yhi);
    cover(x, g.R, g.stride, g.Wp, &xlo, &xhi);
    float s = 0.0f, cnt = 0.0f;
    for (int py = ylo; py <= yhi; ++py)
        for (int px = xlo; px <= xhi; ++px) {
            const size_t o = (((pl * g.R + (y - py * g.stride)) * g.R + (x - px * g.stride)) * g.Hp + py) * g.Wp + px;
            s += __ldg(dmap + o);
            cnt += (__ldg(dmask + o) > 0) ? 1.0f : 0.0f;
        }
    const int n = max(yhi - ylo + 1, 0) * max(xhi - xlo + 1, 0);
    depth[idx] = s / (cnt > 0.0f ? cnt : 1.0f);
    conf[idx] = cnt / (float)n;
}

// img [P,H,W] -> patches [P,R,R,Hp,Wp] (nn.Unfold); mode 0: values divided by num_patches (adjoint of k_fold mode 0), mode 1: plain
__global__ void __launch_bounds__(TPB) k_unfold(const float* __restrict__ img, size_t P, BeGeom g, int mode, float* __restrict__ patches) {
    const size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
    const size_t per = (size_t)g.R * g.R * g.Hp * g.Wp;
    if (idx >= P * per) return;
    const int px = (int)(idx % g.Wp), py = (int)((idx / g.Wp) % g.Hp);
    const int j = (int)((idx / ((size_t)g.Wp * g.Hp)) % g.R), i = (int)((idx / ((size_t)g.Wp * g.Hp * g.R)) % g.R);
    const size_t pl = idx / per;
    const int y = py * g.stride + i, x = px * g.stride + j;
    float v = __ldg(img + (pl * g.H + y) * g.W + x);
    if (mode == 0) {
        int ylo, yhi, xlo, xhi;
        cover(y, g.R, g.stride, g.Hp, &ylo, &yhi);
        cover(x, g.R, g.stride, g.Wp, &xlo, &xhi);
        v /= (float)((yhi - ylo + 1) * (xhi - xlo + 1));
    }
    patches[idx] = v;
}

inline unsigned blocks_for(size_t n) { return (unsigned)((n + TPB - 1) / TPB); }

}  // namespace

// ---------------------------------------------------------------------------------------------------
void be_op_params2dists(const float* params, int K, int B, size_t Lsp, int R, float w, float* dists, cudaStream_t st) {
    k_params2dists<<<blocks_for((size_t)B * Lsp), TPB, 0, st>>>(params, K, B, Lsp, R, w, dists); ++g_be_launches;
}
void be_op_params2dists_bwd(const float* params, int K, const float* gd, int B, size_t Lsp, int R, float w, float* gp, cudaStream_t st) {
    k_params2dists_bwd<<<blocks_for((size_t)B * Lsp), TPB, 0, st>>>(params, K, gd, B, Lsp, R, w, gp); ++g_be_launches;
}
void be_op_indicators(const float* dists, const float* etas, int B, size_t Lsp, int RR, float* wedges, cudaStream_t st) {
    k_indicators<<<blocks_for((size_t)B * RR * Lsp), TPB, 0, st>>>(dists, etas, B, Lsp, RR, wedges); ++g_be_launches;
}
void be_op_indicators_bwd(const float* dists, const float* etas, const float* gw, int B, size_t Lsp, int RR, float* gd, float* ge,
                          cudaStream_t st) {
    k_indicators_bwd<<<blocks_for((size_t)B * Lsp), TPB, 0, st>>>(dists, etas, gw, B, Lsp, RR, gd, ge); ++g_be_launches;
}
void be_op_unary(int op, const float* x, float p0, const BeCam& cam, size_t n, float* y, cudaStream_t st) {
    k_unary<<<blocks_for(n), TPB, 0, st>>>(op, x, p0, cam, n, y); ++g_be_launches;
}
void be_op_unary_bwd(int op, const float* x, const float* gy, float p0, const BeCam& cam, size_t n, float* gx, cudaStream_t st) {
    k_unary_bwd<<<blocks_for(n), TPB, 0, st>>>(op, x, gy, p0, cam, n, gx); ++g_be_launches;
}
static unsigned smish_grid(size_t work) {   // a few CTAs per SM, grid-stride inside
    const size_t want = (work + TPB - 1) / TPB, cap = 148 * 16;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}
void be_op_smish(const float* x, size_t n, float* y, cudaStream_t st) {
    const int v4 = (n % 4 == 0) && (((uintptr_t)x | (uintptr_t)y) % 16 == 0);
    k_smish<<<smish_grid(v4 ? n / 4 : n), TPB, 0, st>>>(x, n, y, v4); ++g_be_launches;
}
void be_op_smish_bwd(const float* x, const float* gy, size_t n, float* gx, cudaStream_t st) {
    const int v4 = (n % 4 == 0) && (((uintptr_t)x | (uintptr_t)gy | (uintptr_t)gx) % 16 == 0);
    k_smish_bwd<<<smish_grid(v4 ? n / 4 : n), TPB, 0, st>>>(x, gy, n, gx, v4); ++g_be_launches;
}
void be_op_depth(const float* e1, const float* e2, const BeCam& cam, size_t n, float* z, cudaStream_t st) {
    k_depth<<<blocks_for(n), TPB, 0, st>>>(e1, e2, cam, n, z); ++g_be_launches;
}
void be_op_depth_bwd(const float* e1, const float* e2, const float* gz, const BeCam& cam, size_t n, float* g1, float* g2, cudaStream_t st) {
    k_depth_bwd<<<blocks_for(n), TPB, 0, st>>>(e1, e2, gz, cam, n, g1, g2); ++g_be_launches;
}
void be_op_inverse3(const float* A, size_t n, float* out, cudaStream_t st) {
    k_inverse3<<<blocks_for(n), TPB, 0, st>>>(A, n, out); ++g_be_launches;
}
void be_op_inverse3_bwd(const float* inv, const float* g, size_t n, float* gA, cudaStream_t st) {
    k_inverse3_bwd<<<blocks_for(n), TPB, 0, st>>>(inv, g, n, gA); ++g_be_launches;
}
void be_op_sobel(const float* img, size_t N, int H, int W, float* out, cudaStream_t st) {
    k_sobel<<<blocks_for(N * (H - 2) * (W - 2)), TPB, 0, st>>>(img, N, H, W, out); ++g_be_launches;
}
void be_op_sobel_bwd(const float* img, const float* gout, size_t N, int H, int W, float* gimg, cudaStream_t st) {
    k_sobel_bwd<<<blocks_for(N * H * W), TPB, 0, st>>>(img, gout, N, H, W, gimg); ++g_be_launches;
}
void be_op_fold(const float* patches, size_t P, const BeGeom& g, int mode, float* out, cudaStream_t st) {
    k_fold<<<blocks_for(P * g.H * g.W), TPB, 0, st>>>(patches, P, g, mode, out); ++g_be_launches;
}
void be_op_fold_depth(const float* dmap, const int* dmask, size_t B, const BeGeom& g, float* depth, float* conf, cudaStream_t st) {
    k_fold_depth<<<blocks_for(B * g.H * g.W), TPB, 0, st>>>(dmap, dmask, B, g, depth, conf); ++g_be_launches;
}
void be_op_unfold(const float* img, size_t P, const BeGeom& g, int mode, float* patches, cudaStream_t st) {
    k_unfold<<<blocks_for(P * g.R * g.R * g.Hp * g.Wp), TPB, 0, st>>>(img, P, g, mode, patches); ++g_be_launches;
}

// ---------------------------------------------------------------------------------------------------
// Glue of the inference driver (blurry_edges_test.py:119-138), SURVEY.md section 8f rank 1
// ---------------------------------------------------------------------------------------------------
namespace {

// image [M,3,H,W] -> vec [M*Hp*Wp, 3, R, R]  (nn.Unfold + permute(0,4,5,1,2,3).reshape of :119-121 in one pass)
__global__ void __launch_bounds__(TPB) k_patch_gather(const float* __restrict__ img, size_t M, BeGeom g, float* __restrict__ vec) {
    const size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
    const size_t per = (size_t)3 * g.R * g.R;
    if (idx >= M * g.Hp * g.Wp * per) return;
    const int j = (int)(idx % g.R), i = (int)((idx / g.R) % g.R), c = (int)((idx / ((size_t)g.R * g.R)) % 3);
    const size_t n = idx / per;
    const int px = (int)(n % g.Wp), py = (int)((n / g.Wp) % g.Hp);
    const size_t m = n / ((size_t)g.Wp * g.Hp);
    vec[idx] = __ldg(img + ((m * 3 + c) * g.H + py * g.stride + i) * g.W + px * g.stride + j);
}

// params [2,L,10] (raw LocalStage output) + colours [2,3,3,Hp,Wp] -> pm [1,L,38] (:123-132):
// per image (xy/3, (wrap(angles)-pi)/pi, eta_coef-0.5, (colours-0.5)*2), the two images concatenated per patch
__global__ void __launch_bounds__(TPB) k_assemble_pm(const float* __restrict__ params, const float* __restrict__ colors, size_t B, size_t L,
                                                     float* __restrict__ pm) {
    const size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (idx >= B * L * 38) return;
    const int k = (int)(idx % 38);
    const size_t l = (idx / 38) % L, b = idx / (38 * L);
    const int m = k / 19, q = k % 19;
    const size_t item = b * 2 + m;
    float v;
    if (q < 10) {
        const float p = __ldg(params + (item * L + l) * 10 + q);
        if (q < 4) v = p / 3.0f;
        else if (q < 8) v = (be_wrap_2pi(p) - BE_PI_F) / BE_PI_F;
        else v = p - 0.5f;
    } else {
        v = (__ldg(colors + (item * 9 + (q - 10)) * L + l) - 0.5f) * 2.0f;      // colours [item][c*3+w][l]
    }
    pm[idx] = v;
}

// eval_depth (utils/metrics.py:3-21) partial sums per image: n, #acc<tau, #acc<tau^2, #acc<tau^3, sum err^2, sum err/gt
__global__ void __launch_bounds__(TPB) k_eval_depth(const float* __restrict__ pred, const float* __restrict__ gt, int H, int W, int crop,
                                                    float z_min, float z_max, float tau, double* __restrict__ out6) {
    __shared__ double sh[TPB][6];
    const size_t img = blockIdx.y;
    const int Hc = H - 2 * crop, Wc = W - 2 * crop;
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (size_t t = (size_t)blockIdx.x * TPB + threadIdx.x; t < (size_t)Hc * Wc; t += (size_t)gridDim.x * TPB) {
        const int y = (int)(t / Wc) + crop, x = (int)(t % Wc) + crop;
        const float raw = pred[(img * H + y) * W + x], gv = gt[(img * H + y) * W + x];
        if (!(raw > 0.0f)) continue;                                         // error_mask = depth_map > 0 (blurry_edges_test.py:148)
        const float p = fminf(fmaxf(raw, z_min), z_max);
        const float err = fabsf(gv - p);
        const float pn = fminf(fmaxf((p - z_min) / (z_max - z_min), 0.0f), 1.0f), gn = fminf(fmaxf((gv - z_min) / (z_max - z_min), 0.0f), 1.0f);
        const float acc = fmaxf(gn / (pn + 1e-8f), pn / (gn + 1e-8f));
        a[0] += 1.0; a[1] += acc < tau; a[2] += acc < tau * tau; a[3] += acc < tau * tau * tau;
        a[4] += (double)err * err; a[5] += (double)err / gv;
    }
    for (int k = 0; k < 6; ++k) sh[threadIdx.x][k] = a[k];
    __syncthreads();
    for (int off = TPB / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off)
            for (int k = 0; k < 6; ++k) sh[threadIdx.x][k] += sh[threadIdx.x + off][k];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int k = 0; k < 6; ++k) atomicAdd(out6 + img * 6 + k, sh[0][k]);
}

}  // namespace

void be_op_patch_gather(const float* img, size_t M, const BeGeom& g, float* vec, cudaStream_t st) {
    k_patch_gather<<<blocks_for(M * g.Hp * g.Wp * 3 * g.R * g.R), TPB, 0, st>>>(img, M, g, vec); ++g_be_launches;
}
void be_op_assemble_pm(const float* params, const float* colors, size_t B, size_t L, float* pm, cudaStream_t st) {
    k_assemble_pm<<<blocks_for(B * L * 38), TPB, 0, st>>>(params, colors, B, L, pm); ++g_be_launches;
}
void be_op_eval_depth(const float* pred, const float* gt, size_t N, int H, int W, int crop, double* out6, cudaStream_t st) {
    dim3 grid(32, (unsigned)N);
    k_eval_depth<<<grid, TPB, 0, st>>>(pred, gt, H, W, crop, 0.75f, 1.18f, 1.25f, out6); ++g_be_launches;
}

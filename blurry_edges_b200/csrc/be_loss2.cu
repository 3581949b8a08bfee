// be_loss2_kernel: global-stage loss forward + analytic backward (global_training.py:62-157), second generation.
//
// Same mathematics as be_loss_kernel<false> of be_train.cu (which stays as the local-stage kernel), restructured after its
// ncu capture (profiles/r1_loss_kernel_full.txt: 19 225 warp-instructions per patch, issue slots 49 % busy, 7 CTA barriers
// and two serial fp64 solves per patch, 2 CTAs/SM):
//   * the ridge colours C and the inverse normal matrix M^-1 of every patch come from the TRAINFWD render pass that has to run
//     first anyway (its solver warp stores a 64-byte record per patch), so the loss kernel has no phase-1 sums, no first
//     reduction, no fp64 solve and two barriers less;
//   * the second solve V = M^-1 A^T G, S = V C^T + C V^T is done redundantly by every warp in fp32 (80 instructions) instead of
//     by warp 0 between two CTA barriers; the chain rule to the raw parameters of patch k-1 is done by one (rotating) warp
//     while the others already work on patch k: 3 CTA barriers per patch instead of 7, no serial section;
//   * the two pixel slots of a thread are the two halves of packed fp32x2 registers (be_pack.cuh) wherever the operands do
//     not come straight from a vector load.
#include "be_internal.h"
#include "be_pack.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int NT = BE_THREADS;               // 224 threads, 7 warps, two pixel slots per thread
constexpr int HALO = BE_MAX_R + 1;
constexpr int NE = NT + 2 * HALO;

__device__ __forceinline__ float warp_reduce16(const float (&v)[16], int lane) {
    float a[8], b[4], c[2];
    bool hi_ = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (hi_ ? v[i + 8] : v[i]) + __shfl_xor_sync(FULL, hi_ ? v[i] : v[i + 8], 16);
    hi_ = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = (hi_ ? a[i + 4] : a[i]) + __shfl_xor_sync(FULL, hi_ ? a[i] : a[i + 4], 8);
    hi_ = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) c[i] = (hi_ ? b[i + 2] : b[i]) + __shfl_xor_sync(FULL, hi_ ? b[i] : b[i + 2], 4);
    hi_ = lane & 2;
    float d = (hi_ ? c[1] : c[0]) + __shfl_xor_sync(FULL, hi_ ? c[0] : c[1], 2);
    d += __shfl_xor_sync(FULL, d, 1);
    return d;
}

template <int RCT>   // RCT = 21: patch size known at compile time (neighbour offsets become immediates), 0: generic
// 128 registers (2 CTAs/SM): capped at 80 for 3 CTAs/SM the kernel spills 284 bytes and is 4 % slower (measured)
__global__ void __maxnreg__(128) be_loss2_kernel(const BeLossArgs a) {
    __shared__ __align__(16) float s_rec[2][BE_REC];
    __shared__ __align__(16) float s_grec[2][BE_GREC];
    __shared__ __align__(16) float s_crec[2][BE_CREC];
    __shared__ float s_axis[BE_MAX_R + 3];
    __shared__ float s_part[BE_WARPS][16];
    __shared__ float s_part3[BE_WARPS][16];
    // Stencil exchange planes in PAIR layout: entry e holds, per channel, (value at pixel e-HALO, value at pixel e-HALO+NT), so the
    // neighbour of both slots of thread t at offset `off` is the one entry t+HALO+off (3 LDS.128 for 6 channels of two pixels).
    // Pixels within HALO of the seam are written twice (as the low half of their own entry and as the high half of the entry NT
    // below); entries outside the patch stay zero.  Row wrap-around needs no mask: a wrapped neighbour is a border pixel, whose
    // Sobel gradients are zero, and only interior pixels (which never wrap) use the rendered-patch plane.
    __shared__ float4 s_X[9 * NE];
    float4* const s_P2 = s_X;                     // [3][NE] rendered patch: (c0,c1) (c2,c3) (c4,c5)
    float4* const s_G2 = s_X + 3 * NE;            // [6][NE] Sobel gradients: gx (c0,c1) (c2,c3) (c4,c5), gy (c0,c1) (c2,c3) (c4,c5)
    __shared__ float4 s_stash[2][NT];            // thread-private: (d1, d2) and (global boundary, bndry_dist) pairs, stage A -> D

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const BeGeom g = a.g;
    const int R = RCT ? RCT : g.R, RR = R * R;

    int blk = blockIdx.x;
    const int run = blk % a.runs_per_row; blk /= a.runs_per_row;
    const int py = blk % g.Hp;
    const int b = blk / g.Hp;
    const int px0 = run * a.G;
    const int n = min(a.G, g.Wp - px0);
    const int y0 = py * g.stride;
    const size_t patch0 = ((size_t)b * g.Hp + py) * g.Wp + px0;

    // records of the first patch: lanes 0-7 table, 8-10 gtable, 11-14 crec
    auto fetch = [&](size_t patch, int l) -> float4 {
        if (l < 8) return __ldg(reinterpret_cast<const float4*>(a.table + patch * BE_REC) + l);
        if (l < 8 + BE_GREC / 4) return __ldg(reinterpret_cast<const float4*>(a.gtable + patch * BE_GREC) + (l - 8));
        return __ldg(reinterpret_cast<const float4*>(a.crec + patch * BE_CREC) + (l - 8 - BE_GREC / 4));
    };
    auto stash = [&](int buf, int l, float4 v) {
        if (l < 8) reinterpret_cast<float4*>(s_rec[buf])[l] = v;
        else if (l < 8 + BE_GREC / 4) reinterpret_cast<float4*>(s_grec[buf])[l - 8] = v;
        else reinterpret_cast<float4*>(s_crec[buf])[l - 8 - BE_GREC / 4] = v;
    };
    constexpr int NFETCH = 8 + BE_GREC / 4 + BE_CREC / 4;
    if (tid < R) s_axis[tid] = be_axis(tid, R);
    if (tid < NFETCH) stash(0, tid, fetch(patch0, tid));

    bool valid[2], interior[2];
    int q[2], pi[2], pj[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int qq = tid + s * NT;
        valid[s] = qq < RR;
        q[s] = valid[s] ? qq : 0;
        pi[s] = q[s] / R; pj[s] = q[s] % R;
        interior[s] = valid[s] && pi[s] >= 1 && pi[s] <= R - 2 && pj[s] >= 1 && pj[s] <= R - 2;
    }
    for (int i = tid; i < 9 * NE; i += NT) s_X[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // store one pair-layout entry (+ the duplicates near the seam) of `nch` channels given as (slot0, slot1) pairs
    auto store_pairs = [&](float4* plane0, const f2* v, int nch) {
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (2 * k < nch) plane0[k * NE + tid + HALO] = make_float4(lo(v[2 * k]), hi(v[2 * k]), lo(v[2 * k + 1]), hi(v[2 * k + 1]));
        if (tid >= NT - HALO) {                    // my low pixel is also the high half of entry tid - NT
            float* d = reinterpret_cast<float*>(plane0 + (tid - NT + HALO)) + 1;
#pragma unroll
            for (int c = 0; c < 12; ++c)
                if (c < nch) d[(c >> 1) * (NE * 4) + (c & 1) * 2] = lo(v[c]);
        }
        if (tid < HALO) {                          // my high pixel is also the low half of entry tid + NT
            float* d = reinterpret_cast<float*>(plane0 + (tid + NT + HALO));
#pragma unroll
            for (int c = 0; c < 12; ++c)
                if (c < nch) d[(c >> 1) * (NE * 4) + (c & 1) * 2] = hi(v[c]);
        }
    };
    const f2 mI = mk2(interior[0] ? 1.0f : 0.0f, interior[1] ? 1.0f : 0.0f);
    const f2 kIs = mul2(bc2(2.0f * a.ks), mI), kIsc = mul2(bc2(2.0f * a.ksc), mI);    // Sobel-loss weights, zero off the interior
    __syncthreads();
    const f2 Y = mk2(s_axis[pi[0]], s_axis[pi[1]]), X = mk2(s_axis[pj[0]], s_axis[pj[1]]);
    const float vm0 = (RCT == BE_MAX_R || valid[0]) ? 1.0f : 0.0f, vm1 = valid[1] ? 1.0f : 0.0f;   // R = 21: every thread has a low pixel
    const float kd = a.gamma_d / (float)(*a.mask_count);
    const size_t TPS = (size_t)a.NB * g.H * g.W * 4;      // floats between consecutive float4 planes of T
    const int np = 12;
    const float k2c = 2.0f * a.kc, k2cc = 2.0f * a.kcc;

    float lossacc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto fold = [&](f2 v) { return (RCT == BE_MAX_R) ? fmaf(hi(v), vm1, lo(v)) : fmaf(hi(v), vm1, lo(v) * vm0); };       // both slots of a thread, padding slots dropped

    // chain rule of one finished patch (global_training.py:141-145 backward): 14 per-patch sums -> 12 raw-parameter gradients
    auto chain = [&](int kp) {
        float t = 0.0f;
        if (lane < 16) {
#pragma unroll
            for (int wv = 0; wv < BE_WARPS; ++wv) t += s_part3[wv][lane];
        }
        float S[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) S[i] = __shfl_sync(FULL, t, i);
        if (lane == 0 && a.grad != nullptr) {
            const float* gr = s_grec[kp & 1];      // deta_dcoef[4], dz_deta[4], xy_scale, ang_scale
            float* out = a.grad + (patch0 + kp) * np;
            const float xs = gr[8], as = gr[9];
            float4 o0, o1, o2;
            o0.x = xs * S[0]; o0.y = xs * S[1]; o0.z = xs * S[4]; o0.w = xs * S[5];
            o1.x = as * (S[2] + S[3]); o1.y = as * S[3]; o1.z = as * (S[6] + S[7]); o1.w = as * S[7];
            o2.x = (S[8] + S[12] * gr[4]) * gr[0];
            o2.y = (S[9] + S[13] * gr[6]) * gr[1];
            o2.z = (S[10] + S[12] * gr[5]) * gr[2];
            o2.w = (S[11] + S[13] * gr[7]) * gr[3];
            float4* o4 = reinterpret_cast<float4*>(out);     // 48-byte rows: 16-byte aligned
            o4[0] = o0; o4[1] = o1; o4[2] = o2;
        }
    };

    for (int k = 0; k < n; ++k) {
        const int cur = k & 1;
        const int x0 = (px0 + k) * g.stride;
        float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (warp == 0 && lane < NFETCH && k + 1 < n) nxt = fetch(patch0 + k + 1, lane);
        auto load_patch = [&](BePatch& P) {
            const float4* q4 = reinterpret_cast<const float4*>(s_rec[cur]);
            const float4 r0 = q4[0], r1 = q4[1], r2 = q4[2], r3 = q4[3], r4 = q4[4];
            P.sn[0] = r0.x; P.sn[1] = r0.y; P.sn[2] = r0.z; P.sn[3] = r0.w;
            P.cs[0] = r1.x; P.cs[1] = r1.y; P.cs[2] = r1.z; P.cs[3] = r1.w;
            P.vx[0] = r2.x; P.vx[1] = r2.y; P.vy[0] = r2.z; P.vy[1] = r2.w;
            P.flip[0] = r3.x; P.flip[1] = r3.y; P.z[0] = r3.z; P.z[1] = r3.w;
            P.inv_eta[0] = r4.x; P.inv_eta[1] = r4.y; P.inv_eta[2] = r4.z; P.inv_eta[3] = r4.w;
        };
        auto load_colors = [&](float* C) {
            const float4* c4 = reinterpret_cast<const float4*>(s_crec[cur]);
            const float4 c0 = c4[0], c1 = c4[1];
            C[0] = c0.x; C[1] = c0.y; C[2] = c0.z; C[3] = c0.w; C[4] = c1.x; C[5] = c1.y; C[6] = c1.z; C[7] = c1.w;
            C[8] = s_crec[cur][8];
        };
        const float* tp[2];      // plane 0 of each slot's pixel in the packed targets
#pragma unroll
        for (int s = 0; s < 2; ++s) tp[s] = a.T + (((size_t)b * g.H + y0 + pi[s]) * g.W + x0 + pj[s]) * 4;

        // ---------------- stage A: distances, soft indicators, render, direct dL/dP ----------------
        f2 h[4], G[6];
        {
            BePatch P;
            float C[9];
            load_patch(P);
            load_colors(C);
            float2 t1[2];
            float4 t2[2], t3[2], t4[2];                  // targets: issued before the arithmetic that hides their latency
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                t1[s] = __ldg(reinterpret_cast<const float2*>(tp[s] + TPS + 2));
                t2[s] = __ldg(reinterpret_cast<const float4*>(tp[s] + 2 * TPS));
                t3[s] = __ldg(reinterpret_cast<const float4*>(tp[s] + 3 * TPS));
                t4[s] = __ldg(reinterpret_cast<const float4*>(tp[s] + 4 * TPS));
            }
            f2 d1, d2;
            be_pixel_dists2(P, X, Y, g.w, &d1, &d2);
            f2 Pv[6];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                h[2 * m] = be_h2(d1, P.inv_eta[2 * m]);
                h[2 * m + 1] = be_h2(d2, P.inv_eta[2 * m + 1]);
                const f2 gg = sub2(bc2(1.0f), h[2 * m + 1]);
                const f2 u0 = mul2(sub2(bc2(1.0f), h[2 * m]), gg), u1 = mul2(h[2 * m], gg), u2 = h[2 * m + 1];
#pragma unroll
                for (int c = 0; c < 3; ++c) Pv[3 * m + c] = fma2(u0, bc2(C[c]), fma2(u1, bc2(C[3 + c]), mul2(u2, bc2(C[6 + c]))));
            }
            store_pairs(s_P2, Pv, 6);
            float gb_[2], bd_[2], Gs[2][6];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                float pv[6];
#pragma unroll
                for (int c = 0; c < 6; ++c) pv[c] = s ? hi(Pv[c]) : lo(Pv[c]);
                const float gt[6] = {t1[s].x, t1[s].y, t2[s].x, t2[s].y, t2[s].z, t2[s].w};
                const float gi[6] = {t3[s].x, t3[s].y, t3[s].z, t3[s].w, t4[s].x, t4[s].y};
                gb_[s] = t4[s].z; bd_[s] = t4[s].w;
                float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const float e1 = pv[c] - gt[c], e2 = pv[c] - gi[c];
                    l0 = fmaf(e1, e1, l0);
                    l1 = fmaf(e2, e2, l1);
                    Gs[s][c] = fmaf(k2cc, e2, k2c * e1);
                }
                lossacc[0] = fmaf(l0, s ? vm1 : vm0, lossacc[0]);
                lossacc[1] = fmaf(l1, s ? vm1 : vm0, lossacc[1]);
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) G[c] = mk2(Gs[0][c], Gs[1][c]);
            s_stash[0][tid] = make_float4(lo(d1), hi(d1), lo(d2), hi(d2));
            s_stash[1][tid] = make_float4(gb_[0], gb_[1], bd_[0], bd_[1]);
        }
        __syncthreads();   // (X1) rendered patch visible

        // chain rule of the previous patch, by one warp, while the others go on
        if (k >= 1 && warp == (k - 1) % BE_WARPS) chain(k - 1);

        // ---------------- stage B: Sobel magnitude of the rendered patch, its loss and gradient (both slots packed) ----------------
        {
            float4 t6[2], t7[2], t8[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                t6[s] = __ldg(reinterpret_cast<const float4*>(tp[s] + 6 * TPS));
                t7[s] = __ldg(reinterpret_cast<const float4*>(tp[s] + 7 * TPS));
                t8[s] = __ldg(reinterpret_cast<const float4*>(tp[s] + 8 * TPS));
            }
            f2 sx[6], sy[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) sx[c] = sy[c] = bc2(0.0f);
#pragma unroll
            for (int oi = -1; oi <= 1; ++oi)
#pragma unroll
                for (int oj = -1; oj <= 1; ++oj) {
                    if (oi == 0 && oj == 0) continue;
                    const float wx = (float)(((oi == 0) ? 2 : 1) * oj);       // sobel_x[oi+1][oj+1]
                    const float wy = (float)(-oi * ((oj == 0) ? 2 : 1));      // sobel_y[oi+1][oj+1]
                    const float4* pn = s_P2 + (tid + HALO + oi * R + oj);
                    const float4 v0 = pn[0], v1 = pn[NE], v2 = pn[2 * NE];
                    const f2 pv[6] = {mk2(v0.x, v0.y), mk2(v0.z, v0.w), mk2(v1.x, v1.y), mk2(v1.z, v1.w), mk2(v2.x, v2.y), mk2(v2.z, v2.w)};
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        if (wx != 0.0f) sx[c] = fma2(bc2(wx), pv[c], sx[c]);
                        if (wy != 0.0f) sy[c] = fma2(bc2(wy), pv[c], sy[c]);
                    }
                }
            const float dgt[2][6] = {{t6[0].x, t6[0].y, t6[0].z, t6[0].w, t7[0].x, t7[0].y}, {t6[1].x, t6[1].y, t6[1].z, t6[1].w, t7[1].x, t7[1].y}};
            const float dgi[2][6] = {{t7[0].z, t7[0].w, t8[0].x, t8[0].y, t8[0].z, t8[0].w}, {t7[1].z, t7[1].w, t8[1].x, t8[1].y, t8[1].z, t8[1].w}};
            f2 gxy[12], l3 = bc2(0.0f), l4 = bc2(0.0f);
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const f2 v = fma2(sx[c], sx[c], fma2(sy[c], sy[c], bc2(1e-8f)));
                const f2 ir = mk2(be_rsqrt(lo(v)), be_rsqrt(hi(v)));          // v >= 1e-8: no denormal handling needed
                const f2 mag = mul2(v, ir);
                const f2 e1 = mk2(lo(mag) - dgt[0][c], hi(mag) - dgt[1][c]), e2 = mk2(lo(mag) - dgi[0][c], hi(mag) - dgi[1][c]);
                l3 = fma2(e1, e1, l3);
                l4 = fma2(e2, e2, l4);
                const f2 gm = mul2(fma2(kIsc, e2, mul2(kIs, e1)), ir);         // zero off the interior
                gxy[c] = mul2(gm, sx[c]);
                gxy[6 + c] = mul2(gm, sy[c]);
            }
            lossacc[3] = fmaf(lo(l3), lo(mI), fmaf(hi(l3), hi(mI), lossacc[3]));
            lossacc[4] = fmaf(lo(l4), lo(mI), fmaf(hi(l4), hi(mI), lossacc[4]));
            store_pairs(s_G2, gxy, 12);
        }
        __syncthreads();   // (X2) Sobel gradients visible

        // ---------------- stage C: Sobel adjoint into G, then A^T G ----------------
        if (warp == 0 && lane < NFETCH && k + 1 < n) stash(cur ^ 1, lane, nxt);     // visible after barrier (R2)
        {
#pragma unroll
            for (int di = -1; di <= 1; ++di)
#pragma unroll
                for (int dj = -1; dj <= 1; ++dj) {
                    if (di == 0 && dj == 0) continue;
                    const float wx = (float)(-dj * ((di == 0) ? 2 : 1));   // weight of gx(i+di, j+dj) in dL/dP(i,j)
                    const float wy = (float)(di * ((dj == 0) ? 2 : 1));    // weight of gy(i+di, j+dj)
                    const float4* pn = s_G2 + (tid + HALO + di * R + dj);
                    if (wx != 0.0f) {
                        const float4 v0 = pn[0], v1 = pn[NE], v2 = pn[2 * NE];
                        G[0] = fma2(bc2(wx), mk2(v0.x, v0.y), G[0]); G[1] = fma2(bc2(wx), mk2(v0.z, v0.w), G[1]);
                        G[2] = fma2(bc2(wx), mk2(v1.x, v1.y), G[2]); G[3] = fma2(bc2(wx), mk2(v1.z, v1.w), G[3]);
                        G[4] = fma2(bc2(wx), mk2(v2.x, v2.y), G[4]); G[5] = fma2(bc2(wx), mk2(v2.z, v2.w), G[5]);
                    }
                    if (wy != 0.0f) {
                        const float4 v0 = pn[3 * NE], v1 = pn[4 * NE], v2 = pn[5 * NE];
                        G[0] = fma2(bc2(wy), mk2(v0.x, v0.y), G[0]); G[1] = fma2(bc2(wy), mk2(v0.z, v0.w), G[1]);
                        G[2] = fma2(bc2(wy), mk2(v1.x, v1.y), G[2]); G[3] = fma2(bc2(wy), mk2(v1.z, v1.w), G[3]);
                        G[4] = fma2(bc2(wy), mk2(v2.x, v2.y), G[4]); G[5] = fma2(bc2(wy), mk2(v2.z, v2.w), G[5]);
                    }
                }
            f2 sums[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) sums[i] = bc2(0.0f);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const f2 gg = sub2(bc2(1.0f), h[2 * m + 1]);
                const f2 u[3] = {mul2(sub2(bc2(1.0f), h[2 * m]), gg), mul2(h[2 * m], gg), h[2 * m + 1]};
#pragma unroll
                for (int wd = 0; wd < 3; ++wd)
#pragma unroll
                    for (int c = 0; c < 3; ++c) sums[3 * wd + c] = fma2(u[wd], G[3 * m + c], sums[3 * wd + c]);
            }
            float ssum[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) ssum[i] = (i < 9) ? fold(sums[i]) : 0.0f;
            const float tot = warp_reduce16(ssum, lane);
            if (!(lane & 1)) s_part[warp][lane >> 1] = tot;
        }
        __syncthreads();   // (R2) A^T G partials visible

        // ---------------- stage D: second solve (every warp), per-pixel backward -> 14 per-patch sums ----------------
        {
            BePatch P;
            float C[9];
            load_patch(P);
            load_colors(C);
            float ny[2][6];
            float zgv[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const float4 q0 = __ldg(reinterpret_cast<const float4*>(tp[s]));
                const float2 q1 = __ldg(reinterpret_cast<const float2*>(tp[s] + TPS));
                ny[s][0] = q0.x; ny[s][1] = q0.y; ny[s][2] = q0.z; ny[s][3] = q0.w; ny[s][4] = q1.x; ny[s][5] = q1.y;
                zgv[s] = __ldg(tp[s] + 5 * TPS);
            }
            const float4 sd = s_stash[0][tid], sg = s_stash[1][tid];
            const f2 d1 = mk2(sd.x, sd.y), d2 = mk2(sd.z, sd.w), gbv = mk2(sg.x, sg.y), bdv = mk2(sg.z, sg.w);
            float V[9], Ssym[6];
            {
                float t = 0.0f;
                if (lane < 16) {
#pragma unroll
                    for (int wv = 0; wv < BE_WARPS; ++wv) t += s_part[wv][lane];
                }
                float AtG[9], Mi[6];
#pragma unroll
                for (int i = 0; i < 9; ++i) AtG[i] = __shfl_sync(FULL, t, i);
#pragma unroll
                for (int i = 0; i < 6; ++i) Mi[i] = s_crec[cur][9 + i];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float b0 = AtG[c], b1 = AtG[3 + c], b2 = AtG[6 + c];
                    V[0 + c] = Mi[0] * b0 + Mi[1] * b1 + Mi[2] * b2;
                    V[3 + c] = Mi[1] * b0 + Mi[3] * b1 + Mi[4] * b2;
                    V[6 + c] = Mi[2] * b0 + Mi[4] * b1 + Mi[5] * b2;
                }
                const int pi_[6] = {0, 0, 0, 1, 1, 2}, pj_[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    float sacc = 0.0f;
#pragma unroll
                    for (int c = 0; c < 3; ++c) sacc += V[3 * pi_[i] + c] * C[3 * pj_[i] + c] + C[3 * pi_[i] + c] * V[3 * pj_[i] + c];
                    Ssym[i] = sacc;
                }
            }
            const float Sm[9] = {Ssym[0], Ssym[1], Ssym[2], Ssym[1], Ssym[3], Ssym[4], Ssym[2], Ssym[4], Ssym[5]};
            f2 sums[14];
#pragma unroll
            for (int i = 0; i < 14; ++i) sums[i] = bc2(0.0f);
            f2 gd1 = bc2(0.0f), gd2 = bc2(0.0f);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const f2 h1 = h[2 * m], h2 = h[2 * m + 1];
                const f2 gg = sub2(bc2(1.0f), h2), g1 = sub2(bc2(1.0f), h1);
                const f2 u[3] = {mul2(g1, gg), mul2(h1, gg), h2};
                f2 gu[3];
#pragma unroll
                for (int wd = 0; wd < 3; ++wd) {
                    // dL/du_w = sum_c (G_c C[w][c] + y_c V[w][c]) - (S u)_w     (be_ridge_backward_pixel)
                    float yv[2];
#pragma unroll
                    for (int s = 0; s < 2; ++s)
                        yv[s] = fmaf(ny[s][3 * m], V[3 * wd], fmaf(ny[s][3 * m + 1], V[3 * wd + 1], ny[s][3 * m + 2] * V[3 * wd + 2]));
                    f2 t = mk2(yv[0], yv[1]);
#pragma unroll
                    for (int c = 0; c < 3; ++c) t = fma2(G[3 * m + c], bc2(C[3 * wd + c]), t);
#pragma unroll
                    for (int v = 0; v < 3; ++v) t = fma2(u[v], bc2(-Sm[3 * wd + v]), t);
                    gu[wd] = t;
                }
                const f2 gh1 = mul2(gg, sub2(gu[1], gu[0]));                                   // be_wedges_backward
                const f2 gh2 = sub2(gu[2], fma2(g1, gu[0], mul2(h1, gu[1])));
                f2 da, de;
                be_h_grad2(d1, P.inv_eta[2 * m], &da, &de);
                gd1 = fma2(gh1, da, gd1); sums[8 + 2 * m] = fma2(gh1, de, sums[8 + 2 * m]);
                be_h_grad2(d2, P.inv_eta[2 * m + 1], &da, &de);
                gd2 = fma2(gh2, da, gd2); sums[9 + 2 * m] = fma2(gh2, de, sums[9 + 2 * m]);
            }
            const f2 lb = be_boundary2(d1, d2);
            const f2 bl = mul2(bdv, lb);
            lossacc[5] += fold(mul2(bl, bl));
            const f2 eb = sub2(lb, gbv);
            lossacc[2] += fold(mul2(eb, eb));
            const f2 glb = fma2(bc2(2.0f * a.kbc), eb, mul2(mul2(bc2(2.0f * a.kbl), bdv), bl));
            float gb1[2], gb2[2], l6[2], s12[2], s13[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const float d1s = s ? hi(d1) : lo(d1), d2s = s ? hi(d2) : lo(d2);
                be_boundary_backward1(d1s, d2s, s ? hi(lb) : lo(lb), s ? hi(glb) : lo(glb), &gb1[s], &gb2[s]);
                const int mk = be_mask(d1s, d2s, false);                                       // global_training.py:84-90,121-127
                const bool on = (zgv[s] != 0.0f) && (mk != 0);
                const float e2 = on ? (((mk == 1) ? P.z[0] : P.z[1]) - zgv[s]) : 0.0f;
                l6[s] = e2 * e2;
                const float ge = on ? 2.0f * kd * e2 : 0.0f;          // select, not multiply: kd is inf when the batch mask is empty
                s12[s] = (mk == 1) ? ge : 0.0f;
                s13[s] = (mk == 1) ? 0.0f : ge;
            }
            lossacc[6] += fold(mk2(l6[0], l6[1]));
            sums[12] = mk2(s12[0], s12[1]); sums[13] = mk2(s13[0], s13[1]);
            gd1 = add2(gd1, mk2(gb1[0], gb1[1])); gd2 = add2(gd2, mk2(gb2[0], gb2[1]));
            be_wedge_backward2(P, 0, X, Y, g.w, gd1, &sums[0]);
            be_wedge_backward2(P, 1, X, Y, g.w, gd2, &sums[4]);
            float ssum[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) ssum[i] = (i < 14) ? fold(sums[i]) : 0.0f;
            const float tot = warp_reduce16(ssum, lane);
            if (!(lane & 1)) s_part3[warp][lane >> 1] = tot;
        }
    }
    __syncthreads();
    if (warp == 0) chain(n - 1);

    // ---------------- per-CTA partial loss sums ----------------
    {
        float sums[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) sums[i] = (i < 7) ? lossacc[i] : 0.0f;
        const float tot = warp_reduce16(sums, lane);
        if (!(lane & 1)) s_part[warp][lane >> 1] = tot;
        __syncthreads();
        if (tid < 8) {
            float t = 0.0f;
#pragma unroll
            for (int wv = 0; wv < BE_WARPS; ++wv) t += s_part[wv][tid];
            a.partials[(size_t)blockIdx.x * 8 + tid] = t;
        }
    }
}

}  // namespace

void be_launch_loss2(const BeLossArgs& a, cudaStream_t st) {
    const int grid = a.NB * a.g.Hp * a.runs_per_row;
    if (a.g.R == 21) be_loss2_kernel<21><<<grid, NT, 0, st>>>(a);
    else be_loss2_kernel<0><<<grid, NT, 0, st>>>(a);
    ++g_be_launches;
}

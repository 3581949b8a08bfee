// sm_100a kernels of the TRAINING side: global-stage loss (global_training.py:62-157) and local-stage loss
// (local_training.py:32-52), forward + analytic backward, plus the small kernels between the two passes.
//
//   be_run3_kernel<TRAINFWD>  (be_run3.cu) renders both images + boundary and folds them -> accumulator [B,H,W,8],
//                             counts the depth-term mask.
//   be_train_targets_kernel   accumulator -> global image / boundary (detached targets, :154-155) and the packed per-pixel target
//                             record T: noisy + ground-truth pixels, global image / boundary, log2(bndry_dist+1), z_gt, the
//                             ground-truth derivative and Sobel(global image) (:106-110,117-118,123-124), so that the loss kernel
//                             fetches everything about one pixel with a few 16-byte loads.
//   be_local_loss_kernel      local-stage loss: one CTA per patch; phase 1 + ridge solve, render, direct dL/dP, Sobel forward
//                             + adjoint through shared memory, A^T G + second solve, per-pixel backward to the per-patch
//                             sums, chain rule to the 10 raw parameters.  (The global-stage loss is be_loss2.cu.)
//   be_loss_reduce_kernel     per-CTA partial sums -> the seven loss terms and the weighted loss.
#include "be_internal.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int T_NY = 0, T_GT = 6, T_GI = 12, T_GB = 18, T_BD = 19, T_DGT = 20, T_DGI = 26, T_ZG = 32;

// T is stored as 8 planes of float4 per pixel ([8][B*H*W] float4, values 0..31) followed by one scalar plane (value 32, z_gt):
// the threads of a warp own neighbouring pixels, so a 16-byte load of one plane touches ~half the sectors of a record-major
// layout and every fetched sector is fully used.  No padding: the loss kernel re-reads the 441-pixel window of every patch and
// lives on the L1 hits of the 19 columns it shares with the previous patch; 33 floats per pixel are 58 KB per window (51 KB when
// img_gt is img_ny and the GT values are never touched), two resident CTAs have 124 KB of L1.
__device__ __forceinline__ size_t t_off(size_t plane_stride, size_t pix, int k) {
    return (k < 32) ? (size_t)(k >> 2) * plane_stride + pix * 4 + (k & 3) : (size_t)8 * plane_stride + pix;
}

__device__ __forceinline__ int cover_1d(int y, int R, int s, int np) {
    const int hi = min(y / s, np - 1);
    const int lo = (y - R + 1 <= 0) ? 0 : (y - R + s) / s;
    return max(hi - lo + 1, 0);
}

__device__ __forceinline__ float warp_reduce16(const float (&v)[16], int lane) {
    float a[8], b[4], c[2];
    bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (hi ? v[i + 8] : v[i]) + __shfl_xor_sync(FULL, hi ? v[i] : v[i + 8], 16);
    hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = (hi ? a[i + 4] : a[i]) + __shfl_xor_sync(FULL, hi ? a[i] : a[i + 4], 8);
    hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) c[i] = (hi ? b[i + 2] : b[i]) + __shfl_xor_sync(FULL, hi ? b[i] : b[i + 2], 4);
    hi = lane & 2;
    float d = (hi ? c[1] : c[0]) + __shfl_xor_sync(FULL, hi ? c[0] : c[1], 2);
    d += __shfl_xor_sync(FULL, d, 1);
    return d;
}

// ---------------------------------------------------------------------------------------------------
// Pairs [b0, b0 + nb) of a workspace laid out for Btot pairs (the host-buffer entry point runs the batch in chunks): `acc`, `T`,
// `gimg`, `gbnd` are the arrays of the WHOLE batch, the plane stride of T is that of Btot pairs.
//
// ONE kernel between the two passes (round 1 had two: normalise, then pack reading the normalised image back): a thread owns a pixel,
// normalises its accumulator cell (fold / closed-form cover count, utils/postprocessing_loss.py:151-173) and - for the Sobel
// magnitude of the global image (:114-117) - the cells of its 8 neighbours (same division, so the values are those the neighbours
// write themselves; the 9 cells come from L1/L2), gathers the pixel's targets and writes the packed record as whole float4s (every
// plane is stored once, 16 bytes per thread, instead of 33 scalar stores into 16-byte-strided planes).
__device__ __forceinline__ void acc_pixel(const float* __restrict__ acc, size_t cell, float n, float (&v)[7]) {
    const float4* src = reinterpret_cast<const float4*>(acc + cell * 8);
    const float4 q0 = __ldg(src), q1 = __ldg(src + 1);
    v[0] = q0.x / n; v[1] = q0.y / n; v[2] = q0.z / n; v[3] = q0.w / n; v[4] = q1.x / n; v[5] = q1.y / n; v[6] = q1.z / n;
}

__global__ void __launch_bounds__(256) be_train_targets_kernel(const float* __restrict__ acc, BeGeom g, int b0, int nb, int Btot,
                                                               const float* __restrict__ img_ny, const float* __restrict__ img_gt,
                                                               const float* __restrict__ bndry_dist, const float* __restrict__ deri,
                                                               const float* __restrict__ bndry_depth, float* __restrict__ T,
                                                               float* __restrict__ gimg, float* __restrict__ gbnd) {
    const bool same_gt = (img_gt == img_ny);          // the GT values are then never read (BeLossArgs::same_gt)
    const size_t HW = (size_t)g.H * g.W;
    const size_t lidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lidx >= (size_t)nb * HW) return;
    const size_t idx = lidx + (size_t)b0 * HW;
    const size_t p = idx % HW;
    const int b = (int)(idx / HW), y = (int)(p / g.W), x = (int)(p % g.W);
    const bool interior = (y >= 1 && y < g.H - 1 && x >= 1 && x < g.W - 1);
    // cover counts of the three rows / columns around the pixel (separable)
    float cy[3], cx[3];
#pragma unroll
    for (int o = -1; o <= 1; ++o) {
        cy[o + 1] = (float)cover_1d(min(max(y + o, 0), g.H - 1), g.R, g.stride, g.Hp);
        cx[o + 1] = (float)cover_1d(min(max(x + o, 0), g.W - 1), g.R, g.stride, g.Wp);
    }
    float v[7];
    acc_pixel(acc, idx, cy[1] * cx[1], v);
    float dgi[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dgt[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (interior) {
        // Sobel magnitude of the normalised global image, utils/postprocessing_loss.py:114-117
        // row by row, so that only one row of neighbours is live: sx = (c - a) + 2 (f - d) + (i - g), sy = (a + 2 b + c) - (g + 2 h + i)
        float sx[6], sy[6];
#pragma unroll
        for (int oi = -1; oi <= 1; ++oi) {
            float l[7], m[7], r[7];
            acc_pixel(acc, (size_t)((long long)idx + (long long)oi * g.W - 1), cy[oi + 1] * cx[0], l);
            acc_pixel(acc, (size_t)((long long)idx + (long long)oi * g.W + 1), cy[oi + 1] * cx[2], r);
            if (oi != 0) acc_pixel(acc, (size_t)((long long)idx + (long long)oi * g.W), cy[oi + 1] * cx[1], m);
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                if (oi == -1) { sx[c] = r[c] - l[c]; sy[c] = l[c] + 2.0f * m[c] + r[c]; }
                else if (oi == 0) sx[c] = sx[c] + 2.0f * (r[c] - l[c]);
                else { sx[c] = sx[c] + (r[c] - l[c]); sy[c] = sy[c] - (l[c] + 2.0f * m[c] + r[c]); }
            }
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            dgi[c] = sqrtf(sx[c] * sx[c] + sy[c] * sy[c] + 1e-8f);
            dgt[c] = deri[((((size_t)b * 2 + c / 3) * (g.H - 2) + (y - 1)) * (g.W - 2) + (x - 1)) * 3 + c % 3];
        }
    }
    float ny[6], gt[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t o = (((size_t)b * 2 + m) * HW + p) * 3 + c;          // dataset-native [B,2,H,W,3]
            ny[3 * m + c] = img_ny[o];
            if (!same_gt) gt[3 * m + c] = img_gt[o];
        }
    const float bd = log2f(bndry_dist[idx] + 1.0f);                           // global_training.py:118
    const size_t PS = (size_t)Btot * HW * 4;                                  // plane stride in floats
    float4* t4 = reinterpret_cast<float4*>(T + idx * 4);
    const size_t P4 = PS / 4;                                                 // ... in float4s
    t4[0] = make_float4(ny[0], ny[1], ny[2], ny[3]);                          // values 0..3
    if (same_gt) {
        *reinterpret_cast<float2*>(t4 + P4) = make_float2(ny[4], ny[5]);      // values 4, 5 (6..11 are never read)
    } else {
        t4[P4] = make_float4(ny[4], ny[5], gt[0], gt[1]);                     // values 4..7
        t4[2 * P4] = make_float4(gt[2], gt[3], gt[4], gt[5]);                 // values 8..11
    }
    t4[3 * P4] = make_float4(v[0], v[1], v[2], v[3]);                         // global image 12..15
    t4[4 * P4] = make_float4(v[4], v[5], v[6], bd);                           // 16, 17, global boundary 18, log2(bndry_dist + 1) 19
    t4[5 * P4] = make_float4(dgt[0], dgt[1], dgt[2], dgt[3]);                 // 20..23
    t4[6 * P4] = make_float4(dgt[4], dgt[5], dgi[0], dgi[1]);                 // 24..27
    t4[7 * P4] = make_float4(dgi[2], dgi[3], dgi[4], dgi[5]);                 // 28..31
    T[t_off(PS, idx, T_ZG)] = bndry_depth[idx];
    if (gimg) {
#pragma unroll
        for (int c = 0; c < 6; ++c) gimg[((size_t)b * 6 + c) * HW + p] = v[c];
    }
    if (gbnd) gbnd[idx] = v[6];
}

// ---------------------------------------------------------------------------------------------------
struct Slot {
    int q, i, j;
    bool valid, interior;
};

// Single launch for the whole step (local_training.py:99-108 calls it on 64 patches): the CTA builds its patch record itself
// (a.l_est != nullptr: no be_setup_kernel launch, no table in HBM), and the last CTA to finish - an atomic ticket - adds the CTAs'
// partial sums in fixed order and writes terms and loss (no be_loss_reduce_kernel launch).  One CTA per patch, at most a few hundred
// CTAs: occupancy is irrelevant, so the kernel takes the registers it wants (no spills).
__global__ void __launch_bounds__(BE_THREADS, 1) be_local_loss_kernel(const BeLossArgs a, const BeLocalTail tail) {
    constexpr int NIMG = 1;
    constexpr int NCH = 3 * NIMG;
    constexpr int RRMAX = BE_MAX_R * BE_MAX_R;

    __shared__ __align__(16) float s_rec[2][BE_REC];
    __shared__ __align__(16) float s_grec[2][BE_GREC];
    __shared__ float s_axis[BE_MAX_R + 3];
    __shared__ float s_part[BE_WARPS][16];
    __shared__ float s_part3[BE_WARPS][16];
    __shared__ float s_col[9], s_V[9], s_S[6];
    __shared__ double s_minv[6];
    __shared__ float4 s_Pa[RRMAX];
    __shared__ float2 s_Pb[RRMAX];
    __shared__ float4 s_gxa[RRMAX], s_gya[RRMAX];
    __shared__ float2 s_gxb[RRMAX], s_gyb[RRMAX];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const BeGeom g = a.g;
    const int R = g.R, RR = R * R;

    int blk = blockIdx.x;
    const int run = blk % a.runs_per_row; blk /= a.runs_per_row;
    const int py = blk % g.Hp;
    const int b = blk / g.Hp;
    const int px0 = run * a.G;
    const int n = min(a.G, g.Wp - px0);
    const size_t patch0 = ((size_t)b * g.Hp + py) * g.Wp + px0;
    const int np = 10;

    if (tid < R) s_axis[tid] = be_axis(tid, R);
    if (a.l_est != nullptr) {          // raw LocalStage output -> record, in place of be_setup_kernel (one patch per CTA: n == 1)
        if (tid == 0) {
            float p[12];
#pragma unroll
            for (int q = 0; q < 12; ++q) p[q] = (q < 10) ? __ldg(a.l_est + patch0 * 10 + q) : 0.0f;
            BePatch P;
            BePatchGrad G;
            be_patch_setup(p, BE_PARAMS_LOCALRAW10, tail.cam, P);
            be_patch_grad_setup(p, BE_PARAMS_LOCALRAW10, tail.cam, P, G);
            float4* rec = reinterpret_cast<float4*>(s_rec[0]);
            rec[0] = make_float4(P.sn[0], P.sn[1], P.sn[2], P.sn[3]);
            rec[1] = make_float4(P.cs[0], P.cs[1], P.cs[2], P.cs[3]);
            rec[2] = make_float4(P.vx[0], P.vx[1], P.vy[0], P.vy[1]);
            rec[3] = make_float4(P.flip[0], P.flip[1], P.z[0], P.z[1]);
            rec[4] = make_float4(P.inv_eta[0], P.inv_eta[1], P.inv_eta[2], P.inv_eta[3]);
            float4* gr = reinterpret_cast<float4*>(s_grec[0]);
            gr[0] = make_float4(G.deta_dcoef[0], G.deta_dcoef[1], G.deta_dcoef[2], G.deta_dcoef[3]);
            gr[1] = make_float4(G.dz_deta[0], G.dz_deta[1], G.dz_deta[2], G.dz_deta[3]);
            gr[2] = make_float4(G.xy_scale, G.ang_scale, 0.f, 0.f);
            if (tail.est_wrapped != nullptr) {      // the reference wraps the angles of the network output in place (local_training.py:33)
#pragma unroll
                for (int q = 4; q < 8; ++q) tail.est_wrapped[patch0 * 10 + q] = be_wrap_2pi(p[q]);
            }
        }
    } else {
        if (tid < 8) reinterpret_cast<float4*>(s_rec[0])[tid] = __ldg(reinterpret_cast<const float4*>(a.table + patch0 * BE_REC) + tid);
        if (tid >= 8 && tid < 8 + BE_GREC / 4)
            reinterpret_cast<float4*>(s_grec[0])[tid - 8] = __ldg(reinterpret_cast<const float4*>(a.gtable + patch0 * BE_GREC) + (tid - 8));
    }

    Slot sl[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int q = tid + s * BE_THREADS;
        sl[s].valid = q < RR;
        sl[s].q = sl[s].valid ? q : 0;
        sl[s].i = sl[s].q / R;
        sl[s].j = sl[s].q % R;
        sl[s].interior = sl[s].valid && sl[s].i >= 1 && sl[s].i <= R - 2 && sl[s].j >= 1 && sl[s].j <= R - 2;
        if (q < RRMAX) {   // border entries of the Sobel-gradient planes stay zero for the whole kernel
            s_gxa[q] = s_gya[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            s_gxb[q] = s_gyb[q] = make_float2(0.f, 0.f);
        }
    }
    __syncthreads();
    const float Y[2] = {s_axis[sl[0].i], s_axis[sl[1].i]};
    const float X[2] = {s_axis[sl[0].j], s_axis[sl[1].j]};

    float lossacc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

    for (int k = 0; k < n; ++k) {
        const int cur = k & 1;
        float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (warp == 0 && k + 1 < n && a.l_est == nullptr) {
            if (lane < 8) nxt = __ldg(reinterpret_cast<const float4*>(a.table + (patch0 + k + 1) * BE_REC) + lane);
            else if (lane < 8 + BE_GREC / 4) nxt = __ldg(reinterpret_cast<const float4*>(a.gtable + (patch0 + k + 1) * BE_GREC) + (lane - 8));
        }
        BePatch P;
        {
            const float4* q4 = reinterpret_cast<const float4*>(s_rec[cur]);
            const float4 r0 = q4[0], r1 = q4[1], r2 = q4[2], r3 = q4[3], r4 = q4[4];
            P.sn[0] = r0.x; P.sn[1] = r0.y; P.sn[2] = r0.z; P.sn[3] = r0.w;
            P.cs[0] = r1.x; P.cs[1] = r1.y; P.cs[2] = r1.z; P.cs[3] = r1.w;
            P.vx[0] = r2.x; P.vx[1] = r2.y; P.vy[0] = r2.z; P.vy[1] = r2.w;
            P.flip[0] = r3.x; P.flip[1] = r3.y; P.z[0] = r3.z; P.z[1] = r3.w;
            P.inv_eta[0] = r4.x; P.inv_eta[1] = r4.y; P.inv_eta[2] = r4.z; P.inv_eta[3] = r4.w;
        }
        auto ld_ny = [&](int s, float* y) {
            const float* p = a.l_ny + (((size_t)b * R + sl[s].i) * R + sl[s].j) * 3;
            y[0] = __ldg(p); y[1] = __ldg(p + 1); y[2] = __ldg(p + 2);
        };

        // ---------------- stage 1: phase 1 ----------------
        float d1[2], d2[2], h[2][2 * NIMG];
        {
            float sums[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) sums[q] = 0.0f;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                be_pixel_dists(P, X[s], Y[s], g.w, &d1[s], &d2[s]);
                if (sl[s].valid) {
                    float y[NCH];
                    ld_ny(s, y);
#pragma unroll
                    for (int m = 0; m < NIMG; ++m) {
                        const float h1 = be_h(d1[s], P.inv_eta[2 * m]), h2 = be_h(d2[s], P.inv_eta[2 * m + 1]);
                        h[s][2 * m] = h1; h[s][2 * m + 1] = h2;
                        float u[3];
                        be_wedges(h1, h2, u);
                        sums[0] = fmaf(u[0], u[0], sums[0]); sums[1] = fmaf(u[0], u[1], sums[1]); sums[2] = fmaf(u[0], u[2], sums[2]);
                        sums[3] = fmaf(u[1], u[1], sums[3]); sums[4] = fmaf(u[1], u[2], sums[4]); sums[5] = fmaf(u[2], u[2], sums[5]);
#pragma unroll
                        for (int wd = 0; wd < 3; ++wd)
#pragma unroll
                            for (int c = 0; c < 3; ++c) sums[6 + 3 * wd + c] = fmaf(u[wd], y[3 * m + c], sums[6 + 3 * wd + c]);
                    }
                }
            }
            const float tot = warp_reduce16(sums, lane);
            if (!(lane & 1)) s_part[warp][lane >> 1] = tot;
        }
        __syncthreads();   // (1)

        // ---------------- stage 2: ridge solve (warp 0, fp64) ----------------
        if (warp == 0) {
            float t = 0.0f;
            if (lane < 16) {
#pragma unroll
                for (int wv = 0; wv < BE_WARPS; ++wv) t += s_part[wv][lane];
            }
            float S[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) S[q] = __shfl_sync(FULL, t, q);
            double Minv[6];
            float C[9];
            be_solve_colors(S, g.lam, Minv, C);
            if (lane == 0) {   // static indices only: a lane-indexed register array would live in local memory
#pragma unroll
                for (int q = 0; q < 9; ++q) s_col[q] = C[q];
#pragma unroll
                for (int q = 0; q < 6; ++q) s_minv[q] = Minv[q];
            }
            if (k + 1 < n) {
                if (lane < 8) reinterpret_cast<float4*>(s_rec[cur ^ 1])[lane] = nxt;
                else if (lane < 8 + BE_GREC / 4) reinterpret_cast<float4*>(s_grec[cur ^ 1])[lane - 8] = nxt;
            }
        }
        __syncthreads();   // (2)

        // ---------------- stage 3: render, direct dL/dP ----------------
        float C[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) C[q] = s_col[q];
        float G[2][NCH];
        float bdv[2] = {0.f, 0.f};
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (sl[s].valid) {
                float Pv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int m = 0; m < NIMG; ++m) {
                    float u[3];
                    be_wedges(h[s][2 * m], h[s][2 * m + 1], u);
#pragma unroll
                    for (int c = 0; c < 3; ++c) Pv[3 * m + c] = fmaf(u[0], C[c], fmaf(u[1], C[3 + c], u[2] * C[6 + c]));
                }
                s_Pa[sl[s].q] = make_float4(Pv[0], Pv[1], Pv[2], Pv[3]);
                s_Pb[sl[s].q] = make_float2(Pv[4], Pv[5]);
                {
                    const float* p = a.l_gt + (((size_t)b * R + sl[s].i) * R + sl[s].j) * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float e1 = Pv[c] - __ldg(p + c);
                        lossacc[0] = fmaf(e1, e1, lossacc[0]);
                        G[s][c] = 2.0f * a.kc * e1;
                    }
                    bdv[s] = __ldg(a.l_bd + ((size_t)b * R + sl[s].i) * R + sl[s].j);
                }
            }
        }
        __syncthreads();   // (3)

        // ---------------- stage 4: Sobel magnitude of the rendered patch, its loss and gradient ----------------
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (sl[s].interior) {
                float sx[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, sy[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int oi = -1; oi <= 1; ++oi)
#pragma unroll
                    for (int oj = -1; oj <= 1; ++oj) {
                        if (oi == 0 && oj == 0) continue;
                        const float wx = (float)(((oi == 0) ? 2 : 1) * oj);       // sobel_x[oi+1][oj+1]
                        const float wy = (float)(-oi * ((oj == 0) ? 2 : 1));      // sobel_y[oi+1][oj+1]
                        const int qn = sl[s].q + oi * R + oj;
                        const float4 pa = s_Pa[qn];
                        const float2 pb = s_Pb[qn];
                        const float pv[6] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y};
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            if (wx != 0.0f) sx[c] = fmaf(wx, pv[c], sx[c]);
                            if (wy != 0.0f) sy[c] = fmaf(wy, pv[c], sy[c]);
                        }
                    }
                float dgt[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                {
                    const float* p = a.l_deri + (((size_t)b * (R - 2) + sl[s].i - 1) * (R - 2) + sl[s].j - 1) * 3;
                    dgt[0] = __ldg(p); dgt[1] = __ldg(p + 1); dgt[2] = __ldg(p + 2);
                }
                float gx[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gy[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const float v = fmaf(sx[c], sx[c], fmaf(sy[c], sy[c], 1e-8f));
                    const float ir = rsqrtf(v);
                    const float mag = v * ir;
                    const float e1 = mag - dgt[c];
                    lossacc[3] = fmaf(e1, e1, lossacc[3]);
                    const float gm = 2.0f * a.ks * e1;
                    gx[c] = gm * sx[c] * ir;
                    gy[c] = gm * sy[c] * ir;
                }
                s_gxa[sl[s].q] = make_float4(gx[0], gx[1], gx[2], gx[3]);
                s_gxb[sl[s].q] = make_float2(gx[4], gx[5]);
                s_gya[sl[s].q] = make_float4(gy[0], gy[1], gy[2], gy[3]);
                s_gyb[sl[s].q] = make_float2(gy[4], gy[5]);
            }
        }
        __syncthreads();   // (4)

        // ---------------- stage 5: Sobel adjoint into G, then A^T G ----------------
        {
            float sums[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) sums[q] = 0.0f;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (sl[s].valid) {
#pragma unroll
                    for (int di = -1; di <= 1; ++di)
#pragma unroll
                        for (int dj = -1; dj <= 1; ++dj) {
                            if (di == 0 && dj == 0) continue;
                            const int io = sl[s].i + di, jo = sl[s].j + dj;
                            if (io < 0 || io >= R || jo < 0 || jo >= R) continue;
                            const float wx = (float)(-dj * ((di == 0) ? 2 : 1));   // weight of gx(i+di, j+dj) in dL/dP(i,j)
                            const float wy = (float)(di * ((dj == 0) ? 2 : 1));    // weight of gy(i+di, j+dj)
                            const int qn = sl[s].q + di * R + dj;
                            if (wx != 0.0f) {
                                const float4 ga = s_gxa[qn];
                                const float2 gb = s_gxb[qn];
                                const float gv[6] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y};
#pragma unroll
                                for (int c = 0; c < NCH; ++c) G[s][c] = fmaf(wx, gv[c], G[s][c]);
                            }
                            if (wy != 0.0f) {
                                const float4 ga = s_gya[qn];
                                const float2 gb = s_gyb[qn];
                                const float gv[6] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y};
#pragma unroll
                                for (int c = 0; c < NCH; ++c) G[s][c] = fmaf(wy, gv[c], G[s][c]);
                            }
                        }
#pragma unroll
                    for (int m = 0; m < NIMG; ++m) {
                        float u[3];
                        be_wedges(h[s][2 * m], h[s][2 * m + 1], u);
#pragma unroll
                        for (int wd = 0; wd < 3; ++wd)
#pragma unroll
                            for (int c = 0; c < 3; ++c) sums[3 * wd + c] = fmaf(u[wd], G[s][3 * m + c], sums[3 * wd + c]);
                    }
                }
            }
            const float tot = warp_reduce16(sums, lane);
            if (!(lane & 1)) s_part[warp][lane >> 1] = tot;
        }
        __syncthreads();   // (5)

        // ---------------- stage 6: second solve (warp 0) ----------------
        if (warp == 0) {
            float t = 0.0f;
            if (lane < 16) {
#pragma unroll
                for (int wv = 0; wv < BE_WARPS; ++wv) t += s_part[wv][lane];
            }
            float AtG[9], V[9], Ssym[6];
#pragma unroll
            for (int q = 0; q < 9; ++q) AtG[q] = __shfl_sync(FULL, t, q);
            double Minv[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) Minv[q] = s_minv[q];
            be_backsolve(Minv, AtG, C, V, Ssym);
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 9; ++q) s_V[q] = V[q];
#pragma unroll
                for (int q = 0; q < 6; ++q) s_S[q] = Ssym[q];
            }
        }
        __syncthreads();   // (6)

        // ---------------- stage 7: per-pixel backward -> 14 per-patch sums ----------------
        {
            float V[9], Ssym[6];
#pragma unroll
            for (int q = 0; q < 9; ++q) V[q] = s_V[q];
#pragma unroll
            for (int q = 0; q < 6; ++q) Ssym[q] = s_S[q];
            float sums[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) sums[q] = 0.0f;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (sl[s].valid) {
                    float y[NCH];
                    ld_ny(s, y);
                    float gd1 = 0.0f, gd2 = 0.0f;
#pragma unroll
                    for (int m = 0; m < NIMG; ++m) {
                        float u[3], gu[3], gh1, gh2, da, de;
                        const float h1 = h[s][2 * m], h2 = h[s][2 * m + 1];
                        be_wedges(h1, h2, u);
                        be_ridge_backward_pixel(&G[s][3 * m], &y[3 * m], u, C, V, Ssym, gu);
                        be_wedges_backward(h1, h2, gu, &gh1, &gh2);
                        be_h_grad(d1[s], P.inv_eta[2 * m], &da, &de);
                        gd1 = fmaf(gh1, da, gd1); sums[8 + 2 * m] = fmaf(gh1, de, sums[8 + 2 * m]);
                        be_h_grad(d2[s], P.inv_eta[2 * m + 1], &da, &de);
                        gd2 = fmaf(gh2, da, gd2); sums[9 + 2 * m] = fmaf(gh2, de, sums[9 + 2 * m]);
                    }
                    const float lb = be_boundary(d1[s], d2[s]);
                    const float bl = bdv[s] * lb;
                    lossacc[5] = fmaf(bl, bl, lossacc[5]);
                    const float glb = 2.0f * a.kbl * bdv[s] * bl;
                    be_boundary_backward(d1[s], d2[s], lb, glb, &gd1, &gd2);
                    be_wedge_backward(P, 0, X[s], Y[s], g.w, gd1, &sums[0]);
                    be_wedge_backward(P, 1, X[s], Y[s], g.w, gd2, &sums[4]);
                }
            }
            const float tot = warp_reduce16(sums, lane);
            if (!(lane & 1)) s_part3[warp][lane >> 1] = tot;
        }
        __syncthreads();   // (7)

        // ---------------- stage 8: chain rule to the raw parameters (warp 0) ----------------
        if (warp == 0 && a.grad != nullptr) {
            float t = 0.0f;
            if (lane < 16) {
#pragma unroll
                for (int wv = 0; wv < BE_WARPS; ++wv) t += s_part3[wv][lane];
            }
            float S[14];
#pragma unroll
            for (int q = 0; q < 14; ++q) S[q] = __shfl_sync(FULL, t, q);
            if (lane == 0) {
                const float* gr = s_grec[cur];      // deta_dcoef[4], dz_deta[4], xy_scale, ang_scale
                float* out = a.grad + (patch0 + k) * np;
                const float xs = gr[8], as = gr[9];
                out[0] = xs * S[0]; out[1] = xs * S[1]; out[2] = xs * S[4]; out[3] = xs * S[5];
                out[4] = as * (S[2] + S[3]); out[5] = as * S[3]; out[6] = as * (S[6] + S[7]); out[7] = as * S[7];
                out[8] = S[8] * gr[0]; out[9] = S[9] * gr[1];
            }
        }
    }

    // ---------------- per-CTA partial loss sums ----------------
    {
        float sums[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) sums[q] = (q < 7) ? lossacc[q] : 0.0f;
        const float tot = warp_reduce16(sums, lane);
        __syncthreads();
        if (!(lane & 1)) s_part[warp][lane >> 1] = tot;
        __syncthreads();
        if (tid < 8) {
            float t = 0.0f;
#pragma unroll
            for (int wv = 0; wv < BE_WARPS; ++wv) t += s_part[wv][tid];
            a.partials[(size_t)blockIdx.x * 8 + tid] = t;
        }
    }
    if (tail.ticket == nullptr) return;
    // ---------------- last CTA: partial sums -> terms and loss (fixed order, fp64) ----------------
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(tail.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (warp == 0) {
        double acc[3] = {0.0, 0.0, 0.0};
        for (int i = lane; i < (int)gridDim.x; i += 32)
#pragma unroll
            for (int t = 0; t < 3; ++t) acc[t] += (double)__ldcg(a.partials + (size_t)i * 8 + tail.sc.src[t]);
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) acc[t] += __shfl_xor_sync(FULL, acc[t], m);
        if (lane == 0) {
            double l = 0.0;
            for (int t = 0; t < tail.sc.nterms; ++t) {
                const double v = acc[t] * tail.sc.scale[t];
                tail.terms[t] = (float)v;
                l += (double)tail.sc.gamma[t] * v;
            }
            *tail.loss = (float)l;
            *tail.ticket = 0u;                       // ready for the next launch
        }
    }
}

// partial sums -> terms (unweighted, as the oracle's `terms`) and the weighted loss
// `true_patches` (may be NULL): patch count of the global batch as all-reduced over the ranks; sc.scale was computed for
// `assumed_patches` (local patches x world), so uneven shards are corrected here by assumed / true.
__global__ void __launch_bounds__(256) be_loss_reduce_kernel(const float* __restrict__ partials, int nblocks, BeLossScale sc,
                                                             const unsigned long long* __restrict__ mask_count,
                                                             const unsigned long long* __restrict__ true_patches, double assumed_patches,
                                                             float* __restrict__ terms, float* __restrict__ loss) {
    __shared__ double s[256][7];
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x)
#pragma unroll
        for (int t = 0; t < 7; ++t) acc[t] += (double)partials[(size_t)i * 8 + t];
#pragma unroll
    for (int t = 0; t < 7; ++t) s[threadIdx.x][t] = acc[t];
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off)
#pragma unroll
            for (int t = 0; t < 7; ++t) s[threadIdx.x][t] += s[threadIdx.x + off][t];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double l = 0.0;
        const double fix = (true_patches != nullptr) ? assumed_patches / (double)(*true_patches) : 1.0;
        for (int t = 0; t < sc.nterms; ++t) {
            const int src = sc.src[t];
            double v = s[0][src] * sc.scale[t] * fix;
            if (sc.masked[t]) v = s[0][src] / (double)(*mask_count);     // 0/0 -> NaN, as the reference (global_training.py:127)
            terms[t] = (float)v;
            l += (double)sc.gamma[t] * v;
        }
        *loss = (float)l;
    }
}

// deferred depth normaliser: grad[:, 8:12] += grad_depth / (mask count of the whole batch).  An empty mask (count 0) behaves like the
// fused path and like the reference (global_training.py:127, 0/0): the depth share of every patch is 0 * inf = NaN.
// Uneven data-parallel shards: the kernel's scales assumed `assumed_patches`; the rest of the gradient is corrected by assumed / true.
__global__ void __launch_bounds__(256) be_grad_depth_fixup_kernel(float* __restrict__ grad, const float4* __restrict__ gd,
                                                                   const unsigned long long* __restrict__ mask_count,
                                                                   const unsigned long long* __restrict__ true_patches, double assumed_patches,
                                                                   size_t npatch) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npatch) return;
    const float inv = 1.0f / (float)(*mask_count);          // inf when the batch mask is empty
    const float4 d = gd[i];
    float4* g = reinterpret_cast<float4*>(grad + i * 12);
    float4 v = g[2];
    if (true_patches != nullptr && (double)(*true_patches) != assumed_patches) {
        const float fix = (float)(assumed_patches / (double)(*true_patches));
        float4 a = g[0], b = g[1];
        a.x *= fix; a.y *= fix; a.z *= fix; a.w *= fix; b.x *= fix; b.y *= fix; b.z *= fix; b.w *= fix;
        v.x *= fix; v.y *= fix; v.z *= fix; v.w *= fix;
        g[0] = a; g[1] = b;
    }
    v.x = fmaf(d.x, inv, v.x); v.y = fmaf(d.y, inv, v.y); v.z = fmaf(d.z, inv, v.z); v.w = fmaf(d.w, inv, v.w);
    g[2] = v;
}

}  // namespace

void be_launch_grad_depth_fixup(float* grad, const float* grad_depth, const unsigned long long* mask_count, const unsigned long long* true_patches,
                                double assumed_patches, size_t npatch, cudaStream_t st) {
    be_grad_depth_fixup_kernel<<<(unsigned)((npatch + 255) / 256), 256, 0, st>>>(grad, reinterpret_cast<const float4*>(grad_depth), mask_count,
                                                                                 true_patches, assumed_patches, npatch);
    ++g_be_launches;
}

void be_launch_train_targets(const float* acc, const BeGeom& g, int b0, int nb, int Btot, const float* img_ny, const float* img_gt,
                             const float* bndry_dist, const float* deri, const float* bndry_depth, float* T, float* gimg, float* gbnd,
                             cudaStream_t st) {
    const size_t n = (size_t)nb * g.H * g.W;
    be_train_targets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(acc, g, b0, nb, Btot, img_ny, img_gt, bndry_dist, deri, bndry_depth, T,
                                                                         gimg, gbnd);
    ++g_be_launches;
}

void be_launch_loss(const BeLossArgs& a, const BeLocalTail& tail, cudaStream_t st) {
    const int grid = a.NB * a.g.Hp * a.runs_per_row;
    be_local_loss_kernel<<<grid, BE_THREADS, 0, st>>>(a, tail);
    ++g_be_launches;
}

void be_launch_loss_reduce(const float* partials, int nblocks, const BeLossScale& sc, const unsigned long long* mask_count,
                           const unsigned long long* true_patches, double assumed_patches, float* terms, float* loss, cudaStream_t st) {
    be_loss_reduce_kernel<<<1, 256, 0, st>>>(partials, nblocks, sc, mask_count, true_patches, assumed_patches, terms, loss);
    ++g_be_launches;
}

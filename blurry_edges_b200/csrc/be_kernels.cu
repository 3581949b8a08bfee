// sm_100a kernels of the Blurry-Edges render -> fold -> depth path (forward / inference side).
//
// Work decomposition:
//   be_setup_kernel      one thread per patch: parameters -> 128-byte record (sin/cos of the 4 edges, flips,
//                        etas, analytic depths).  Keeps every transcendental that is constant per patch off the
//                        critical path of the renderer.
//   be_run_kernel<MODE>  one CTA (7 warps) walks a RUN of consecutive patches of one patch row.  The R*R pixels of
//                        the sliding window are owned by fixed "slots" (row i, column residue r = x mod R), two
//                        slots per thread, so an image pixel stays with the same thread for all <= ceil(R/stride)
//                        patches that cover it inside the run: its RGB values are loaded once, and its overlap sums
//                        (the Fold of utils/postprocessing_loss.py:151-173) accumulate in REGISTERS.  Only when the
//                        pixel leaves the window are the 15 sums flushed with four 16-byte vector reductions
//                        (REDG.E.ADD.F32x4) into an interleaved [B,H,W,16] accumulator.  The unfolded
//                        [B,..,R,R,Hp,Wp] tensors of the reference are never materialised.
//                        Per patch: phase 1 (distances, soft indicators, normal-equation partial sums) ->
//                        warp transposing reduction -> fp64 3x3 ridge solve by warp 0 -> phase 2 (four renders,
//                        boundary, depth mask/map) into the register accumulators.
//   be_normalise_kernel  accumulator -> the six planar maps (divide by the closed-form cover count / depth count).
#include "be_internal.h"

long long g_be_launches = 0;

namespace {

constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------------------
// per-patch records
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) be_setup_kernel(const float* __restrict__ est, int param_mode, int npatch,
                                                       BeCam cam, float* __restrict__ table, float* __restrict__ gtable) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= npatch) return;
    const int np = (param_mode == BE_PARAMS_LOCAL10 || param_mode == BE_PARAMS_LOCALRAW10) ? 10 : 12;
    float p[12];
    for (int k = 0; k < 12; ++k) p[k] = (k < np) ? est[(size_t)n * np + k] : 0.0f;
    BePatch P;
    be_patch_setup(p, param_mode, cam, P);
    float4* rec = reinterpret_cast<float4*>(table + (size_t)n * BE_REC);
    rec[0] = make_float4(P.sn[0], P.sn[1], P.sn[2], P.sn[3]);
    rec[1] = make_float4(P.cs[0], P.cs[1], P.cs[2], P.cs[3]);
    rec[2] = make_float4(P.vx[0], P.vx[1], P.vy[0], P.vy[1]);
    rec[3] = make_float4(P.flip[0], P.flip[1], P.z[0], P.z[1]);
    rec[4] = make_float4(P.inv_eta[0], P.inv_eta[1], P.inv_eta[2], P.inv_eta[3]);
    rec[5] = make_float4(P.eta[0], P.eta[1], P.eta[2], P.eta[3]);
    rec[6] = make_float4(0.f, 0.f, 0.f, 0.f);
    rec[7] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gtable) {   // chain-rule scalars of the backward pass
        BePatchGrad G;
        be_patch_grad_setup(p, param_mode, cam, P, G);
        float4* gr = reinterpret_cast<float4*>(gtable + (size_t)n * BE_GREC);
        gr[0] = make_float4(G.deta_dcoef[0], G.deta_dcoef[1], G.deta_dcoef[2], G.deta_dcoef[3]);
        gr[1] = make_float4(G.dz_deta[0], G.dz_deta[1], G.dz_deta[2], G.dz_deta[3]);
        gr[2] = make_float4(G.xy_scale, G.ang_scale, 0.f, 0.f);
    }
}

__device__ __forceinline__ void load_record(const float* rec, BePatch& P) {
    const float4* q = reinterpret_cast<const float4*>(rec);
    const float4 a = q[0], b = q[1], c = q[2], d = q[3], e = q[4];
    P.sn[0] = a.x; P.sn[1] = a.y; P.sn[2] = a.z; P.sn[3] = a.w;
    P.cs[0] = b.x; P.cs[1] = b.y; P.cs[2] = b.z; P.cs[3] = b.w;
    P.vx[0] = c.x; P.vx[1] = c.y; P.vy[0] = c.z; P.vy[1] = c.w;
    P.flip[0] = d.x; P.flip[1] = d.y; P.z[0] = d.z; P.z[1] = d.w;
    P.inv_eta[0] = e.x; P.inv_eta[1] = e.y; P.inv_eta[2] = e.z; P.inv_eta[3] = e.w;
}

// Sum 16 per-lane values over the warp with 15 shuffles instead of 80: at every butterfly step a lane keeps half
// of its values and ships the other half.  On return lane l holds the warp total of v[l >> 1].
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

__device__ __forceinline__ float ld_img(const BeImg& im, int b, int m, int c, int y, int x) {
    return __ldg(im.p + b * im.sb + m * im.sm + c * im.sc + y * im.sy + x * im.sx);
}

// ---------------------------------------------------------------------------------------------------
// the renderer
// ---------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __maxnreg__(96) be_run_kernel(const BeRunArgs a) {
    constexpr bool INFER = (MODE == BE_RUN_INFER);
    constexpr bool TRAIN = (MODE == BE_RUN_TRAINFWD);   // global-loss forward: 2 renders + boundary -> 7 folded planes
    constexpr bool FOLD = INFER || TRAIN;
    constexpr int NIMG = FOLD ? 2 : 1;           // images whose pixels enter the normal equations
    constexpr int NACC = INFER ? 15 : (TRAIN ? 7 : 1);
    constexpr int ACCW = INFER ? BE_ACC : 8;     // floats per pixel of the fold accumulator

    __shared__ __align__(16) float s_rec[2][BE_REC];
    __shared__ float s_axis[BE_MAX_R + 3];
    __shared__ float s_part[BE_WARPS][16];
    __shared__ float s_col[12];                  // C[3*w+c] (9), inv_refoc[2]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const BeGeom g = a.g;
    const int R = g.R, RR = R * R;

    // which run is this CTA?
    int blk = blockIdx.x;
    const int run = blk % a.runs_per_row; blk /= a.runs_per_row;
    const int py = blk % g.Hp;
    const int b = blk / g.Hp;
    const int px0 = run * a.G;
    const int n = min(a.G, g.Wp - px0);
    const int y0 = py * g.stride;
    const size_t patch0 = ((size_t)b * g.Hp + py) * g.Wp + px0;

    if (tid < R) s_axis[tid] = be_axis(tid, R);
    if (tid < 8) reinterpret_cast<float4*>(s_rec[0])[tid] = __ldg(reinterpret_cast<const float4*>(a.table + patch0 * BE_REC) + tid);

    // slot bookkeeping: slot s -> (row i, residue r); local column j slides by -stride per patch
    bool valid[2];
    int si[2], j[2];
    float Y[2];
    float pix[2][3 * NIMG];
    float acc[2][NACC];
    float zg[2] = {0.0f, 0.0f};                  // TRAIN: ground-truth boundary depth at the slot's pixel
    unsigned mcount = 0;                         // TRAIN: #[z_gt != 0 and mask != 0]   (global_training.py:125-127)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int slot = tid + s * BE_THREADS;
        valid[s] = slot < RR;
        si[s] = valid[s] ? slot / R : 0;
        j[s] = valid[s] ? slot % R : 0;
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc[s][q] = 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        Y[s] = s_axis[si[s]];
        const int x = px0 * g.stride + j[s];
#pragma unroll
        for (int m = 0; m < NIMG; ++m)
#pragma unroll
            for (int c = 0; c < 3; ++c) pix[s][3 * m + c] = ld_img(a.img, b, m, c, y0 + si[s], x);
        if (TRAIN) zg[s] = __ldg(a.zgt + ((size_t)b * g.H + y0 + si[s]) * g.W + x);
    }

    const float inv_sharp = 1.0f / (BE_SQRT2_F * BE_ETA_SHARP);

    for (int k = 0; k < n; ++k) {
        const int cur = k & 1;
        float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (warp == 0 && lane < 8 && k + 1 < n)
            nxt = __ldg(reinterpret_cast<const float4*>(a.table + (patch0 + k + 1) * BE_REC) + lane);

        BePatch P;
        load_record(s_rec[cur], P);

        // ---------------- phase 1: distances, soft indicators, normal-equation partial sums ----------------
        float d1[2], d2[2], h[2][2 * NIMG];
        float sums[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) sums[q] = 0.0f;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const float X = s_axis[j[s]];
            be_pixel_dists(P, X, Y[s], g.w, &d1[s], &d2[s]);
            if (valid[s]) {
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
                        for (int c = 0; c < 3; ++c) sums[6 + 3 * wd + c] = fmaf(u[wd], pix[s][3 * m + c], sums[6 + 3 * wd + c]);
                }
                if (INFER) {
                    const int mk = be_mask(d1[s], d2[s], a.densify_w != 0);
                    sums[15] += (mk == 1) ? 1.0f : ((mk == 2) ? 1024.0f : 0.0f);   // two exact counters in one float
                }
                if (TRAIN) mcount += (be_mask(d1[s], d2[s], false) != 0 && zg[s] != 0.0f) ? 1u : 0u;
            }
        }
        const float tot = warp_reduce16(sums, lane);
        if (!(lane & 1)) s_part[warp][lane >> 1] = tot;
        __syncthreads();

        // ---------------- ridge solve (warp 0, fp64) ----------------
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
            if (FOLD && lane == 0) {   // static indices only: a lane-indexed register array would live in local memory
#pragma unroll
                for (int q = 0; q < 9; ++q) s_col[q] = C[q];
            }
            if (INFER) {
                if (lane == 9 || lane == 10) {
                    const int cnt = (int)S[15];
                    const int have = (lane == 9) ? (cnt & 1023) : (cnt >> 10);
                    const float sg = have > 0 ? be_refocus_sigma(a.cam, (lane == 9) ? P.z[0] : P.z[1]) : BE_ETA_SHARP;   // blurry_edges_test.py:66-72
                    s_col[lane] = 1.0f / (BE_SQRT2_F * sg);
                }
            }
            if (!FOLD && lane == 0) {
                // colours [NB][3(channel)][3(wedge)][Hp][Wp]  (blurry_edges_test.py:27 permute)
                float* dst = a.colors + (size_t)b * 9 * g.Hp * g.Wp + (size_t)py * g.Wp + px0 + k;
#pragma unroll
                for (int wd = 0; wd < 3; ++wd)
#pragma unroll
                    for (int c = 0; c < 3; ++c) dst[(size_t)(c * 3 + wd) * g.Hp * g.Wp] = C[3 * wd + c];
            }
            if (lane < 8 && k + 1 < n) reinterpret_cast<float4*>(s_rec[cur ^ 1])[lane] = nxt;
        }
        __syncthreads();

        // ---------------- phase 2: renders, boundary, depth -> register accumulators ----------------
        if (INFER) {
            float C[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) C[q] = s_col[q];
            const float inv_r1 = s_col[9], inv_r2 = s_col[10];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (valid[s]) {
                    float u[3];
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        be_wedges(h[s][2 * m], h[s][2 * m + 1], u);
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            acc[s][3 * m + c] += fmaf(u[0], C[c], fmaf(u[1], C[3 + c], u[2] * C[6 + c]));
                    }
                    be_wedges(be_h(d1[s], inv_sharp), be_h(d2[s], inv_sharp), u);            // blurry_edges_test.py:63-64
#pragma unroll
                    for (int c = 0; c < 3; ++c) acc[s][6 + c] += fmaf(u[0], C[c], fmaf(u[1], C[3 + c], u[2] * C[6 + c]));
                    be_wedges(be_h(d1[s], inv_r1), be_h(d2[s], inv_r2), u);                  // :73-74
#pragma unroll
                    for (int c = 0; c < 3; ++c) acc[s][9 + c] += fmaf(u[0], C[c], fmaf(u[1], C[3 + c], u[2] * C[6 + c]));
                    acc[s][12] += be_boundary(d1[s], d2[s]);                                  // :59-61
                    const int mk = be_mask(d1[s], d2[s], a.densify_w != 0);                  // :47-57
                    acc[s][13] += (mk == 1) ? P.z[0] : ((mk == 2) ? P.z[1] : 0.0f);
                    acc[s][14] += (mk > 0) ? 1.0f : 0.0f;
                }
            }
        }

        if (TRAIN) {
            float C[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) C[q] = s_col[q];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (valid[s]) {
                    float u[3];
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        be_wedges(h[s][2 * m], h[s][2 * m + 1], u);
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            acc[s][3 * m + c] += fmaf(u[0], C[c], fmaf(u[1], C[3 + c], u[2] * C[6 + c]));
                    }
                    acc[s][6] += be_boundary(d1[s], d2[s]);
                }
            }
        }

        // ---------------- slide the window by one patch ----------------
        const bool last = (k + 1 == n);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (!valid[s]) continue;
            const int jn = j[s] - g.stride;
            if (FOLD && (last || jn < 0)) {
                const int x = (px0 + k) * g.stride + j[s];
                float4* dst = reinterpret_cast<float4*>(a.acc + (((size_t)b * g.H + y0 + si[s]) * g.W + x) * ACCW);
                atomicAdd(dst + 0, make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]));
                if (INFER) {
                    atomicAdd(dst + 1, make_float4(acc[s][4], acc[s][5], acc[s][6], acc[s][7 % NACC]));
                    atomicAdd(dst + 2, make_float4(acc[s][8 % NACC], acc[s][9 % NACC], acc[s][10 % NACC], acc[s][11 % NACC]));
                    atomicAdd(dst + 3, make_float4(acc[s][12 % NACC], acc[s][13 % NACC], acc[s][14 % NACC], 0.0f));
                } else {
                    atomicAdd(dst + 1, make_float4(acc[s][4 % NACC], acc[s][5 % NACC], acc[s][6 % NACC], 0.0f));
                }
#pragma unroll
                for (int q = 0; q < NACC; ++q) acc[s][q] = 0.0f;
            }
            if (!last) {
                if (jn < 0) {
                    j[s] = jn + R;
                    const int x = (px0 + k + 1) * g.stride + j[s];
#pragma unroll
                    for (int m = 0; m < NIMG; ++m)
#pragma unroll
                        for (int c = 0; c < 3; ++c) pix[s][3 * m + c] = ld_img(a.img, b, m, c, y0 + si[s], x);
                    if (TRAIN) zg[s] = __ldg(a.zgt + ((size_t)b * g.H + y0 + si[s]) * g.W + x);
                } else {
                    j[s] = jn;
                }
            }
        }
    }
    if (TRAIN) {
        mcount = __reduce_add_sync(FULL, mcount);
        if (lane == 0 && mcount) atomicAdd(a.mask_count, (unsigned long long)mcount);
    }
}

// ---------------------------------------------------------------------------------------------------
// accumulator -> planar maps       (divisions of utils/postprocessing_loss.py:151-173, threshold blurry_edges_test.py:144)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cover_1d(int y, int R, int s, int np) {
    const int hi = min(y / s, np - 1);
    const int lo = (y - R + 1 <= 0) ? 0 : (y - R + s) / s;
    return max(hi - lo + 1, 0);
}

__global__ void __launch_bounds__(256) be_normalise_kernel(const float* __restrict__ acc, BeGeom g, int B, float thres,
                                                           float* __restrict__ image, float* __restrict__ sharp,
                                                           float* __restrict__ refoc, float* __restrict__ bndry,
                                                           float* __restrict__ depth, float* __restrict__ conf,
                                                           float* __restrict__ depth_thr) {
    const size_t HW = (size_t)g.H * g.W;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * HW) return;
    const int b = (int)(idx / HW);
    const size_t p = idx % HW;
    const int y = (int)(p / g.W), x = (int)(p % g.W);
    const float4* src = reinterpret_cast<const float4*>(acc + idx * BE_ACC);
    const float4 q0 = src[0], q1 = src[1], q2 = src[2], q3 = src[3];
    const float n = (float)(cover_1d(y, g.R, g.stride, g.Hp) * cover_1d(x, g.R, g.stride, g.Wp));
    float* im = image + (size_t)b * 6 * HW + p;
    im[0] = q0.x / n; im[HW] = q0.y / n; im[2 * HW] = q0.z / n;
    im[3 * HW] = q0.w / n; im[4 * HW] = q1.x / n; im[5 * HW] = q1.y / n;
    float* sh = sharp + (size_t)b * 3 * HW + p;
    sh[0] = q1.z / n; sh[HW] = q1.w / n; sh[2 * HW] = q2.x / n;
    float* rf = refoc + (size_t)b * 3 * HW + p;
    rf[0] = q2.y / n; rf[HW] = q2.z / n; rf[2 * HW] = q2.w / n;
    bndry[idx] = q3.x / n;
    const float cnt = q3.z;
    const float dz = q3.y / (cnt > 0.0f ? cnt : 1.0f);
    const float cf = cnt / n;
    depth[idx] = dz;
    conf[idx] = cf;
    if (depth_thr) depth_thr[idx] = (cf > thres) ? dz : 0.0f;
}

__global__ void __launch_bounds__(256) be_refold_kernel(const float* __restrict__ unf, BeGeom g, int M, float* __restrict__ img) {
    const size_t HW = (size_t)g.H * g.W;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)M * 3 * HW) return;
    const size_t mc = idx / HW;
    const size_t p = idx % HW;
    const int y = (int)(p / g.W), x = (int)(p % g.W);
    const int py = min(y / g.stride, g.Hp - 1), px = min(x / g.stride, g.Wp - 1);
    const int i = y - py * g.stride, jj = x - px * g.stride;
    float v = 0.0f;
    if (i < g.R && jj < g.R) v = unf[(((mc * g.R + i) * g.R + jj) * g.Hp + py) * g.Wp + px];
    img[idx] = v;
}

__global__ void __launch_bounds__(256) be_cover_count_kernel(BeGeom g, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.H * g.W) return;
    out[idx] = (float)(cover_1d(idx / g.W, g.R, g.stride, g.Hp) * cover_1d(idx % g.W, g.R, g.stride, g.Wp));
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
void be_launch_setup(const float* est, int param_mode, int npatch, const BeCam& cam, float* table, float* gtable, cudaStream_t st) {
    be_setup_kernel<<<(npatch + 127) / 128, 128, 0, st>>>(est, param_mode, npatch, cam, table, gtable);
    ++g_be_launches;
}

void be_launch_run(int mode, const BeRunArgs& a, cudaStream_t st) {
    const int grid = a.NB * a.g.Hp * a.runs_per_row;
    if (mode == BE_RUN_INFER) be_run_kernel<BE_RUN_INFER><<<grid, BE_THREADS, 0, st>>>(a);
    else if (mode == BE_RUN_TRAINFWD) be_run_kernel<BE_RUN_TRAINFWD><<<grid, BE_THREADS, 0, st>>>(a);
    else be_run_kernel<BE_RUN_COLORS><<<grid, BE_THREADS, 0, st>>>(a);
    ++g_be_launches;
}

void be_launch_normalise(const float* acc, const BeGeom& g, int B, float thres, float* image, float* sharp, float* refoc,
                         float* bndry, float* depth, float* conf, float* depth_thr, cudaStream_t st) {
    const size_t n = (size_t)B * g.H * g.W;
    be_normalise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(acc, g, B, thres, image, sharp, refoc, bndry, depth, conf, depth_thr);
    ++g_be_launches;
}

void be_launch_refold(const float* unfolded, const BeGeom& g, int M, float* image, cudaStream_t st) {
    const size_t n = (size_t)M * 3 * g.H * g.W;
    be_refold_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(unfolded, g, M, image);
    ++g_be_launches;
}

void be_launch_cover_count(const BeGeom& g, float* out, cudaStream_t st) {
    be_cover_count_kernel<<<(g.H * g.W + 255) / 256, 256, 0, st>>>(g, out);
    ++g_be_launches;
}

// be_loss2_kernel: global-stage loss forward + analytic backward (global_training.py:62-157).
//
// Same mathematics as the local-stage kernel of be_train.cu, restructured after the ncu captures of its predecessors
// (profiles/r1_loss_kernel_full.txt: 19 225 warp-instructions per patch, 7 CTA barriers and two serial fp64 solves per patch;
// profiles/r1j_train_kernels_full.txt: 11 293, three CTA barriers, every warp repeating the second solve):
//   * the ridge colours C and the inverse normal matrix M^-1 of every patch come from the TRAINFWD render pass that has to run
//     first anyway (its solver warp stores a 64-byte record per patch): no phase-1 sums, no first reduction, no fp64 solve here;
//   * 7 render warps (two packed pixel slots per thread, be_pack.cuh) + one HELPER warp.  Everything that crosses lanes goes to the
//     helper through shared memory: it adds up the lanes' A^T G partial sums, does the second solve V = M^-1 A^T G,
//     S = V C^T + C V^T once (not once per warp), later adds up the 14 backward sums and applies the chain rule to the raw
//     parameters, and prefetches the next patch's records.  The render warps hand over with bar.arrive (non-blocking) and wait
//     for V, S on an mbarrier, behind the part of the backward pass that does not depend on them (edge geometry, boundary and
//     depth terms); the two remaining barriers (the Sobel stencil exchanges) are named barriers of the render warps only;
//   * the kernel is bound by the L1/shared-memory data pipe, so the stencil stages move as little as the algebra allows: the
//     forward Sobel runs on the wedge weights u1, u2 of the two images (the filter is linear, P_c = C0_c + u1 (C1-C0)_c +
//     u2 (C2-C0)_c: 4 values per neighbour instead of 6 channels), the adjoint on the gradients projected onto C1-C0 and C2-C0
//     (only dL/du_k - dL/du_0 enters the backward: 8 values instead of 12), and the adjoint's share of A^T G is accumulated
//     locally (sum_q gx_c(q) Sobel_x(u_w)(q) + gy_c(q) Sobel_y(u_w)(q)).  Neighbours are exchanged in PAIR layout (see s_X below)
//     so that the stencil FMAs are packed too; R = 21 is a template constant (neighbour offsets become immediates); targets
//     are addressed with 32-bit offsets (no spills at 128 registers).
#include "be_internal.h"
#include "be_pack.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int NT = BE_THREADS;               // 224 render threads, 7 warps, two pixel slots per thread
constexpr int NTHR = NT + 32;                // + helper warp
constexpr int HALO = BE_MAX_R + 1;
constexpr int NE = NT + 2 * HALO;
constexpr int NPLANE = 6;                    // stencil exchange planes: 2 (wedge weights) + 4 (projected Sobel gradients)
#ifndef BE_LOSS2_PAD_SMEM
#define BE_LOSS2_PAD_SMEM 0            // diagnosis only: extra dynamic shared memory, e.g. 65536 forces one CTA per SM
#endif
constexpr size_t DYN_SMEM = sizeof(float4) * (NPLANE * NE + BE_WARPS * 32 * 3 + BE_WARPS * 32 * 4 + 2 * NT) + BE_LOSS2_PAD_SMEM;

template <int ID, int COUNT> __device__ __forceinline__ void bar_sync_id() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
template <int ID, int COUNT> __device__ __forceinline__ void bar_arrive_id() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
constexpr int BAR_RENDER = 1, BAR_ATG = 2, BAR_SUMS = 3;      // 0 = __syncthreads

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    unsigned ok;
    do {
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    } while (!ok);
}

// Read-only loads the compiler may not merge with an earlier load of the same address: when img_gt is img_ny, stage D re-reads
// the values stage A has already used instead of keeping them in registers across the two stencil stages.
// Diagnosis builds (results wrong on purpose; -DBE_LOSS2_DIAG=n via BE_NVCC_EXTRA): 1 = every patch of a run reads the first patch's
// target window (L1 hits), 2 = no target loads at all, 3 = no stencil loads from shared memory, 4 = no stencil barriers.
#ifndef BE_LOSS2_DIAG
#define BE_LOSS2_DIAG 0
#endif
#if BE_LOSS2_DIAG == 2
#define BE_TLD4(p) make_float4(0.25f, 0.5f, 0.75f, 0.125f)
#define BE_TLD2(p) make_float2(0.25f, 0.5f)
#define BE_TLD1(p) 0.9f
#else
#define BE_TLD4(p) __ldg(reinterpret_cast<const float4*>(p))
#define BE_TLD2(p) __ldg(reinterpret_cast<const float2*>(p))
#define BE_TLD1(p) __ldg(p)
#endif
#if BE_LOSS2_DIAG == 3
#define BE_SLD4(p) make_float4(0.25f, 0.5f, 0.75f, 0.125f)
#else
#define BE_SLD4(p) (*(p))
#endif

__device__ __forceinline__ float4 ldg_again4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldg_again2(const float* p) {
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

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

// RCT = 21: patch size known at compile time (neighbour offsets become immediates), 0: generic.  SAMEGT: img_gt is img_ny (the
// training call of the reference, global_training.py:210): the colour term reads the NY values and the GT values stay out of L1.
template <int RCT, bool SAMEGT>
__global__ void __launch_bounds__(NTHR, 2) be_loss2_kernel(const BeLossArgs a) {
    __shared__ __align__(16) float s_rec[2][BE_REC];
    __shared__ __align__(16) float s_grec[2][BE_GREC];
    __shared__ __align__(16) float s_crec[2][BE_CREC];
    __shared__ float s_axis[BE_MAX_R + 3];
    __shared__ float s_part[BE_WARPS][16];          // end of kernel: loss partial sums per warp
    __shared__ __align__(16) float s_VS[12];        // (V_k - V_0)[3], (S_k - S_0)[3] for k = 1, 2 of the current patch (helper -> render warps)
    __shared__ unsigned long long s_vready;         // mbarrier: s_VS published
    // Stencil exchange planes in PAIR layout: entry e holds, per channel, (value at pixel e-HALO, value at pixel e-HALO+NT), so the
    // neighbour of both slots of thread t at offset `off` is the one entry t+HALO+off (3 LDS.128 for 6 channels of two pixels).
    // Pixels within HALO of the seam are written twice (as the low half of their own entry and as the high half of the entry NT
    // below); entries outside the patch stay zero.  Row wrap-around needs no mask: a wrapped neighbour is a border pixel, whose
    // Sobel gradients are zero, and only interior pixels (which never wrap) use the rendered-patch plane.
    // 57 KB of dynamic shared memory (above the 48 KB static limit).  Two CTAs need a 132 KB carve-out, which leaves 124 KB of L1
    // for the packed targets: one patch reads 441 x 144 B = 63 KB of them and shares 19 of its 21 columns with the next patch.
    extern __shared__ float4 s_dyn[];
    float4* const s_X = s_dyn;                    // [NPLANE][NE]
    float4* const s_P2 = s_X;                     // [2][NE] wedge weights of the rendered patch: (u1, u2) image 1, (u1, u2) image 2
    float4* const s_G2 = s_X + 2 * NE;            // [4][NE] projected Sobel gradients: PX (img 1), PX (img 2), PY (img 1), PY (img 2)
    float4* const s_partA = s_X + NPLANE * NE;    // [7*32][3] per-lane A^T G partial sums (9 of 12 floats used), rows of 48 bytes
    float4* const s_partB = s_partA + BE_WARPS * 32 * 3;   // [7*32][4] per-lane backward sums (14 of 16 used), rows of 64 bytes, swizzled
    float4* const s_stash0 = s_partB + BE_WARPS * 32 * 4;  // [NT] thread-private (d1, d2) pairs, stage A -> D
    float4* const s_stash1 = s_stash0 + NT;                // [NT] thread-private (global boundary, bndry_dist) pairs

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
    const int np = 12;

    // per-patch records: lanes 0-7 table, 8-10 gtable, 11-14 crec
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
    if (tid == NTHR - 1) mbar_init(&s_vready, 1);
    for (int i = tid; i < NPLANE * NE; i += NTHR) s_X[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    if (warp == BE_WARPS) {
        // =========================================== helper warp ===========================================
        for (int k = 0; k < n; ++k) {
            const int cur = k & 1;
            float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane < NFETCH && k + 1 < n) nxt = fetch(patch0 + k + 1, lane);
            bar_sync_id<BAR_ATG, NTHR>();                  // A^T G partial sums of patch k are in s_partA
            if (lane < NFETCH && k + 1 < n) stash(cur ^ 1, lane, nxt);       // nobody reads the other record buffer any more
            const int gq = lane >> 2, qq = lane & 3;
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (qq < 3) {
#pragma unroll
                for (int wv = 0; wv < BE_WARPS; ++wv)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float4 v = s_partA[(wv * 32 + gq + 8 * jj) * 3 + qq];
                        t4.x += v.x; t4.y += v.y; t4.z += v.z; t4.w += v.w;
                    }
            }
#pragma unroll
            for (int m = 4; m <= 16; m <<= 1) {
                t4.x += __shfl_xor_sync(FULL, t4.x, m); t4.y += __shfl_xor_sync(FULL, t4.y, m);
                t4.z += __shfl_xor_sync(FULL, t4.z, m); t4.w += __shfl_xor_sync(FULL, t4.w, m);
            }
            float AtG[9];
            AtG[0] = __shfl_sync(FULL, t4.x, 0); AtG[1] = __shfl_sync(FULL, t4.y, 0); AtG[2] = __shfl_sync(FULL, t4.z, 0);
            AtG[3] = __shfl_sync(FULL, t4.w, 0); AtG[4] = __shfl_sync(FULL, t4.x, 1); AtG[5] = __shfl_sync(FULL, t4.y, 1);
            AtG[6] = __shfl_sync(FULL, t4.z, 1); AtG[7] = __shfl_sync(FULL, t4.w, 1); AtG[8] = __shfl_sync(FULL, t4.x, 2);
            {   // second solve of the ridge backward: V = M^-1 (A^T G), S = V C^T + C V^T (be_backsolve, fp32)
                float C[9], Mi[6], V[9], Ssym[6];
#pragma unroll
                for (int i = 0; i < 9; ++i) C[i] = s_crec[cur][i];
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
                if (lane == 0) {   // V_k - V_0 and S_k - S_0 (k = 1, 2): all the per-pixel backward needs (u_0 = 1 - u_1 - u_2)
                    float4* o = reinterpret_cast<float4*>(s_VS);
                    o[0] = make_float4(V[3] - V[0], V[4] - V[1], V[5] - V[2], V[6] - V[0]);
                    o[1] = make_float4(V[7] - V[1], V[8] - V[2], Ssym[1] - Ssym[0], Ssym[3] - Ssym[1]);
                    o[2] = make_float4(Ssym[4] - Ssym[2], Ssym[2] - Ssym[0], Ssym[4] - Ssym[1], Ssym[5] - Ssym[2]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_vready);
            bar_sync_id<BAR_SUMS, NTHR>();                 // the 14 backward sums of patch k are in s_partB
            t4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int wv = 0; wv < BE_WARPS; ++wv)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int l = gq + 8 * jj;
                    const float4 v = s_partB[(wv * 32 + l) * 4 + ((qq + (l >> 1)) & 3)];
                    t4.x += v.x; t4.y += v.y; t4.z += v.z; t4.w += v.w;
                }
#pragma unroll
            for (int m = 4; m <= 16; m <<= 1) {
                t4.x += __shfl_xor_sync(FULL, t4.x, m); t4.y += __shfl_xor_sync(FULL, t4.y, m);
                t4.z += __shfl_xor_sync(FULL, t4.z, m); t4.w += __shfl_xor_sync(FULL, t4.w, m);
            }
            float S[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                S[4 * q + 0] = __shfl_sync(FULL, t4.x, q); S[4 * q + 1] = __shfl_sync(FULL, t4.y, q);
                S[4 * q + 2] = __shfl_sync(FULL, t4.z, q); S[4 * q + 3] = __shfl_sync(FULL, t4.w, q);
            }
            // chain rule of global_training.py:141-145 backward: 14 per-patch sums -> 12 raw-parameter gradients
            if (lane == 0 && a.grad != nullptr) {
                const float* gr = s_grec[cur];      // deta_dcoef[4], dz_deta[4], xy_scale, ang_scale
                const float xs = gr[8], as = gr[9];
                float4 o0, o1, o2;
                o0.x = xs * S[0]; o0.y = xs * S[1]; o0.z = xs * S[4]; o0.w = xs * S[5];
                o1.x = as * (S[2] + S[3]); o1.y = as * S[3]; o1.z = as * (S[6] + S[7]); o1.w = as * S[7];
                const float4 zd = make_float4(S[12] * gr[4] * gr[0], S[13] * gr[6] * gr[1], S[12] * gr[5] * gr[2], S[13] * gr[7] * gr[3]);
                const bool defer = a.defer_depth != 0;                 // depth share kept apart until the batch mask count is known
                // empty batch mask: the reference's depth term is 0/0 and autograd hands NaN to every eta coefficient
                // (global_training.py:127); the deferred path gets the same from 0 * (1/0) in be_grad_depth_fixup_kernel
                const float empty = (!defer && *a.mask_count == 0ull) ? __int_as_float(0x7fc00000) : 0.0f;
                o2.x = S[8] * gr[0] + (defer ? 0.0f : zd.x + empty);
                o2.y = S[9] * gr[1] + (defer ? 0.0f : zd.y + empty);
                o2.z = S[10] * gr[2] + (defer ? 0.0f : zd.z + empty);
                o2.w = S[11] * gr[3] + (defer ? 0.0f : zd.w + empty);
                float4* o4 = reinterpret_cast<float4*>(a.grad + (patch0 + k) * np);     // 48-byte rows: 16-byte aligned
                o4[0] = o0; o4[1] = o1; o4[2] = o2;
                if (defer) reinterpret_cast<float4*>(a.grad_depth)[patch0 + k] = zd;
            }
        }
        __syncthreads();      // matches the render warps' barrier before the loss partial sums
        __syncthreads();
        return;
    }

    // =========================================== render warps ===========================================
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
    const f2 Y = mk2(s_axis[pi[0]], s_axis[pi[1]]), X = mk2(s_axis[pj[0]], s_axis[pj[1]]);
    const float vm0 = (RCT == BE_MAX_R || valid[0]) ? 1.0f : 0.0f, vm1 = valid[1] ? 1.0f : 0.0f;   // R = 21: every thread has a low pixel
    // depth term: gamma_d / (mask count of the whole batch); with a deferred normaliser (a.grad_depth) the count is applied later
    const float kd = a.defer_depth ? a.gamma_d : a.gamma_d / (float)(*a.mask_count);
    const unsigned TPS = (unsigned)a.NBT * g.H * g.W * 4;   // floats between consecutive float4 planes of T (32-bit offsets: be_launch_loss2 checks the size)
    const size_t bT = (size_t)(b + a.b0);                   // pair index inside the targets of the whole batch
    const unsigned toff0[2] = {(unsigned)(((bT * g.H + y0 + pi[0]) * g.W + pj[0]) * 4), (unsigned)(((bT * g.H + y0 + pi[1]) * g.W + pj[1]) * 4)};
    const float k2c = 2.0f * a.kc, k2cc = 2.0f * a.kcc;

    float lossacc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto fold = [&](f2 v) { return (RCT == BE_MAX_R) ? fmaf(hi(v), vm1, lo(v)) : fmaf(hi(v), vm1, lo(v) * vm0); };       // both slots of a thread, padding slots dropped

    for (int k = 0; k < n; ++k) {
        const int cur = k & 1;
        const unsigned x0 = (unsigned)((px0 + k) * g.stride);
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
#if BE_LOSS2_DIAG == 1
        const unsigned tp[2] = {toff0[0] + 4u * (unsigned)(px0 * g.stride), toff0[1] + 4u * (unsigned)(px0 * g.stride)};
#else
        const unsigned tp[2] = {toff0[0] + 4u * x0, toff0[1] + 4u * x0};      // element offset of plane 0 of each slot's pixel in the packed targets
#endif

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
                if (SAMEGT) {                            // gt = ny: values 0..5 (plane 0 and the first half of plane 1)
                    t2[s] = BE_TLD4(a.T + tp[s]);
                    t1[s] = BE_TLD2(a.T + (tp[s] + TPS));
                } else {                                 // values 6..11 (second half of plane 1 and plane 2)
                    t1[s] = BE_TLD2(a.T + (tp[s] + TPS + 2u));
                    t2[s] = BE_TLD4(a.T + (tp[s] + 2u * TPS));
                }
                t3[s] = BE_TLD4(a.T + (tp[s] + 3u * TPS));
                t4[s] = BE_TLD4(a.T + (tp[s] + 4u * TPS));
            }
            f2 d1, d2;
            be_pixel_dists2(P, X, Y, g.w, &d1, &d2);
            f2 Pv[6], U[4];                           // U = (u1, u2) of image 1, (u1, u2) of image 2
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                h[2 * m] = be_h2(d1, P.inv_eta[2 * m]);
                h[2 * m + 1] = be_h2(d2, P.inv_eta[2 * m + 1]);
                const f2 gg = sub2(bc2(1.0f), h[2 * m + 1]);
                const f2 u0 = mul2(sub2(bc2(1.0f), h[2 * m]), gg), u1 = mul2(h[2 * m], gg), u2 = h[2 * m + 1];
#pragma unroll
                for (int c = 0; c < 3; ++c) Pv[3 * m + c] = fma2(u0, bc2(C[c]), fma2(u1, bc2(C[3 + c]), mul2(u2, bc2(C[6 + c]))));
                U[2 * m] = u1; U[2 * m + 1] = u2;
            }
            // The Sobel filter is linear and P_c = C0_c + u1 (C1_c - C0_c) + u2 (C2_c - C0_c): the stencil stage needs the neighbours'
            // u1, u2 of the two images (4 values), not their 6 rendered channels - a third less shared-memory traffic in stage B.
            store_pairs(s_P2, U, 4);
            float gb_[2], bd_[2], Gs[2][6];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                float pv[6];
#pragma unroll
                for (int c = 0; c < 6; ++c) pv[c] = s ? hi(Pv[c]) : lo(Pv[c]);
                const float gt[6] = {SAMEGT ? t2[s].x : t1[s].x, SAMEGT ? t2[s].y : t1[s].y, SAMEGT ? t2[s].z : t2[s].x,
                                     SAMEGT ? t2[s].w : t2[s].y, SAMEGT ? t1[s].x : t2[s].z, SAMEGT ? t1[s].y : t2[s].w};
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
            s_stash0[tid] = make_float4(lo(d1), hi(d1), lo(d2), hi(d2));
            s_stash1[tid] = make_float4(gb_[0], gb_[1], bd_[0], bd_[1]);
        }
#if BE_LOSS2_DIAG != 4
        bar_sync_id<BAR_RENDER, NT>();   // (X1) rendered patch visible to the render warps
#endif

        // ---------------- stage B: Sobel magnitude of the rendered patch, its loss and gradient (both slots packed) ----------------
        {
            float4 t6[2], t7[2], t8[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                t6[s] = BE_TLD4(a.T + (tp[s] + 5u * TPS));      // values 20..31: derivative targets
                t7[s] = BE_TLD4(a.T + (tp[s] + 6u * TPS));
                t8[s] = BE_TLD4(a.T + (tp[s] + 7u * TPS));
            }
            f2 ux[4], uy[4];                           // Sobel responses of (u1, u2) of image 1 and of image 2
#pragma unroll
            for (int c = 0; c < 4; ++c) ux[c] = uy[c] = bc2(0.0f);
#pragma unroll
            for (int oi = -1; oi <= 1; ++oi)
#pragma unroll
                for (int oj = -1; oj <= 1; ++oj) {
                    if (oi == 0 && oj == 0) continue;
                    const float wx = (float)(((oi == 0) ? 2 : 1) * oj);       // sobel_x[oi+1][oj+1]
                    const float wy = (float)(-oi * ((oj == 0) ? 2 : 1));      // sobel_y[oi+1][oj+1]
                    const float4* pn = s_P2 + (tid + HALO + oi * R + oj);
                    const float4 v0 = BE_SLD4(pn), v1 = BE_SLD4(pn + NE);
                    const f2 uv[4] = {mk2(v0.x, v0.y), mk2(v0.z, v0.w), mk2(v1.x, v1.y), mk2(v1.z, v1.w)};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (wx != 0.0f) ux[c] = fma2(bc2(wx), uv[c], ux[c]);
                        if (wy != 0.0f) uy[c] = fma2(bc2(wy), uv[c], uy[c]);
                    }
                }
            f2 sx[6], sy[6];
            float D1[3], D2[3];
            {
                const float* cc = s_crec[cur];
#pragma unroll
                for (int c = 0; c < 3; ++c) { D1[c] = cc[3 + c] - cc[c]; D2[c] = cc[6 + c] - cc[c]; }
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        sx[3 * m + c] = fma2(ux[2 * m], bc2(D1[c]), mul2(ux[2 * m + 1], bc2(D2[c])));
                        sy[3 * m + c] = fma2(uy[2 * m], bc2(D1[c]), mul2(uy[2 * m + 1], bc2(D2[c])));
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
                const f2 gm = mul2(mul2(fma2(bc2(2.0f * a.ksc), e2, mul2(bc2(2.0f * a.ks), e1)), mI), ir);   // zero off the interior
                gxy[c] = mul2(gm, sx[c]);
                gxy[6 + c] = mul2(gm, sy[c]);
            }
            lossacc[3] = fmaf(lo(l3), lo(mI), fmaf(hi(l3), hi(mI), lossacc[3]));
            lossacc[4] = fmaf(lo(l4), lo(mI), fmaf(hi(l4), hi(mI), lossacc[4]));
            // What the rest of the backward pass needs from the adjoint of the Sobel filter, A_c = K^T (gx_c, gy_c):
            //  * dL/du_1 - dL/du_0 and dL/du_2 - dL/du_0 use sum_c A_c (C_k - C_0)_c = K^T applied to the PROJECTED fields
            //    sum_c (C_k - C_0)_c gx_c, sum_c (C_k - C_0)_c gy_c: 8 values per pixel go through shared memory instead of 12;
            //  * A^T G uses sum_p u_w(p) A_c(p) = sum_q gx_c(q) Sobel_x(u_w)(q) + gy_c(q) Sobel_y(u_w)(q), which is local to q.
            f2 proj[8];                              // (PX1, PX2) image 1, (PX1, PX2) image 2, then the same for PY
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                proj[2 * m] = fma2(gxy[3 * m], bc2(D1[0]), fma2(gxy[3 * m + 1], bc2(D1[1]), mul2(gxy[3 * m + 2], bc2(D1[2]))));
                proj[2 * m + 1] = fma2(gxy[3 * m], bc2(D2[0]), fma2(gxy[3 * m + 1], bc2(D2[1]), mul2(gxy[3 * m + 2], bc2(D2[2]))));
                proj[4 + 2 * m] = fma2(gxy[6 + 3 * m], bc2(D1[0]), fma2(gxy[7 + 3 * m], bc2(D1[1]), mul2(gxy[8 + 3 * m], bc2(D1[2]))));
                proj[5 + 2 * m] = fma2(gxy[6 + 3 * m], bc2(D2[0]), fma2(gxy[7 + 3 * m], bc2(D2[1]), mul2(gxy[8 + 3 * m], bc2(D2[2]))));
            }
            store_pairs(s_G2, proj, 8);
            f2 sums[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) sums[i] = bc2(0.0f);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const f2 gg = sub2(bc2(1.0f), h[2 * m + 1]);
                const f2 u[3] = {mul2(sub2(bc2(1.0f), h[2 * m]), gg), mul2(h[2 * m], gg), h[2 * m + 1]};
                const f2 UX[3] = {neg2(add2(ux[2 * m], ux[2 * m + 1])), ux[2 * m], ux[2 * m + 1]};      // Sobel(u0) = -Sobel(u1) - Sobel(u2)
                const f2 UY[3] = {neg2(add2(uy[2 * m], uy[2 * m + 1])), uy[2 * m], uy[2 * m + 1]};
#pragma unroll
                for (int wd = 0; wd < 3; ++wd)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        sums[3 * wd + c] = fma2(u[wd], G[3 * m + c], fma2(UX[wd], gxy[3 * m + c], fma2(UY[wd], gxy[6 + 3 * m + c], sums[3 * wd + c])));
            }
            float4* row = s_partA + (warp * 32 + lane) * 3;
            row[0] = make_float4(fold(sums[0]), fold(sums[1]), fold(sums[2]), fold(sums[3]));
            row[1] = make_float4(fold(sums[4]), fold(sums[5]), fold(sums[6]), fold(sums[7]));
            row[2] = make_float4(fold(sums[8]), 0.0f, 0.0f, 0.0f);
        }
        bar_arrive_id<BAR_ATG, NTHR>();  // A^T G partial sums handed to the helper
#if BE_LOSS2_DIAG != 4
        bar_sync_id<BAR_RENDER, NT>();   // (X2) projected Sobel gradients visible
#endif

        // ---------------- stage C: adjoint of the Sobel filter on the projected fields ----------------
        f2 AE[4];                                    // sum_c A_c (C_k - C_0)_c for (k = 1, 2) of image 1, then of image 2
#pragma unroll
        for (int i = 0; i < 4; ++i) AE[i] = bc2(0.0f);
#pragma unroll
        for (int di = -1; di <= 1; ++di)
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                if (di == 0 && dj == 0) continue;
                const float wx = (float)(-dj * ((di == 0) ? 2 : 1));   // weight of gx(i+di, j+dj) in dL/dP(i,j)
                const float wy = (float)(di * ((dj == 0) ? 2 : 1));    // weight of gy(i+di, j+dj)
                const float4* pn = s_G2 + (tid + HALO + di * R + dj);
                if (wx != 0.0f) {
                    const float4 v0 = BE_SLD4(pn), v1 = BE_SLD4(pn + NE);
                    AE[0] = fma2(bc2(wx), mk2(v0.x, v0.y), AE[0]); AE[1] = fma2(bc2(wx), mk2(v0.z, v0.w), AE[1]);
                    AE[2] = fma2(bc2(wx), mk2(v1.x, v1.y), AE[2]); AE[3] = fma2(bc2(wx), mk2(v1.z, v1.w), AE[3]);
                }
                if (wy != 0.0f) {
                    const float4 v0 = BE_SLD4(pn + 2 * NE), v1 = BE_SLD4(pn + 3 * NE);
                    AE[0] = fma2(bc2(wy), mk2(v0.x, v0.y), AE[0]); AE[1] = fma2(bc2(wy), mk2(v0.z, v0.w), AE[1]);
                    AE[2] = fma2(bc2(wy), mk2(v1.x, v1.y), AE[2]); AE[3] = fma2(bc2(wy), mk2(v1.z, v1.w), AE[3]);
                }
            }

        // ---------------- stage D: per-pixel backward -> 14 per-patch sums -> helper ----------------
        {
            BePatch P;
            float C[9];
            load_patch(P);
            load_colors(C);
            float ny[2][6];
            float zgv[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
#if BE_LOSS2_DIAG == 2
                const float4 q0 = BE_TLD4(a.T); const float2 q1 = BE_TLD2(a.T);
#else
                const float4 q0 = SAMEGT ? ldg_again4(a.T + tp[s]) : __ldg(reinterpret_cast<const float4*>(a.T + tp[s]));
                const float2 q1 = SAMEGT ? ldg_again2(a.T + (tp[s] + TPS)) : __ldg(reinterpret_cast<const float2*>(a.T + (tp[s] + TPS)));
#endif
                ny[s][0] = q0.x; ny[s][1] = q0.y; ny[s][2] = q0.z; ny[s][3] = q0.w; ny[s][4] = q1.x; ny[s][5] = q1.y;
                zgv[s] = BE_TLD1(a.T + (8u * TPS + (tp[s] >> 2)));        // value 32: the scalar plane
            }
            const float4 sd = s_stash0[tid], sg = s_stash1[tid];
            const f2 d1 = mk2(sd.x, sd.y), d2 = mk2(sd.z, sd.w), gbv = mk2(sg.x, sg.y), bdv = mk2(sg.z, sg.w);
            // -- part 1: boundary and depth terms, which do not depend on V, S (the helper has had stage C to reduce and solve) --
            f2 sums[14];
            const f2 lb = be_boundary2(d1, d2);
            const f2 bl = mul2(bdv, lb);
            lossacc[5] += fold(mul2(bl, bl));
            const f2 eb = sub2(lb, gbv);
            lossacc[2] += fold(mul2(eb, eb));
            const f2 glb = fma2(bc2(2.0f * a.kbc), eb, mul2(mul2(bc2(2.0f * a.kbl), bdv), bl));
            f2 gd1, gd2;
            {
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
                gd1 = mk2(gb1[0], gb1[1]); gd2 = mk2(gb2[0], gb2[1]);
            }
            // -- part 2: the ridge backward through V, S --
            mbar_wait(&s_vready, (unsigned)(k & 1));
            float DV[2][3], DS[2][3], D1[3], D2[3];     // V_k - V_0, S_k - S_0 (k = 1, 2; from the helper), C_k - C_0
            {
                const float4* vs = reinterpret_cast<const float4*>(s_VS);
                const float4 v0 = vs[0], v1 = vs[1], v2 = vs[2];
                DV[0][0] = v0.x; DV[0][1] = v0.y; DV[0][2] = v0.z; DV[1][0] = v0.w; DV[1][1] = v1.x; DV[1][2] = v1.y;
                DS[0][0] = v1.z; DS[0][1] = v1.w; DS[0][2] = v2.x; DS[1][0] = v2.y; DS[1][1] = v2.z; DS[1][2] = v2.w;
#pragma unroll
                for (int c = 0; c < 3; ++c) { D1[c] = C[3 + c] - C[c]; D2[c] = C[6 + c] - C[c]; }
            }
#pragma unroll
            for (int i = 8; i < 12; ++i) sums[i] = bc2(0.0f);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const f2 h1 = h[2 * m], h2 = h[2 * m + 1];
                const f2 gg = sub2(bc2(1.0f), h2), g1 = sub2(bc2(1.0f), h1);
                const f2 u[3] = {mul2(g1, gg), mul2(h1, gg), h2};
                // E_k = dL/du_k - dL/du_0 = sum_c (G_c + A_c)(C_k - C_0)_c + y_c (V_k - V_0)_c - ((S_k - S_0) u)     (be_ridge_backward_pixel)
                f2 E[2];
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const float* Dk = kk ? D2 : D1;
                    float yv[2];
#pragma unroll
                    for (int s = 0; s < 2; ++s)
                        yv[s] = fmaf(ny[s][3 * m], DV[kk][0], fmaf(ny[s][3 * m + 1], DV[kk][1], ny[s][3 * m + 2] * DV[kk][2]));
                    f2 t = add2(AE[2 * m + kk], mk2(yv[0], yv[1]));
#pragma unroll
                    for (int c = 0; c < 3; ++c) t = fma2(G[3 * m + c], bc2(Dk[c]), t);
#pragma unroll
                    for (int v = 0; v < 3; ++v) t = fma2(u[v], bc2(-DS[kk][v]), t);
                    E[kk] = t;
                }
                const f2 gh1 = mul2(gg, E[0]);                                                 // be_wedges_backward
                const f2 gh2 = sub2(E[1], mul2(h1, E[0]));
                f2 da, de;
                be_h_grad2(d1, P.inv_eta[2 * m], &da, &de);
                gd1 = fma2(gh1, da, gd1); sums[8 + 2 * m] = fma2(gh1, de, sums[8 + 2 * m]);
                be_h_grad2(d2, P.inv_eta[2 * m + 1], &da, &de);
                gd2 = fma2(gh2, da, gd2); sums[9 + 2 * m] = fma2(gh2, de, sums[9 + 2 * m]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) sums[i] = bc2(0.0f);
            be_wedge_backward2(P, 0, X, Y, g.w, gd1, &sums[0]);
            be_wedge_backward2(P, 1, X, Y, g.w, gd2, &sums[4]);
            float ssum[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) ssum[i] = (i < 14) ? fold(sums[i]) : 0.0f;
            {   // row of this lane, float4 columns rotated by (lane >> 1): conflict-free 16-byte stores and helper loads
                float4* row = s_partB + (warp * 32 + lane) * 4;
                const int rot = lane >> 1;
#pragma unroll
                for (int c = 0; c < 4; ++c) row[(c + rot) & 3] = make_float4(ssum[4 * c], ssum[4 * c + 1], ssum[4 * c + 2], ssum[4 * c + 3]);
            }
        }
        bar_arrive_id<BAR_SUMS, NTHR>();
    }

    // ---------------- per-CTA partial loss sums ----------------
    {
        float sums[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) sums[i] = (i < 7) ? lossacc[i] : 0.0f;
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
}

}  // namespace

void be_launch_loss2(const BeLossArgs& a, cudaStream_t st) {
    const int grid = a.NB * a.g.Hp * a.runs_per_row;
    static bool configured[4][BE_MAX_DEVICES] = {};
    be_opt_in_smem(be_loss2_kernel<21, false>, DYN_SMEM, configured[0]);
    be_opt_in_smem(be_loss2_kernel<21, true>, DYN_SMEM, configured[1]);
    be_opt_in_smem(be_loss2_kernel<0, false>, DYN_SMEM, configured[2]);
    be_opt_in_smem(be_loss2_kernel<0, true>, DYN_SMEM, configured[3]);
    if (a.g.R == 21) {
        if (a.same_gt) be_loss2_kernel<21, true><<<grid, NTHR, DYN_SMEM, st>>>(a);
        else be_loss2_kernel<21, false><<<grid, NTHR, DYN_SMEM, st>>>(a);
    } else {
        if (a.same_gt) be_loss2_kernel<0, true><<<grid, NTHR, DYN_SMEM, st>>>(a);
        else be_loss2_kernel<0, false><<<grid, NTHR, DYN_SMEM, st>>>(a);
    }
    ++g_be_launches;
}

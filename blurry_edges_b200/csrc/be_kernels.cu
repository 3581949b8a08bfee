// Small sm_100a kernels around the renderer of the Blurry-Edges render -> fold -> depth path.
//
// Work decomposition:
//   be_setup_kernel      one thread per patch: parameters -> 128-byte record (sin/cos of the 4 edges, flips,
//                        etas, analytic depths).  Keeps every transcendental that is constant per patch off the
//                        critical path of the renderer.
//   (renderer + fused fold: be_run3.cu; loss kernels: be_train.cu, be_loss2.cu)
//   be_normalise_kernel  accumulator -> the six planar maps (divide by the closed-form cover count / depth count).
#include "be_internal.h"

long long g_be_launches = 0;

namespace {

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
    // 1 / (sqrt2 * refocus sigma) of the two wedge depths (utils/depth_etas.py:36-37, blurry_edges_test.py:66-74): per-patch
    // constants, kept off the solver warp's critical path
    rec[6] = make_float4(1.0f / (BE_SQRT2_F * be_refocus_sigma(cam, P.z[0])), 1.0f / (BE_SQRT2_F * be_refocus_sigma(cam, P.z[1])), 0.f, 0.f);
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

__device__ __forceinline__ int cover_1d(int y, int R, int s, int np) {
    const int hi = min(y / s, np - 1);
    const int lo = (y - R + 1 <= 0) ? 0 : (y - R + s) / s;
    return max(hi - lo + 1, 0);
}

// `rows`, `y_first`: the accumulator (and the outputs) hold image rows [y_first, y_first + rows) of the g.H rows of the image - the
// whole image (rows = g.H, y_first = 0) or the row band one rank of a block-sharded big-image job owns.
__global__ void __launch_bounds__(256) be_normalise_kernel(const float* __restrict__ acc, BeGeom g, int B, int rows, int y_first, float thres,
                                                           float* __restrict__ image, float* __restrict__ sharp,
                                                           float* __restrict__ refoc, float* __restrict__ bndry,
                                                           float* __restrict__ depth, float* __restrict__ conf,
                                                           float* __restrict__ depth_thr) {
    const size_t HW = (size_t)rows * g.W;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * HW) return;
    const int b = (int)(idx / HW);
    const size_t p = idx % HW;
    const int y = y_first + (int)(p / g.W), x = (int)(p % g.W);
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

// Deterministic fold, second half: acc[b][y][x][:] = sum over the patch rows py that cover image row y, in ASCENDING py, of the slab
// cell stage[b][py][y - py*stride][x][:] that the CTA of (b, py) wrote with a plain store (be_run3.cu).  One thread per float4 of a
// pixel; consecutive threads read consecutive 16 bytes.  The sum order is fixed, so repeated launches are bit-identical
// (torch.use_deterministic_algorithms, global_training.py:177 / utils/util_func.py:17-19).  Overwrites acc: no memset needed.
__global__ void __launch_bounds__(256) be_stage_reduce_kernel(const float4* __restrict__ stage, BeGeom g, int B, int q4, float4* __restrict__ acc) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t HW = (size_t)g.H * g.W;
    if (idx >= (size_t)B * HW * q4) return;
    const int q = (int)(idx % q4);
    const size_t pix = idx / q4;
    const int b = (int)(pix / HW);
    const int y = (int)((pix % HW) / g.W), x = (int)(pix % g.W);
    const int hi = min(y / g.stride, g.Hp - 1);
    const int lo = (y - g.R + 1 <= 0) ? 0 : (y - g.R + g.stride) / g.stride;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x <= (g.Wp - 1) * g.stride + g.R - 1) {
        for (int py = lo; py <= hi; ++py) {
            const float4 v = stage[((((size_t)b * g.Hp + py) * g.R + (y - py * g.stride)) * g.W + x) * q4 + q];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    acc[idx] = s;
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

void be_launch_normalise(const float* acc, const BeGeom& g, int B, int rows, int y_first, float thres, float* image, float* sharp, float* refoc,
                         float* bndry, float* depth, float* conf, float* depth_thr, cudaStream_t st) {
    const size_t n = (size_t)B * rows * g.W;
    be_normalise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(acc, g, B, rows, y_first, thres, image, sharp, refoc, bndry, depth, conf,
                                                                     depth_thr);
    ++g_be_launches;
}

void be_launch_stage_reduce(const float* stage, const BeGeom& g, int B, int accw, float* acc, cudaStream_t st) {
    const size_t n = (size_t)B * g.H * g.W * (accw / 4);
    be_stage_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(stage), g, B, accw / 4,
                                                                        reinterpret_cast<float4*>(acc));
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

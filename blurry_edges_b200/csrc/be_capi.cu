// C ABI of libblurry_edges_b200.so (declared in include/blurry_edges_b200.h).  Plain pointers and sizes only.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "../../include/blurry_edges_b200.h"
#include "be_internal.h"

namespace {

thread_local char g_err[512] = "";

int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

#define BE_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) return fail("%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

#define BE_REQUIRE(cond, ...) \
    do {                      \
        if (!(cond)) return fail(__VA_ARGS__); \
    } while (0)

}  // namespace

struct be_ctx {
    be_config cfg;
    BeGeom g;
    BeCam cam;
    double consts[8];
    int device;
    // HBM workspace
    float* table;       // [2*max_batch*L][BE_REC]
    float* acc;         // [max_batch][H][W][BE_ACC]
    size_t table_bytes, acc_bytes;
    // staging for the be_host_* entry points (lazily allocated)
    float* st_est;
    float* st_img;
    float* st_out;
    size_t st_bytes;
    cudaStream_t st_stream;
    // optional per-kernel timing of the last be_render_fold_fwd call (be_ctx_set_timing)
    int timing;
    cudaEvent_t ev[5];
};

namespace {

int check_ctx(const be_ctx* c) {
    BE_REQUIRE(c != nullptr, "null context");
    int dev = -1;
    BE_CUDA(cudaGetDevice(&dev));
    BE_REQUIRE(dev == c->device, "context was created on device %d but the current device is %d", c->device, dev);
    return 0;
}

int check_layout(const be_image_layout* l) {
    BE_REQUIRE(l != nullptr, "null image layout");
    return 0;
}

BeImg make_img(const float* p, const be_image_layout* l) {
    BeImg im;
    im.p = p; im.sb = l->sb; im.sm = l->sm; im.sc = l->sc; im.sy = l->sy; im.sx = l->sx;
    return im;
}

// PostProcessGlobalBase.__init__ / DepthEtas.__init__ constants (utils/postprocessing_loss.py:14,137-138,
// utils/depth_etas.py:4-21).  Pure host arithmetic: python-float (double) where the reference uses python floats,
// fp32 where it uses fp32 tensors.
void be_derive(const be_config* cfg, BeGeom* g, BeCam* cam, double* consts) {
    g->R = cfg->R; g->stride = cfg->stride; g->H = cfg->H; g->W = cfg->W;
    g->Hp = (cfg->H - cfg->R) / cfg->stride + 1;
    g->Wp = (cfg->W - cfg->R) / cfg->stride + 1;
    g->w = (float)cfg->w;
    const double al = cfg->alpha_lambda * (double)(cfg->R * cfg->R);
    const double lam = al * al;
    g->lam = (float)lam;   // `ridge` is an fp32 tensor (:122,:133)
    const double s = cfg->cam_s, r1 = cfg->cam_rho_1, r2 = cfg->cam_rho_2;
    const int nf = cfg->R / 2;
    const double numerator = 2.0 * s * s * (r2 - r1);
    const double k_const = -s * (r1 - r2) * (r1 * s + r2 * s - 2.0);
    const double k_root = nf * cfg->cam_pixel_pitch * cfg->cam_mag / cfg->cam_sigma_cam;
    // intercept: torch.abs(torch.tensor(s*(rho_2-rho_1))) * sigma_cam / pixel_pitch / mag / norm_factor, an fp32 chain
    volatile float icpt = fabsf((float)(s * (r2 - r1)));
    icpt = icpt * (float)cfg->cam_sigma_cam;
    icpt = icpt / (float)cfg->cam_pixel_pitch;
    icpt = icpt / (float)cfg->cam_mag;
    icpt = icpt / (float)nf;
    cam->numerator = (float)numerator; cam->k_fac = (float)(k_root * k_root); cam->k_const = (float)k_const;
    cam->k_root = (float)k_root; cam->intercept = icpt;
    const float tw = (float)(M_PI / 4.0), tm = (float)(3.0 * M_PI / 4.0);
    cam->sin_w = sinf(tw); cam->cos_w = cosf(tw); cam->sin_m = sinf(tm); cam->cos_m = cosf(tm);
    cam->s = (float)s; cam->rho_prime = (float)cfg->rho_prime;
    consts[0] = numerator; consts[1] = k_const; consts[2] = k_root; consts[3] = k_root * k_root;
    consts[4] = icpt; consts[5] = lam; consts[6] = g->Hp; consts[7] = g->Wp;
}

int validate_cfg(const be_config* cfg) {
    BE_REQUIRE(cfg != nullptr, "null config");
    BE_REQUIRE(cfg->R >= 3 && cfg->R <= BE_MAX_R, "R=%d unsupported (3..%d)", cfg->R, BE_MAX_R);
    BE_REQUIRE(cfg->stride >= 1 && cfg->stride < cfg->R, "stride=%d must be in [1,R)", cfg->stride);
    BE_REQUIRE(cfg->H >= cfg->R && cfg->W >= cfg->R, "image %dx%d smaller than the patch", cfg->H, cfg->W);
    return 0;
}

int pick_runs(const BeGeom& g, int* G, int* runs) {
    // one CTA per patch row unless the row is very long (big images): then chunks of <= 96 patches
    const int maxG = 96;
    *runs = (g.Wp + maxG - 1) / maxG;
    *G = (g.Wp + *runs - 1) / *runs;
    return 0;
}

}  // namespace

extern "C" {

int be_abi_version(void) { return BE_ABI_VERSION; }
const char* be_last_error(void) { return g_err; }
int64_t be_launch_count(void) { return g_be_launches; }

int be_ctx_create(be_ctx** out, const be_config* cfg) {
    BE_REQUIRE(out && cfg, "null argument");
    *out = nullptr;
    if (validate_cfg(cfg)) return 1;
    BE_REQUIRE(cfg->max_batch >= 1, "max_batch must be >= 1");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    BE_REQUIRE(e == cudaSuccess && ndev > 0, "no CUDA device: this library has no CPU fallback (%s)", cudaGetErrorString(e));

    be_ctx* c = new (std::nothrow) be_ctx();
    BE_REQUIRE(c, "out of host memory");
    memset(c, 0, sizeof(*c));
    c->cfg = *cfg;
    BE_CUDA(cudaGetDevice(&c->device));
    be_derive(cfg, &c->g, &c->cam, c->consts);
    BeGeom& g = c->g;

    const size_t L = (size_t)g.Hp * g.Wp;
    c->table_bytes = 2 * (size_t)cfg->max_batch * L * BE_REC * sizeof(float);
    c->acc_bytes = (size_t)cfg->max_batch * g.H * g.W * BE_ACC * sizeof(float);
    if (cudaMalloc(&c->table, c->table_bytes) != cudaSuccess || cudaMalloc(&c->acc, c->acc_bytes) != cudaSuccess) {
        cudaFree(c->table);
        delete c;
        return fail("cudaMalloc of the %.1f MiB workspace failed", (c->table_bytes + c->acc_bytes) / 1048576.0);
    }
    *out = c;
    return 0;
}

int be_derive_constants(const be_config* cfg, double* out8) {
    BE_REQUIRE(out8, "null argument");
    if (validate_cfg(cfg)) return 1;
    BeGeom g; BeCam cam;
    be_derive(cfg, &g, &cam, out8);
    return 0;
}

int be_ctx_destroy(be_ctx* c) {
    if (!c) return 0;
    cudaFree(c->table); cudaFree(c->acc);
    cudaFree(c->st_est); cudaFree(c->st_img); cudaFree(c->st_out);
    if (c->st_stream) cudaStreamDestroy(c->st_stream);
    for (int i = 0; i < 5; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    delete c;
    return 0;
}

int64_t be_ctx_workspace_bytes(const be_ctx* c) { return c ? (int64_t)(c->table_bytes + c->acc_bytes + c->st_bytes) : 0; }

int be_ctx_constants(const be_ctx* c, double* out8) {
    BE_REQUIRE(c && out8, "null argument");
    memcpy(out8, c->consts, sizeof(c->consts));
    return 0;
}

int be_ctx_set_timing(be_ctx* c, int32_t enable) {
    if (check_ctx(c)) return 1;
    if (enable && !c->ev[0])
        for (int i = 0; i < 5; ++i) BE_CUDA(cudaEventCreate(&c->ev[i]));
    c->timing = enable;
    return 0;
}

int be_ctx_last_timing(be_ctx* c, float* ms4) {
    if (check_ctx(c)) return 1;
    BE_REQUIRE(ms4 && c->timing && c->ev[0], "timing is not enabled");
    BE_CUDA(cudaEventSynchronize(c->ev[4]));
    for (int i = 0; i < 4; ++i) BE_CUDA(cudaEventElapsedTime(&ms4[i], c->ev[i], c->ev[i + 1]));
    return 0;
}

int be_cover_count(be_ctx* c, float* dev_out, void* stream) {
    if (check_ctx(c)) return 1;
    BE_REQUIRE(dev_out, "null output");
    be_launch_cover_count(c->g, dev_out, (cudaStream_t)stream);
    BE_CUDA(cudaGetLastError());
    return 0;
}

int be_refold_image(be_ctx* c, const float* dev_unfolded, int32_t M, float* dev_image, void* stream) {
    if (check_ctx(c)) return 1;
    BE_REQUIRE(dev_unfolded && dev_image, "null pointer");
    BE_REQUIRE(M >= 0, "negative image count");
    if (M == 0) return 0;
    be_launch_refold(dev_unfolded, c->g, M, dev_image, (cudaStream_t)stream);
    BE_CUDA(cudaGetLastError());
    return 0;
}

int be_colors_fwd(be_ctx* c, const float* dev_est, int32_t param_mode, const float* dev_img, const be_image_layout* layout,
                  int32_t M, float* dev_colors, void* stream) {
    if (check_ctx(c) || check_layout(layout)) return 1;
    if (M == 0) return 0;
    BE_REQUIRE(dev_est && dev_img && dev_colors, "null pointer");
    BE_REQUIRE(param_mode == BE_PARAMS_LOCAL10 || param_mode == BE_PARAMS_LOCALRAW10, "be_colors_fwd takes 10-parameter patches");
    BE_REQUIRE(M >= 0 && M <= 2 * c->cfg.max_batch, "M=%d exceeds 2*max_batch=%d", M, 2 * c->cfg.max_batch);
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int L = c->g.Hp * c->g.Wp;
    be_launch_setup(dev_est, param_mode, M * L, c->cam, c->table, st);
    BeRunArgs a;
    memset(&a, 0, sizeof(a));
    a.table = c->table; a.img = make_img(dev_img, layout); a.colors = dev_colors;
    a.g = c->g; a.cam = c->cam; a.NB = M;
    pick_runs(c->g, &a.G, &a.runs_per_row);
    be_launch_run(BE_RUN_COLORS, a, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}

int be_render_fold_fwd(be_ctx* c, const float* dev_est, int32_t param_mode, const float* dev_img, const be_image_layout* layout,
                       int32_t B, int32_t densify_w, float* dev_image, float* dev_sharp, float* dev_refoc, float* dev_bndry,
                       float* dev_depth, float* dev_conf, float* dev_depth_thr, void* stream) {
    if (check_ctx(c) || check_layout(layout)) return 1;
    if (B == 0) return 0;
    BE_REQUIRE(dev_est && dev_img && dev_image && dev_sharp && dev_refoc && dev_bndry && dev_depth && dev_conf, "null pointer");
    BE_REQUIRE(param_mode == BE_PARAMS_RESTORED12 || param_mode == BE_PARAMS_RAW12, "be_render_fold_fwd takes 12-parameter patches");
    BE_REQUIRE(B >= 0 && B <= c->cfg.max_batch, "B=%d exceeds max_batch=%d", B, c->cfg.max_batch);
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const BeGeom& g = c->g;
    const int L = g.Hp * g.Wp;
    const bool tm = c->timing != 0;
    if (tm) cudaEventRecord(c->ev[0], st);
    BE_CUDA(cudaMemsetAsync(c->acc, 0, (size_t)B * g.H * g.W * BE_ACC * sizeof(float), st));
    if (tm) cudaEventRecord(c->ev[1], st);
    be_launch_setup(dev_est, param_mode, B * L, c->cam, c->table, st);
    if (tm) cudaEventRecord(c->ev[2], st);
    BeRunArgs a;
    memset(&a, 0, sizeof(a));
    a.table = c->table; a.img = make_img(dev_img, layout); a.acc = c->acc;
    a.g = g; a.cam = c->cam; a.NB = B; a.densify_w = densify_w;
    pick_runs(g, &a.G, &a.runs_per_row);
    be_launch_run(BE_RUN_INFER, a, st);
    if (tm) cudaEventRecord(c->ev[3], st);
    const float thres = densify_w ? 0.0f : 0.05f;   // blurry_edges_test.py:109-112
    be_launch_normalise(c->acc, g, B, thres, dev_image, dev_sharp, dev_refoc, dev_bndry, dev_depth, dev_conf, dev_depth_thr, st);
    if (tm) cudaEventRecord(c->ev[4], st);
    BE_CUDA(cudaGetLastError());
    return 0;
}

int be_host_render_fold(be_ctx* c, const float* est, int32_t param_mode, const float* img, const be_image_layout* layout,
                        int32_t B, int32_t densify_w, float* image, float* sharp, float* refoc, float* bndry, float* depth,
                        float* conf, float* depth_thr) {
    if (check_ctx(c) || check_layout(layout)) return 1;
    if (B == 0) return 0;
    BE_REQUIRE(est && img && image && sharp && refoc && bndry && depth && conf, "null pointer");
    BE_REQUIRE(B >= 0 && B <= c->cfg.max_batch, "B=%d exceeds max_batch=%d", B, c->cfg.max_batch);
    if (B == 0) return 0;
    const BeGeom& g = c->g;
    const size_t HW = (size_t)g.H * g.W, L = (size_t)g.Hp * g.Wp;
    const size_t mb = (size_t)c->cfg.max_batch;
    if (!c->st_est) {
        BE_CUDA(cudaStreamCreateWithFlags(&c->st_stream, cudaStreamNonBlocking));
        BE_CUDA(cudaMalloc(&c->st_est, mb * L * 12 * sizeof(float)));
        BE_CUDA(cudaMalloc(&c->st_img, mb * 6 * HW * sizeof(float)));
        BE_CUDA(cudaMalloc(&c->st_out, mb * 16 * HW * sizeof(float)));
        c->st_bytes = mb * (L * 12 + 22 * HW) * sizeof(float);
    }
    cudaStream_t st = c->st_stream;
    BE_CUDA(cudaMemcpyAsync(c->st_est, est, (size_t)B * L * 12 * sizeof(float), cudaMemcpyHostToDevice, st));
    BE_CUDA(cudaMemcpyAsync(c->st_img, img, (size_t)B * 6 * HW * sizeof(float), cudaMemcpyHostToDevice, st));
    float* o = c->st_out;
    float* d_image = o;                  float* d_sharp = o + (size_t)B * 6 * HW;
    float* d_refoc = o + (size_t)B * 9 * HW;   float* d_bndry = o + (size_t)B * 12 * HW;
    float* d_depth = o + (size_t)B * 13 * HW;  float* d_conf = o + (size_t)B * 14 * HW;
    float* d_thr = o + (size_t)B * 15 * HW;
    if (be_render_fold_fwd(c, c->st_est, param_mode, c->st_img, layout, B, densify_w, d_image, d_sharp, d_refoc, d_bndry,
                           d_depth, d_conf, d_thr, (void*)st))
        return 1;
    const size_t f = sizeof(float);
    BE_CUDA(cudaMemcpyAsync(image, d_image, (size_t)B * 6 * HW * f, cudaMemcpyDeviceToHost, st));
    BE_CUDA(cudaMemcpyAsync(sharp, d_sharp, (size_t)B * 3 * HW * f, cudaMemcpyDeviceToHost, st));
    BE_CUDA(cudaMemcpyAsync(refoc, d_refoc, (size_t)B * 3 * HW * f, cudaMemcpyDeviceToHost, st));
    BE_CUDA(cudaMemcpyAsync(bndry, d_bndry, (size_t)B * HW * f, cudaMemcpyDeviceToHost, st));
    BE_CUDA(cudaMemcpyAsync(depth, d_depth, (size_t)B * HW * f, cudaMemcpyDeviceToHost, st));
    BE_CUDA(cudaMemcpyAsync(conf, d_conf, (size_t)B * HW * f, cudaMemcpyDeviceToHost, st));
    if (depth_thr) BE_CUDA(cudaMemcpyAsync(depth_thr, d_thr, (size_t)B * HW * f, cudaMemcpyDeviceToHost, st));
    BE_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"

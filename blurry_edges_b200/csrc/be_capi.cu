// C ABI of libblurry_edges_b200.so (declared in include/blurry_edges_b200.h).  Plain pointers and sizes only.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
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

constexpr int BE_HOST_CHUNKS = 16;  // maximum pipeline depth of the host-buffer entry point
constexpr int BE_TRAIN_EVENTS = 10;
constexpr int BE_TRAIN_SETS = 64;    // ring of event sets: the per-kernel times of the last 64 training steps can be read without a host
                                     // synchronisation inside the timed loop

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
    cudaStream_t st_streams[3];
    cudaEvent_t st_events[2 * BE_HOST_CHUNKS];
    // block descriptors of the blocked (big-image) entry points
    BeBlock* blk_dev;
    BeBlock* blk_pin;
    int blk_cap;
    int blk_n;          // descriptors currently on the device (upload_blocks skips an identical list)
    cudaEvent_t blk_ev;
    // training workspace (lazily allocated by the loss entry points)
    float* gtable;      // [max_batch*L][BE_GREC]
    float* crec;        // [max_batch*L][BE_CREC]
    float* T;           // [max_batch][H][W][BE_TW]
    float* partials;    // [max_batch*Hp*runs][8]
    float* lpart;       // local loss: [max_batch][8] partial sums + the ticket of its last-CTA reduction
    int same_gt;        // the last be_global_loss_stage1 call had img_gt == img_ny
    int train_B;        // pairs of the batch whose stage 1 ran last
    int train_parts;    // rows of `partials` the last loss-kernel launch(es) wrote
    size_t train_bytes;
    // staging of be_host_global_loss (lazily allocated)
    float *ht_raw, *ht_ny, *ht_gt, *ht_bd, *ht_deri, *ht_zg, *ht_grad, *ht_gdep, *ht_scal;
    int ht_want_grad;
    cudaEvent_t tev_ring[BE_TRAIN_SETS][BE_TRAIN_EVENTS];
    cudaEvent_t* tev;   // per-kernel timing: the event set of the current training step (be_ctx_last_train_timing / be_ctx_train_timing_at)
    int tset;           // index of `tev` in the ring
    long long tsteps;   // training steps timed since timing was enabled
    // deterministic fold (be_ctx_set_deterministic): per-patch-row slabs [max_batch][Hp][R][W][accw], allocated on first use
    int deterministic;
    float* stage;
    size_t stage_bytes;
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

void launch_run(int mode, const BeRunArgs& a, cudaStream_t st) { be_launch_run3(mode, a, st); }

// Split every patch row into `runs` runs of G consecutive patches (one CTA each).  Long runs amortise the per-CTA start-up
// (pixel loads, pipeline fill, the final flush of all 441 accumulators: ~`ovh` patch-times); but the grid runs in waves of 148 * ctas_per_sm CTAs and
// a partly filled last wave costs a whole one, which matters for small batches and for the chunks of the host pipeline.  Pick the
// split that minimises  waves * (G + ovh)  over runs of 8..96 patches.
int pick_runs(const BeGeom& g, int items, int ctas_per_sm, double ovh, int* G, int* runs) {
    const int maxG = 96, minG = 8;
    const long long rows = (long long)items * g.Hp, slots = 148LL * ctas_per_sm;
    double best = 1e300;
    int bestG = g.Wp < maxG ? g.Wp : maxG;
    for (int r = (g.Wp + maxG - 1) / maxG; r <= g.Wp; ++r) {
        const int Gr = (g.Wp + r - 1) / r;
        if (Gr < minG && r > 1) break;
        const int rr = (g.Wp + Gr - 1) / Gr;
        const long long waves = (rows * rr + slots - 1) / slots;
        const double cost = (double)waves * (Gr + ovh);
        if (cost < best * 0.999) { best = cost; bestG = Gr; }
    }
    *G = bestG;
    *runs = (g.Wp + bestG - 1) / bestG;
    return 0;
}
constexpr int RUN_CTAS = 2, COLORS_CTAS = 3, LOSS_CTAS = 2;        // resident CTAs per SM of be_run3_kernel / be_loss2_kernel
constexpr double RUN_OVH = 5.0, LOSS_OVH = 1.0;   // calibrated: G=64 vs G=32 at 64 pairs differ by 1.9 % (be_run3)

}  // namespace

static int ensure_stage(be_ctx* c, int accw) {
    const size_t need = (size_t)c->cfg.max_batch * c->g.Hp * c->g.R * c->g.W * accw * sizeof(float);
    if (c->stage && c->stage_bytes >= need) return 0;
    if (c->stage) { BE_CUDA(cudaDeviceSynchronize()); cudaFree(c->stage); c->stage = nullptr; c->stage_bytes = 0; }
    if (cudaMalloc(&c->stage, need) != cudaSuccess)
        return fail("cudaMalloc of the %.1f MiB staging slabs of the deterministic fold failed", need / 1048576.0);
    c->stage_bytes = need;
    return 0;
}

static int ensure_acc(be_ctx* c) {
    if (c->acc) return 0;
    const size_t bytes = (size_t)c->cfg.max_batch * c->g.H * c->g.W * BE_ACC * sizeof(float);
    BE_CUDA(cudaMalloc(&c->acc, bytes));
    c->acc_bytes = bytes;
    return 0;
}

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
    c->acc_bytes = 0;                                    // the fold accumulator is allocated by the first entry point that folds
    if (cudaMalloc(&c->table, c->table_bytes) != cudaSuccess) {
        delete c;
        return fail("cudaMalloc of the %.1f MiB patch-record table failed", c->table_bytes / 1048576.0);
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
    cudaFree(c->table); cudaFree(c->acc); cudaFree(c->stage);
    cudaFree(c->st_est); cudaFree(c->st_img); cudaFree(c->st_out);
    cudaFree(c->gtable); cudaFree(c->T); cudaFree(c->partials); cudaFree(c->crec); cudaFree(c->lpart);
    cudaFree(c->ht_raw); cudaFree(c->ht_ny); cudaFree(c->ht_gt); cudaFree(c->ht_bd); cudaFree(c->ht_deri); cudaFree(c->ht_zg);
    cudaFree(c->ht_grad); cudaFree(c->ht_gdep); cudaFree(c->ht_scal);
    for (int k = 0; k < BE_TRAIN_SETS; ++k)
        for (int i = 0; i < BE_TRAIN_EVENTS; ++i) if (c->tev_ring[k][i]) cudaEventDestroy(c->tev_ring[k][i]);
    cudaFree(c->blk_dev); cudaFreeHost(c->blk_pin);
    if (c->blk_ev) cudaEventDestroy(c->blk_ev);
    for (int i = 0; i < 3; ++i) if (c->st_streams[i]) cudaStreamDestroy(c->st_streams[i]);
    for (int i = 0; i < 2 * BE_HOST_CHUNKS; ++i) if (c->st_events[i]) cudaEventDestroy(c->st_events[i]);
    for (int i = 0; i < 5; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    delete c;
    return 0;
}

int64_t be_ctx_workspace_bytes(const be_ctx* c) { return c ? (int64_t)(c->table_bytes + c->acc_bytes + c->st_bytes + c->train_bytes + c->stage_bytes) : 0; }

int be_ctx_set_deterministic(be_ctx* c, int32_t enable) {
    if (check_ctx(c)) return 1;
    c->deterministic = enable ? 1 : 0;
    return 0;
}

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
    c->tsteps = 0;
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
    be_launch_setup(dev_est, param_mode, M * L, c->cam, c->table, nullptr, st);
    BeRunArgs a;
    memset(&a, 0, sizeof(a));
    a.table = c->table; a.img = make_img(dev_img, layout); a.colors = dev_colors;
    a.g = c->g; a.cam = c->cam; a.NB = M; a.accH = c->g.H; a.accW = c->g.W;
    pick_runs(c->g, M, COLORS_CTAS, RUN_OVH, &a.G, &a.runs_per_row);
    launch_run(BE_RUN_COLORS, a, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}

// pass B on pairs [b0, b0+B) of the context's workspace (records and accumulator are per pair, so disjoint ranges may be in
// flight on different streams)
static int render_fold_range(be_ctx* c, const float* dev_est, int32_t param_mode, const float* dev_img, const be_image_layout* layout,
                             int b0, int32_t B, int32_t densify_w, float* dev_image, float* dev_sharp, float* dev_refoc, float* dev_bndry,
                             float* dev_depth, float* dev_conf, float* dev_depth_thr, cudaStream_t st, bool tm) {
    const BeGeom& g = c->g;
    const int L = g.Hp * g.Wp;
    float* table = c->table + (size_t)b0 * L * BE_REC;
    float* acc = c->acc + (size_t)b0 * g.H * g.W * BE_ACC;
    const bool det = c->deterministic != 0;
    if (det && ensure_stage(c, BE_ACC)) return 1;
    if (tm) cudaEventRecord(c->ev[0], st);
    if (!det) BE_CUDA(cudaMemsetAsync(acc, 0, (size_t)B * g.H * g.W * BE_ACC * sizeof(float), st));
    if (tm) cudaEventRecord(c->ev[1], st);
    be_launch_setup(dev_est, param_mode, B * L, c->cam, table, nullptr, st);
    if (tm) cudaEventRecord(c->ev[2], st);
    BeRunArgs a;
    memset(&a, 0, sizeof(a));
    a.table = table; a.img = make_img(dev_img, layout); a.acc = acc;
    a.g = g; a.cam = c->cam; a.NB = B; a.densify_w = densify_w; a.accH = g.H; a.accW = g.W;
    pick_runs(g, B, RUN_CTAS, RUN_OVH, &a.G, &a.runs_per_row);
    if (det) {   // one CTA per patch row writes its slab with plain stores; a fixed-order pass folds the slabs into the accumulator
        a.G = g.Wp; a.runs_per_row = 1;
        a.stage = c->stage + (size_t)b0 * g.Hp * g.R * g.W * BE_ACC;
    }
    launch_run(BE_RUN_INFER, a, st);
    if (tm) cudaEventRecord(c->ev[3], st);
    if (det) be_launch_stage_reduce(a.stage, g, B, BE_ACC, acc, st);
    const float thres = densify_w ? 0.0f : 0.05f;   // blurry_edges_test.py:109-112
    be_launch_normalise(acc, g, B, g.H, 0, thres, dev_image, dev_sharp, dev_refoc, dev_bndry, dev_depth, dev_conf, dev_depth_thr, st);
    if (tm) cudaEventRecord(c->ev[4], st);
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
    if (ensure_acc(c)) return 1;
    return render_fold_range(c, dev_est, param_mode, dev_img, layout, 0, B, densify_w, dev_image, dev_sharp, dev_refoc, dev_bndry,
                             dev_depth, dev_conf, dev_depth_thr, (cudaStream_t)stream, c->timing != 0);
}

// ---------------------------------------------------------------------------------------------------
// training entry points
// ---------------------------------------------------------------------------------------------------
static int ensure_train_ws(be_ctx* c) {
    if (ensure_acc(c)) return 1;
    if (c->gtable) return 0;
    const BeGeom& g = c->g;
    const size_t mb = (size_t)c->cfg.max_batch, L = (size_t)g.Hp * g.Wp;
    const int runs = (g.Wp + 7) / 8;   // worst case of pick_runs
    const size_t b1 = mb * L * BE_GREC * sizeof(float), b2 = mb * g.H * g.W * BE_TW * sizeof(float);
    const size_t b3 = mb * g.Hp * runs * 8 * sizeof(float);
    const size_t b4 = mb * L * BE_CREC * sizeof(float);
    BE_CUDA(cudaMalloc(&c->gtable, b1));
    BE_CUDA(cudaMalloc(&c->T, b2));
    BE_CUDA(cudaMalloc(&c->partials, b3));
    BE_CUDA(cudaMalloc(&c->crec, b4));
    c->train_bytes = b1 + b2 + b3 + b4;
    return 0;
}

static int ensure_train_events(be_ctx* c) {
    if (c->tev) return 0;
    for (int k = 0; k < BE_TRAIN_SETS; ++k)
        for (int i = 0; i < BE_TRAIN_EVENTS; ++i) BE_CUDA(cudaEventCreate(&c->tev_ring[k][i]));
    c->tset = BE_TRAIN_SETS - 1;
    c->tev = c->tev_ring[c->tset];
    return 0;
}

// Stage 1 on pairs [b0, b0 + nb) of whole-batch device arrays holding Btot pairs: records, TRAINFWD render + fold, global maps,
// packed targets.  The mask count is ADDED to *dev_mask_count (the caller zeroes it once per batch).  Disjoint ranges may be in
// flight at the same time (every workspace array is indexed by pair).
static int loss_stage1_range(be_ctx* c, const float* dev_raw, const float* dev_img_ny, const float* dev_img_gt, const float* dev_bndry_dist,
                             const float* dev_deri, const float* dev_bndry_depth, int b0, int nb, int Btot, float* dev_global_image,
                             float* dev_global_bndry, int64_t* dev_mask_count, cudaStream_t st, bool tm, int parts = 3) {
    const BeGeom& g = c->g;
    const int L = g.Hp * g.Wp;
    const size_t HW = (size_t)g.H * g.W;
    const bool det = c->deterministic != 0;
    if (parts & 1) {   // render: records, TRAINFWD render + fold; the mask count is final when this part is done
        if (det && ensure_stage(c, 8)) return 1;
        if (tm) cudaEventRecord(c->tev[0], st);
        if (!det) BE_CUDA(cudaMemsetAsync(c->acc + (size_t)b0 * HW * 8, 0, (size_t)nb * HW * 8 * sizeof(float), st));
        if (tm) cudaEventRecord(c->tev[1], st);
        be_launch_setup(dev_raw + (size_t)b0 * L * 12, BE_PARAMS_RAW12, nb * L, c->cam, c->table + (size_t)b0 * L * BE_REC,
                        c->gtable + (size_t)b0 * L * BE_GREC, st);
        if (tm) cudaEventRecord(c->tev[2], st);
        BeRunArgs a;
        memset(&a, 0, sizeof(a));
        a.table = c->table + (size_t)b0 * L * BE_REC; a.acc = c->acc + (size_t)b0 * HW * 8;
        a.img.p = dev_img_ny + (size_t)b0 * 6 * HW; a.img.sb = 6 * HW; a.img.sm = 3 * HW; a.img.sc = 1; a.img.sy = 3 * g.W; a.img.sx = 3;   // [B,2,H,W,3]
        a.zgt = dev_bndry_depth + (size_t)b0 * HW; a.mask_count = reinterpret_cast<unsigned long long*>(dev_mask_count);
        a.crec = c->crec + (size_t)b0 * L * BE_CREC;
        a.g = g; a.cam = c->cam; a.NB = nb; a.accH = g.H; a.accW = g.W;
        pick_runs(g, nb, RUN_CTAS, RUN_OVH, &a.G, &a.runs_per_row);
        if (det) {
            a.G = g.Wp; a.runs_per_row = 1;
            a.stage = c->stage + (size_t)b0 * g.Hp * g.R * g.W * 8;
        }
        launch_run(BE_RUN_TRAINFWD, a, st);
        if (tm) cudaEventRecord(c->tev[3], st);
        if (det) be_launch_stage_reduce(a.stage, g, nb, 8, c->acc + (size_t)b0 * HW * 8, st);
    }
    if (parts & 2) {   // targets: global maps + packed per-pixel targets of the loss kernel
        if (tm && !(parts & 1)) cudaEventRecord(c->tev[3], st);
        // one kernel: normalised global maps + packed per-pixel targets (the timing hook keeps its slot for the former normalise
        // launch: it now reads ~0)
        if (tm) cudaEventRecord(c->tev[4], st);
        be_launch_train_targets(c->acc, g, b0, nb, Btot, dev_img_ny, dev_img_gt, dev_bndry_dist, dev_deri, dev_bndry_depth, c->T,
                                dev_global_image, dev_global_bndry, st);
        if (tm) cudaEventRecord(c->tev[5], st);
    }
    BE_CUDA(cudaGetLastError());
    return 0;
}

static int global_loss_stage1_parts(be_ctx* c, const float* dev_raw, const float* dev_img_ny, const float* dev_img_gt,
                                    const float* dev_bndry_dist, const float* dev_deri, const float* dev_bndry_depth, int32_t B,
                                    float* dev_global_image, float* dev_global_bndry, int64_t* dev_mask_count, void* stream, int parts) {
    if (check_ctx(c)) return 1;
    if (B == 0) return 0;
    BE_REQUIRE(dev_raw && dev_img_ny && dev_img_gt && dev_bndry_dist && dev_deri && dev_bndry_depth && dev_mask_count, "null pointer");
    BE_REQUIRE(B > 0 && B <= c->cfg.max_batch, "B=%d exceeds max_batch=%d", B, c->cfg.max_batch);
    if (ensure_train_ws(c)) return 1;
    if (c->timing && ensure_train_events(c)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (parts & 1) {
        BE_CUDA(cudaMemsetAsync(dev_mask_count, 0, sizeof(int64_t), st));
        c->same_gt = (dev_img_gt == dev_img_ny);
        c->train_B = B;
        if (c->timing) {                  // a new step: next event set of the ring
            c->tset = (c->tset + 1) % BE_TRAIN_SETS;
            c->tev = c->tev_ring[c->tset];
            ++c->tsteps;
        }
    } else {
        BE_REQUIRE(c->train_B == B, "be_global_loss_stage1_render must run first on the same batch");
    }
    return loss_stage1_range(c, dev_raw, dev_img_ny, dev_img_gt, dev_bndry_dist, dev_deri, dev_bndry_depth, 0, B, B, dev_global_image,
                             dev_global_bndry, dev_mask_count, st, c->timing != 0, parts);
}

int be_global_loss_stage1(be_ctx* c, const float* dev_raw, const float* dev_img_ny, const float* dev_img_gt,
                          const float* dev_bndry_dist, const float* dev_deri, const float* dev_bndry_depth, int32_t B,
                          float* dev_global_image, float* dev_global_bndry, int64_t* dev_mask_count, void* stream) {
    return global_loss_stage1_parts(c, dev_raw, dev_img_ny, dev_img_gt, dev_bndry_dist, dev_deri, dev_bndry_depth, B, dev_global_image,
                                    dev_global_bndry, dev_mask_count, stream, 3);
}

int be_global_loss_stage1_render(be_ctx* c, const float* dev_raw, const float* dev_img_ny, const float* dev_img_gt,
                                 const float* dev_bndry_dist, const float* dev_deri, const float* dev_bndry_depth, int32_t B,
                                 int64_t* dev_mask_count, void* stream) {
    return global_loss_stage1_parts(c, dev_raw, dev_img_ny, dev_img_gt, dev_bndry_dist, dev_deri, dev_bndry_depth, B, nullptr, nullptr,
                                    dev_mask_count, stream, 1);
}

int be_global_loss_stage1_targets(be_ctx* c, const float* dev_raw, const float* dev_img_ny, const float* dev_img_gt,
                                  const float* dev_bndry_dist, const float* dev_deri, const float* dev_bndry_depth, int32_t B,
                                  float* dev_global_image, float* dev_global_bndry, int64_t* dev_mask_count, void* stream) {
    return global_loss_stage1_parts(c, dev_raw, dev_img_ny, dev_img_gt, dev_bndry_dist, dev_deri, dev_bndry_depth, B, dev_global_image,
                                    dev_global_bndry, dev_mask_count, stream, 2);
}

struct LossScales {
    BeLossScale sc;
    float kc, kcc, kbc, ks, ksc, kbl, gamma_d;
};

static LossScales loss_scales(const BeGeom& g, const double* gammas7, int64_t global_patches) {
    LossScales o;
    memset(&o, 0, sizeof(o));
    const double RR = (double)g.R * g.R, Ri2 = (double)(g.R - 2) * (g.R - 2), Np = (double)global_patches;
    o.sc.nterms = 7;
    const double norm[7] = {2 * RR * Np, 2 * RR * Np, RR * Np, 2 * Ri2 * Np, 2 * Ri2 * Np, RR * Np, 1.0};
    for (int t = 0; t < 7; ++t) { o.sc.src[t] = t; o.sc.scale[t] = 1.0 / norm[t]; o.sc.gamma[t] = (float)gammas7[t]; o.sc.masked[t] = (t == 6); }
    o.kc = (float)(gammas7[0] / norm[0]); o.kcc = (float)(gammas7[1] / norm[1]); o.kbc = (float)(gammas7[2] / norm[2]);
    o.ks = (float)(gammas7[3] / norm[3]); o.ksc = (float)(gammas7[4] / norm[4]); o.kbl = (float)(gammas7[5] / norm[5]);
    o.gamma_d = (float)gammas7[6];
    return o;
}

// The loss kernel on pairs [b0, b0 + nb) of a batch of Btot pairs whose stage 1 has run.  dev_grad / dev_grad_depth are the arrays of
// the whole batch; the CTAs' partial sums go to c->partials + part_off rows.  Returns the number of partial rows in *nparts.
static int loss_kernel_range(be_ctx* c, int b0, int nb, int Btot, const LossScales& k, const int64_t* dev_mask_count, float* dev_grad,
                             float* dev_grad_depth, bool defer, int part_off, int* nparts, cudaStream_t st) {
    const BeGeom& g = c->g;
    const size_t L = (size_t)g.Hp * g.Wp;
    BeLossArgs a;
    memset(&a, 0, sizeof(a));
    a.table = c->table + (size_t)b0 * L * BE_REC; a.gtable = c->gtable + (size_t)b0 * L * BE_GREC; a.crec = c->crec + (size_t)b0 * L * BE_CREC;
    a.T = c->T; a.b0 = b0; a.NBT = Btot;
    a.grad = dev_grad ? dev_grad + (size_t)b0 * L * 12 : nullptr;
    a.grad_depth = (defer && dev_grad_depth) ? dev_grad_depth + (size_t)b0 * L * 4 : nullptr;
    a.defer_depth = defer ? 1 : 0;
    a.mask_count = reinterpret_cast<const unsigned long long*>(dev_mask_count);
    a.g = g; a.NB = nb; a.same_gt = c->same_gt;
    pick_runs(g, nb, LOSS_CTAS, LOSS_OVH, &a.G, &a.runs_per_row);
    a.partials = c->partials + (size_t)part_off * 8;
    a.kc = k.kc; a.kcc = k.kcc; a.kbc = k.kbc; a.ks = k.ks; a.ksc = k.ksc; a.kbl = k.kbl; a.gamma_d = k.gamma_d;
    be_launch_loss2(a, st);
    *nparts = nb * g.Hp * a.runs_per_row;
    return 0;
}

// Both halves of stage 2.  `launch`: the loss kernel; `finish`: the partial sums -> terms and loss (needs the mask count), and, when the
// depth share of the gradient was kept apart (dev_grad_depth), its normalisation.  be_global_loss_stage2 = launch + finish with the
// count already known; be_global_loss_stage2_launch / _finish let a data-parallel caller overlap the all-reduce of the count with
// the loss kernel.
static int loss_stage2(be_ctx* c, int32_t B, const double* gammas7, int64_t global_patches, const int64_t* dev_mask_count,
                       const int64_t* dev_true_patches, float* dev_terms, float* dev_loss, float* dev_grad, float* dev_grad_depth, bool launch,
                       bool finish, void* stream) {
    if (check_ctx(c)) return 1;
    if (B == 0) return 0;
    BE_REQUIRE(gammas7, "null pointer");
    BE_REQUIRE(!finish || (dev_mask_count && dev_terms && dev_loss), "null pointer");
    BE_REQUIRE(launch && finish ? true : (dev_grad == nullptr) == (dev_grad_depth == nullptr), "the split stage 2 needs dev_grad and dev_grad_depth together");
    BE_REQUIRE(launch && finish ? dev_mask_count != nullptr : true, "null pointer");
    BE_REQUIRE(B > 0 && B <= c->cfg.max_batch, "B=%d exceeds max_batch=%d", B, c->cfg.max_batch);
    BE_REQUIRE(c->gtable && c->train_B == B, "be_global_loss_stage1 must run first on the same batch (stage 1 saw %d pairs, stage 2 got %d)", c->train_B, B);
    BE_REQUIRE(global_patches > 0, "global_patches must be positive");
    BE_REQUIRE((double)B * c->g.H * c->g.W * BE_TW < 4.0e9, "batch of %d %dx%d pairs exceeds the 32-bit target offsets of the loss kernel", B, c->g.H, c->g.W);
    cudaStream_t st = (cudaStream_t)stream;
    const BeGeom& g = c->g;
    const LossScales k = loss_scales(g, gammas7, global_patches);
    const bool tm = c->timing != 0 && c->tev;
    if (launch) {
        if (tm) cudaEventRecord(c->tev[6], st);
        if (loss_kernel_range(c, 0, B, B, k, dev_mask_count, dev_grad, dev_grad_depth, !finish, 0, &c->train_parts, st)) return 1;
        if (tm) cudaEventRecord(c->tev[7], st);
    }
    if (finish) {
        if (tm) cudaEventRecord(c->tev[8], st);
        const unsigned long long* tp = reinterpret_cast<const unsigned long long*>(dev_true_patches);
        be_launch_loss_reduce(c->partials, c->train_parts, k.sc, reinterpret_cast<const unsigned long long*>(dev_mask_count), tp,
                              (double)global_patches, dev_terms, dev_loss, st);
        if (!launch && dev_grad && dev_grad_depth)
            be_launch_grad_depth_fixup(dev_grad, dev_grad_depth, reinterpret_cast<const unsigned long long*>(dev_mask_count), tp,
                                       (double)global_patches, (size_t)B * g.Hp * g.Wp, st);
        if (tm) cudaEventRecord(c->tev[9], st);
    }
    BE_CUDA(cudaGetLastError());
    return 0;
}

int be_global_loss_stage2(be_ctx* c, int32_t B, const double* gammas7, int64_t global_patches, const int64_t* dev_mask_count,
                          float* dev_terms, float* dev_loss, float* dev_grad, void* stream) {
    return loss_stage2(c, B, gammas7, global_patches, dev_mask_count, nullptr, dev_terms, dev_loss, dev_grad, nullptr, true, true, stream);
}

int be_global_loss_stage2_launch(be_ctx* c, int32_t B, const double* gammas7, int64_t global_patches, float* dev_grad,
                                 float* dev_grad_depth, void* stream) {
    return loss_stage2(c, B, gammas7, global_patches, nullptr, nullptr, nullptr, nullptr, dev_grad, dev_grad_depth, true, false, stream);
}

int be_global_loss_stage2_finish(be_ctx* c, int32_t B, const double* gammas7, int64_t global_patches, const int64_t* dev_mask_count,
                                 const int64_t* dev_true_patches, float* dev_terms, float* dev_loss, float* dev_grad, float* dev_grad_depth,
                                 void* stream) {
    return loss_stage2(c, B, gammas7, global_patches, dev_mask_count, dev_true_patches, dev_terms, dev_loss, dev_grad, dev_grad_depth, false, true,
                       stream);
}

int be_ctx_train_timing_at(be_ctx* c, int32_t steps_back, float* ms7) {
    if (check_ctx(c)) return 1;
    BE_REQUIRE(ms7 && c->timing && c->tev && c->tsteps > 0, "timing is not enabled (be_ctx_set_timing) or no training step has run since");
    BE_REQUIRE(steps_back >= 0 && steps_back < BE_TRAIN_SETS && steps_back < c->tsteps, "only the last %d timed steps are kept (asked for %d back of %lld)",
               BE_TRAIN_SETS, steps_back, c->tsteps);
    cudaEvent_t* ev = c->tev_ring[(c->tset - steps_back + BE_TRAIN_SETS) % BE_TRAIN_SETS];
    BE_CUDA(cudaEventSynchronize(ev[9]));
    static const int from[7] = {0, 1, 2, 3, 4, 6, 8};
    for (int i = 0; i < 7; ++i) BE_CUDA(cudaEventElapsedTime(&ms7[i], ev[from[i]], ev[from[i] + 1]));
    return 0;
}

int be_ctx_last_train_timing(be_ctx* c, float* ms7) { return be_ctx_train_timing_at(c, 0, ms7); }

static cudaEvent_t* g_trace_ev = nullptr;      // BE_HOST_TRACE: events of the last be_host_global_loss_begin
static int g_trace_chunks = 0, g_trace_nb[BE_HOST_CHUNKS] = {0};

// ---- host-buffer form of the training step (the e2e path of bench.py; what a ctypes binding on the reference side calls with numpy
// arrays).  begin: H2D in chunks of pairs on an internal copy stream, per chunk stage 1 and the loss kernel with the depth normaliser
// deferred, on the CALLER's stream (so that a data-parallel caller can all-reduce the count stream-ordered after it);
// end: reduce, depth fix-up, D2H of terms / loss / grad, synchronise. ----
static int ensure_train_staging(be_ctx* c) {
    if (c->ht_raw) return 0;
    const BeGeom& g = c->g;
    const size_t mb = (size_t)c->cfg.max_batch, HW = (size_t)g.H * g.W, L = (size_t)g.Hp * g.Wp, f = sizeof(float);
    const size_t dHW = (size_t)(g.H - 2) * (g.W - 2);
    BE_CUDA(cudaMalloc(&c->ht_raw, mb * L * 12 * f));
    BE_CUDA(cudaMalloc(&c->ht_ny, mb * 6 * HW * f));
    BE_CUDA(cudaMalloc(&c->ht_gt, mb * 6 * HW * f));
    BE_CUDA(cudaMalloc(&c->ht_bd, mb * HW * f));
    BE_CUDA(cudaMalloc(&c->ht_deri, mb * 6 * dHW * f));
    BE_CUDA(cudaMalloc(&c->ht_zg, mb * HW * f));
    BE_CUDA(cudaMalloc(&c->ht_grad, mb * L * 12 * f));
    BE_CUDA(cudaMalloc(&c->ht_gdep, mb * L * 4 * f));
    BE_CUDA(cudaMalloc(&c->ht_scal, 8 * f + sizeof(int64_t) * 2));
    c->st_bytes += mb * (L * 28 + 14 * HW + 6 * dHW) * f + 48;
    if (!c->st_streams[0])
        for (int i = 0; i < 3; ++i) BE_CUDA(cudaStreamCreateWithFlags(&c->st_streams[i], cudaStreamNonBlocking));
    if (!c->st_events[0])
        for (int i = 0; i < 2 * BE_HOST_CHUNKS; ++i) BE_CUDA(cudaEventCreateWithFlags(&c->st_events[i], cudaEventDisableTiming));
    return 0;
}

int be_host_global_loss_begin(be_ctx* c, const float* raw, const float* img_ny, const float* img_gt, const float* bndry_dist,
                              const float* deri, const float* bndry_depth, int32_t B, const double* gammas7, int64_t global_patches,
                              int32_t want_grad, int64_t* dev_mask_count, void* stream) {
    if (check_ctx(c)) return 1;
    BE_REQUIRE(raw && img_ny && img_gt && bndry_dist && deri && bndry_depth && gammas7 && dev_mask_count, "null pointer");
    BE_REQUIRE(B > 0 && B <= c->cfg.max_batch, "B=%d must be in [1, max_batch=%d]", B, c->cfg.max_batch);
    BE_REQUIRE(global_patches > 0, "global_patches must be positive");
    BE_REQUIRE((double)B * c->g.H * c->g.W * BE_TW < 4.0e9, "batch of %d %dx%d pairs exceeds the 32-bit target offsets of the loss kernel", B, c->g.H, c->g.W);
    if (ensure_train_ws(c) || ensure_train_staging(c)) return 1;
    const BeGeom& g = c->g;
    const size_t HW = (size_t)g.H * g.W, L = (size_t)g.Hp * g.Wp, f = sizeof(float), dHW = (size_t)(g.H - 2) * (g.W - 2);
    cudaStream_t s_k = (cudaStream_t)stream, s_in = c->st_streams[0];
    const bool same = (img_gt == img_ny);
    c->same_gt = same;
    c->train_B = B;
    const float* d_gt = same ? c->ht_ny : c->ht_gt;
    const LossScales k = loss_scales(g, gammas7, global_patches);
    // the copy stream must not overtake the previous call's kernels, which may still read the staging buffers
    BE_CUDA(cudaEventRecord(c->st_events[2 * BE_HOST_CHUNKS - 1], s_k));
    BE_CUDA(cudaStreamWaitEvent(s_in, c->st_events[2 * BE_HOST_CHUNKS - 1], 0));
    BE_CUDA(cudaMemsetAsync(dev_mask_count, 0, sizeof(int64_t), s_k));
    // Chunks grow 1:3:5:7: a small first chunk gets the kernels going after ~6 % of the H2D traffic, the later chunks are large enough
    // to fill whole waves of CTAs (a chunk of n pairs costs at least one wave: ~0.2 ms), and every chunk's copy is done before the
    // kernels of the previous ones are (H2D 45 MB ~ 0.9 ms, kernels 2.9 ms at 32 pairs).  BE_HOST_TRAIN_WGT="a,b,c,..." overrides.
    static int wgt[BE_HOST_CHUNKS] = {1, 3, 5, 7};
    static int nw = 4;
    static const bool parsed = [] {
        const char* e = getenv("BE_HOST_TRAIN_WGT");
        if (e) {
            int k = 0;
            while (*e && k < BE_HOST_CHUNKS - 1) { const int v = atoi(e); wgt[k++] = v < 1 ? 1 : v; while (*e && *e != ',') ++e; if (*e == ',') ++e; }
            if (k > 0) nw = k;
        }
        return true;
    }();
    (void)parsed;
    int bounds[BE_HOST_CHUNKS + 1];
    int nchunk = nw;
    {
        int tot_w = 0, acc_w = 0;
        for (int i = 0; i < nw; ++i) tot_w += wgt[i];
        bounds[0] = 0;
        for (int i = 0; i < nw; ++i) { acc_w += wgt[i]; bounds[i + 1] = (int)(((long long)B * acc_w + tot_w / 2) / tot_w); }
        bounds[nw] = B;
    }
    int part_off = 0;
    // Chunks alternate between the caller's stream and an internal one: the last, partly filled wave of chunk i's loss kernel then
    // shares the GPU with chunk i+1's stage-1 kernels instead of leaving SMs idle (a chunk of n pairs is a whole number of waves
    // only by accident).  The internal stream starts after the caller's stream reaches this call and is joined back at the end.
    static const int nstreams = [] { const char* e = getenv("BE_HOST_TRAIN_STREAMS"); const int v = e ? atoi(e) : 2; return v == 1 ? 1 : 2; }();
    cudaStream_t s_alt = c->st_streams[2];
    if (nstreams == 2) {
        BE_CUDA(cudaEventRecord(c->st_events[2 * BE_HOST_CHUNKS - 2], s_k));         // after the memset of the mask count
        BE_CUDA(cudaStreamWaitEvent(s_alt, c->st_events[2 * BE_HOST_CHUNKS - 2], 0));
    }
    static const bool trace = getenv("BE_HOST_TRACE") != nullptr;       // timeline of one call on stderr (tuning aid)
    static cudaEvent_t tev[3 * BE_HOST_CHUNKS + 2];
    if (trace && !tev[0]) for (auto& e : tev) cudaEventCreate(&e);
    if (trace) { cudaEventRecord(tev[3 * BE_HOST_CHUNKS], s_in); g_trace_chunks = nchunk; }
    for (int i = 0; i < nchunk; ++i) {
        const int b0 = bounds[i], b1 = bounds[i + 1], nb = b1 - b0;
        if (nb <= 0) continue;
        BE_CUDA(cudaMemcpyAsync(c->ht_raw + b0 * L * 12, raw + b0 * L * 12, nb * L * 12 * f, cudaMemcpyHostToDevice, s_in));
        BE_CUDA(cudaMemcpyAsync(c->ht_ny + b0 * 6 * HW, img_ny + b0 * 6 * HW, nb * 6 * HW * f, cudaMemcpyHostToDevice, s_in));
        if (!same) BE_CUDA(cudaMemcpyAsync(c->ht_gt + b0 * 6 * HW, img_gt + b0 * 6 * HW, nb * 6 * HW * f, cudaMemcpyHostToDevice, s_in));
        BE_CUDA(cudaMemcpyAsync(c->ht_bd + b0 * HW, bndry_dist + b0 * HW, nb * HW * f, cudaMemcpyHostToDevice, s_in));
        BE_CUDA(cudaMemcpyAsync(c->ht_deri + b0 * 6 * dHW, deri + b0 * 6 * dHW, nb * 6 * dHW * f, cudaMemcpyHostToDevice, s_in));
        BE_CUDA(cudaMemcpyAsync(c->ht_zg + b0 * HW, bndry_depth + b0 * HW, nb * HW * f, cudaMemcpyHostToDevice, s_in));
        BE_CUDA(cudaEventRecord(c->st_events[i], s_in));
        if (trace) cudaEventRecord(tev[3 * i], s_in);
        cudaStream_t s_c = (nstreams == 2 && (i & 1)) ? s_alt : s_k;
        BE_CUDA(cudaStreamWaitEvent(s_c, c->st_events[i], 0));
        if (loss_stage1_range(c, c->ht_raw, c->ht_ny, d_gt, c->ht_bd, c->ht_deri, c->ht_zg, b0, nb, B, nullptr, nullptr, dev_mask_count, s_c, false))
            return 1;
        if (trace) cudaEventRecord(tev[3 * i + 1], s_c);
        int np_ = 0;
        if (loss_kernel_range(c, b0, nb, B, k, nullptr, want_grad ? c->ht_grad : nullptr, want_grad ? c->ht_gdep : nullptr, true, part_off, &np_, s_c))
            return 1;
        if (trace) { cudaEventRecord(tev[3 * i + 2], s_c); g_trace_ev = tev; g_trace_nb[i] = nb; }
        part_off += np_;
    }
    if (nstreams == 2) {                 // join: everything after this call on the caller's stream sees all chunks
        BE_CUDA(cudaEventRecord(c->st_events[2 * BE_HOST_CHUNKS - 3], s_alt));
        BE_CUDA(cudaStreamWaitEvent(s_k, c->st_events[2 * BE_HOST_CHUNKS - 3], 0));
    }
    c->train_parts = part_off;
    c->ht_want_grad = want_grad;
    BE_CUDA(cudaGetLastError());
    return 0;
}

int be_host_global_loss_end(be_ctx* c, int32_t B, const double* gammas7, int64_t global_patches, const int64_t* dev_mask_count,
                            const int64_t* dev_true_patches, float* terms7, float* loss1, float* grad, void* stream) {
    if (check_ctx(c)) return 1;
    BE_REQUIRE(gammas7 && dev_mask_count && terms7 && loss1, "null pointer");
    BE_REQUIRE(c->ht_raw && c->train_B == B && B > 0, "be_host_global_loss_begin must run first on the same batch");
    BE_REQUIRE(!grad || c->ht_want_grad, "be_host_global_loss_begin ran with want_grad = 0");
    const BeGeom& g = c->g;
    const size_t L = (size_t)g.Hp * g.Wp;
    cudaStream_t s_k = (cudaStream_t)stream;
    const LossScales k = loss_scales(g, gammas7, global_patches);
    float* d_terms = c->ht_scal;
    const unsigned long long* tp = reinterpret_cast<const unsigned long long*>(dev_true_patches);
    be_launch_loss_reduce(c->partials, c->train_parts, k.sc, reinterpret_cast<const unsigned long long*>(dev_mask_count), tp, (double)global_patches,
                          d_terms, d_terms + 7, s_k);
    BE_CUDA(cudaMemcpyAsync(terms7, d_terms, 7 * sizeof(float), cudaMemcpyDeviceToHost, s_k));
    BE_CUDA(cudaMemcpyAsync(loss1, d_terms + 7, sizeof(float), cudaMemcpyDeviceToHost, s_k));
    if (grad) {
        be_launch_grad_depth_fixup(c->ht_grad, c->ht_gdep, reinterpret_cast<const unsigned long long*>(dev_mask_count), tp, (double)global_patches,
                                   (size_t)B * L, s_k);
        BE_CUDA(cudaMemcpyAsync(grad, c->ht_grad, (size_t)B * L * 12 * sizeof(float), cudaMemcpyDeviceToHost, s_k));
    }
    BE_CUDA(cudaGetLastError());
    if (g_trace_ev) cudaEventRecord(g_trace_ev[3 * BE_HOST_CHUNKS + 1], s_k);
    BE_CUDA(cudaStreamSynchronize(s_k));
    if (g_trace_ev) {
        cudaEvent_t* tev = g_trace_ev;
        float a, b2, d;
        for (int i = 0; i < g_trace_chunks; ++i) {
            cudaEventElapsedTime(&a, tev[3 * BE_HOST_CHUNKS], tev[3 * i]); cudaEventElapsedTime(&b2, tev[3 * BE_HOST_CHUNKS], tev[3 * i + 1]);
            cudaEventElapsedTime(&d, tev[3 * BE_HOST_CHUNKS], tev[3 * i + 2]);
            fprintf(stderr, "train chunk %d (%d pairs): h2d done %.3f  stage 1 done %.3f  loss kernel done %.3f ms\n", i, g_trace_nb[i], a, b2, d);
        }
        cudaEventElapsedTime(&a, tev[3 * BE_HOST_CHUNKS], tev[3 * BE_HOST_CHUNKS + 1]);
        fprintf(stderr, "reduce + fix-up + d2h done %.3f ms\n", a);
    }
    return 0;
}

// Single-process form of the step on host buffers, two phases.  The depth term's normaliser is the mask count of the WHOLE batch
// (global_training.py:127), which the device knows as soon as every chunk is rendered.  Phase 1 therefore renders chunk after chunk
// behind the H2D copies of (est, image pair, z_gt); the copies of the loss-only inputs (bndry_dist, deri) follow on the same copy
// stream.  Phase 2 packs the targets and runs the loss kernel per chunk with the FINAL count (no deferred normaliser, no fix-up pass),
// so a chunk's gradient is final when its loss kernel ends and goes home while the next chunk's loss kernel runs: what is left after
// the last kernel is the D2H of the last - smallest - chunk.  (The deferred form below, be_host_global_loss_begin/_end, stays for
// data-parallel callers, whose count is only final after an all-reduce; it pays reduce + fix-up + the D2H of the whole gradient,
// ~0.14 ms at 32 pairs, after the last kernel.)  BE_HOST_TRAIN_MODE=deferred selects the old schedule for comparison.
static int host_global_loss_two_phase(be_ctx* c, const float* raw, const float* img_ny, const float* img_gt, const float* bndry_dist,
                                      const float* deri, const float* bndry_depth, int32_t B, const double* gammas7, float* terms7,
                                      float* loss1, float* grad) {
    BE_REQUIRE(raw && img_ny && img_gt && bndry_dist && deri && bndry_depth && gammas7 && terms7 && loss1, "null pointer");
    BE_REQUIRE(B > 0 && B <= c->cfg.max_batch, "B=%d must be in [1, max_batch=%d]", B, c->cfg.max_batch);
    BE_REQUIRE((double)B * c->g.H * c->g.W * BE_TW < 4.0e9, "batch of %d %dx%d pairs exceeds the 32-bit target offsets of the loss kernel", B, c->g.H, c->g.W);
    const BeGeom& g = c->g;
    const size_t HW = (size_t)g.H * g.W, L = (size_t)g.Hp * g.Wp, f = sizeof(float), dHW = (size_t)(g.H - 2) * (g.W - 2);
    cudaStream_t s_in = c->st_streams[0], s_k = c->st_streams[1], s_alt = c->st_streams[2];
    int64_t* cnt = reinterpret_cast<int64_t*>(c->ht_scal + 8);
    const bool same = (img_gt == img_ny);
    c->same_gt = same;
    c->train_B = B;
    const float* d_gt = same ? c->ht_ny : c->ht_gt;
    const LossScales k = loss_scales(g, gammas7, (int64_t)B * (int64_t)L);
    constexpr int MAXC = 8;                                  // events: [i] render inputs of chunk i, [8 + j] loss inputs of loss chunk j, [16 + j] loss kernel j done
    // Phase 1 is paced by the copies (a pair's render inputs take about as long to arrive as the pair takes to render): a small first
    // chunk, then equal ones.  The loss-only inputs follow in the order phase 2 consumes them; its first chunk is half the batch (those
    // inputs are up when the last render ends), the last one is small (its gradient's way home is all that is left after the last kernel).
    // Measured at 32 pairs: 3.02 ms per call against 3.20 ms for the deferred schedule (profiles/r2_experiments.txt).
    static int wgt[MAXC] = {1, 3, 4, 4, 4}, wgt2[MAXC] = {8, 7, 1};
    static int nw = 5, nw2 = 3;
    static const bool parsed = [] {
        auto parse = [](const char* name, int* w, int* n) {
            const char* e = getenv(name);
            if (!e) return;
            int k = 0;
            while (*e && k < MAXC) { const int v = atoi(e); w[k++] = v < 1 ? 1 : v; while (*e && *e != ',') ++e; if (*e == ',') ++e; }
            if (k > 0) *n = k;
        };
        parse("BE_HOST_TRAIN_WGT", wgt, &nw);
        parse("BE_HOST_TRAIN_WGT2", wgt2, &nw2);
        return true;
    }();
    (void)parsed;
    int bounds[MAXC + 1], bounds2[MAXC + 1];
    auto split = [B](const int* w, int n, int* bd) {
        int tot_w = 0, acc_w = 0;
        for (int i = 0; i < n; ++i) tot_w += w[i];
        bd[0] = 0;
        for (int i = 0; i < n; ++i) { acc_w += w[i]; bd[i + 1] = (int)(((long long)B * acc_w + tot_w / 2) / tot_w); }
        bd[n] = B;
    };
    split(wgt, nw, bounds);
    split(wgt2, nw2, bounds2);
    static const bool trace = getenv("BE_HOST_TRACE") != nullptr;
    static cudaEvent_t tev[4 * MAXC + 2];
    if (trace && !tev[0]) for (auto& e : tev) cudaEventCreate(&e);
    cudaEvent_t* ev = c->st_events;
    BE_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int64_t), s_k));
    BE_CUDA(cudaEventRecord(ev[31], s_k));
    BE_CUDA(cudaStreamWaitEvent(s_alt, ev[31], 0));
    if (trace) cudaEventRecord(tev[4 * MAXC], s_in);
    // ---- phase 1: render inputs up (small chunk first), render behind them; then the loss-only inputs (big chunk first) ----
    for (int i = 0; i < nw; ++i) {
        const int b0 = bounds[i], nb = bounds[i + 1] - b0;
        if (nb <= 0) continue;
        BE_CUDA(cudaMemcpyAsync(c->ht_raw + b0 * L * 12, raw + b0 * L * 12, nb * L * 12 * f, cudaMemcpyHostToDevice, s_in));
        BE_CUDA(cudaMemcpyAsync(c->ht_ny + b0 * 6 * HW, img_ny + b0 * 6 * HW, nb * 6 * HW * f, cudaMemcpyHostToDevice, s_in));
        if (!same) BE_CUDA(cudaMemcpyAsync(c->ht_gt + b0 * 6 * HW, img_gt + b0 * 6 * HW, nb * 6 * HW * f, cudaMemcpyHostToDevice, s_in));
        BE_CUDA(cudaMemcpyAsync(c->ht_zg + b0 * HW, bndry_depth + b0 * HW, nb * HW * f, cudaMemcpyHostToDevice, s_in));
        BE_CUDA(cudaEventRecord(ev[i], s_in));
        cudaStream_t s_c = (i & 1) ? s_alt : s_k;
        BE_CUDA(cudaStreamWaitEvent(s_c, ev[i], 0));
        if (loss_stage1_range(c, c->ht_raw, c->ht_ny, d_gt, c->ht_bd, c->ht_deri, c->ht_zg, b0, nb, B, nullptr, nullptr, cnt, s_c, false, 1)) return 1;
        if (trace) { cudaEventRecord(tev[4 * i], s_in); cudaEventRecord(tev[4 * i + 1], s_c); }
    }
    for (int j = 0; j < nw2; ++j) {                          // loss-only inputs, in the order phase 2 consumes them
        const int b0 = bounds2[j], nb = bounds2[j + 1] - b0;
        if (nb <= 0) continue;
        BE_CUDA(cudaMemcpyAsync(c->ht_bd + b0 * HW, bndry_dist + b0 * HW, nb * HW * f, cudaMemcpyHostToDevice, s_in));
        BE_CUDA(cudaMemcpyAsync(c->ht_deri + b0 * 6 * dHW, deri + b0 * 6 * dHW, nb * 6 * dHW * f, cudaMemcpyHostToDevice, s_in));
        BE_CUDA(cudaEventRecord(ev[8 + j], s_in));
    }
    // every chunk rendered = the mask count is final: both kernel streams wait for each other
    BE_CUDA(cudaEventRecord(ev[30], s_alt));
    BE_CUDA(cudaStreamWaitEvent(s_k, ev[30], 0));
    BE_CUDA(cudaEventRecord(ev[29], s_k));
    BE_CUDA(cudaStreamWaitEvent(s_alt, ev[29], 0));
    // ---- phase 2: targets + loss kernel per chunk, gradient home behind each ----
    // loss chunks alternate between the two kernel streams: the partly filled last wave of one shares the GPU with the start of the next
    static const int streams2 = [] { const char* e = getenv("BE_HOST_TRAIN_STREAMS2"); return (e && atoi(e) == 1) ? 1 : 2; }();
    int part_off = 0;
    for (int j = 0; j < nw2; ++j) {
        const int b0 = bounds2[j], nb = bounds2[j + 1] - b0;
        if (nb <= 0) continue;
        cudaStream_t s_c = (streams2 == 2 && (j & 1)) ? s_alt : s_k;
        BE_CUDA(cudaStreamWaitEvent(s_c, ev[8 + j], 0));
        if (loss_stage1_range(c, c->ht_raw, c->ht_ny, d_gt, c->ht_bd, c->ht_deri, c->ht_zg, b0, nb, B, nullptr, nullptr, cnt, s_c, false, 2)) return 1;
        int np_ = 0;
        if (loss_kernel_range(c, b0, nb, B, k, cnt, grad ? c->ht_grad : nullptr, nullptr, false, part_off, &np_, s_c)) return 1;
        part_off += np_;
        if (trace) cudaEventRecord(tev[4 * j + 2], s_c);
        if (grad) {
            BE_CUDA(cudaEventRecord(ev[16 + j], s_c));
            BE_CUDA(cudaStreamWaitEvent(s_in, ev[16 + j], 0));
            BE_CUDA(cudaMemcpyAsync(grad + b0 * L * 12, c->ht_grad + b0 * L * 12, nb * L * 12 * f, cudaMemcpyDeviceToHost, s_in));
            if (trace) cudaEventRecord(tev[4 * j + 3], s_in);
        }
    }
    BE_CUDA(cudaEventRecord(ev[30], s_alt));
    BE_CUDA(cudaStreamWaitEvent(s_k, ev[30], 0));
    c->train_parts = part_off;
    float* d_terms = c->ht_scal;
    be_launch_loss_reduce(c->partials, part_off, k.sc, reinterpret_cast<const unsigned long long*>(cnt), nullptr, (double)B * (double)L, d_terms,
                          d_terms + 7, s_k);
    BE_CUDA(cudaMemcpyAsync(terms7, d_terms, 7 * sizeof(float), cudaMemcpyDeviceToHost, s_k));
    BE_CUDA(cudaMemcpyAsync(loss1, d_terms + 7, sizeof(float), cudaMemcpyDeviceToHost, s_k));
    if (trace) cudaEventRecord(tev[4 * MAXC + 1], s_k);
    BE_CUDA(cudaGetLastError());
    BE_CUDA(cudaStreamSynchronize(s_k));
    BE_CUDA(cudaStreamSynchronize(s_in));
    if (trace) {
        float a[2];
        for (int i = 0; i < nw; ++i) {
            if (bounds[i + 1] <= bounds[i]) continue;
            cudaEventElapsedTime(&a[0], tev[4 * MAXC], tev[4 * i]); cudaEventElapsedTime(&a[1], tev[4 * MAXC], tev[4 * i + 1]);
            fprintf(stderr, "render chunk %d (%d pairs): inputs up %.3f  rendered %.3f ms\n", i, bounds[i + 1] - bounds[i], a[0], a[1]);
        }
        for (int j = 0; j < nw2; ++j) {
            if (bounds2[j + 1] <= bounds2[j]) continue;
            a[1] = -1.f;
            cudaEventElapsedTime(&a[0], tev[4 * MAXC], tev[4 * j + 2]);
            if (grad) cudaEventElapsedTime(&a[1], tev[4 * MAXC], tev[4 * j + 3]);
            fprintf(stderr, "loss chunk %d (%d pairs): loss kernel done %.3f  gradient home %.3f ms\n", j, bounds2[j + 1] - bounds2[j], a[0], a[1]);
        }
        cudaEventElapsedTime(&a[0], tev[4 * MAXC], tev[4 * MAXC + 1]);
        fprintf(stderr, "reduce + terms home %.3f ms\n", a[0]);
    }
    return 0;
}

int be_host_global_loss(be_ctx* c, const float* raw, const float* img_ny, const float* img_gt, const float* bndry_dist, const float* deri,
                        const float* bndry_depth, int32_t B, const double* gammas7, float* terms7, float* loss1, float* grad) {
    if (check_ctx(c)) return 1;
    if (B == 0) return 0;
    if (ensure_train_ws(c) || ensure_train_staging(c)) return 1;
    static const bool deferred = [] { const char* e = getenv("BE_HOST_TRAIN_MODE"); return e && !strcmp(e, "deferred"); }();
    if (!deferred) return host_global_loss_two_phase(c, raw, img_ny, img_gt, bndry_dist, deri, bndry_depth, B, gammas7, terms7, loss1, grad);
    int64_t* cnt = reinterpret_cast<int64_t*>(c->ht_scal + 8);
    const int64_t np_ = (int64_t)B * c->g.Hp * c->g.Wp;
    if (be_host_global_loss_begin(c, raw, img_ny, img_gt, bndry_dist, deri, bndry_depth, B, gammas7, np_, grad != nullptr, cnt, c->st_streams[1])) return 1;
    return be_host_global_loss_end(c, B, gammas7, np_, cnt, nullptr, terms7, loss1, grad, c->st_streams[1]);
}

int be_local_loss(be_ctx* c, float* dev_est, const float* dev_img_ny, const float* dev_img_gt, const float* dev_bndry_dist,
                  const float* dev_deri, int32_t B, double beta_bndry_loc, double beta_smthns, int32_t wrap_in_place, float* dev_terms,
                  float* dev_loss, float* dev_grad, void* stream) {
    if (check_ctx(c)) return 1;
    if (B == 0) return 0;
    BE_REQUIRE(dev_est && dev_img_ny && dev_img_gt && dev_bndry_dist && dev_deri && dev_terms && dev_loss, "null pointer");
    const BeGeom& g = c->g;
    BE_REQUIRE(g.H == g.R && g.W == g.R, "be_local_loss needs a context whose image is one %dx%d patch (got %dx%d)", g.R, g.R, g.H, g.W);
    BE_REQUIRE(B > 0 && B <= c->cfg.max_batch, "B=%d exceeds max_batch=%d", B, c->cfg.max_batch);
    if (!c->lpart) {   // per-CTA partial sums + the ticket of the last-CTA reduction
        BE_CUDA(cudaMalloc(&c->lpart, (size_t)c->cfg.max_batch * 8 * sizeof(float) + 16));
        BE_CUDA(cudaMemset(c->lpart, 0, (size_t)c->cfg.max_batch * 8 * sizeof(float) + 16));
    }
    cudaStream_t st = (cudaStream_t)stream;
    const double RR = (double)g.R * g.R, Ri2 = (double)(g.R - 2) * (g.R - 2), Np = (double)B;
    BeLossArgs a;
    memset(&a, 0, sizeof(a));
    a.grad = dev_grad; a.partials = c->lpart;
    a.l_est = dev_est; a.l_ny = dev_img_ny; a.l_gt = dev_img_gt; a.l_bd = dev_bndry_dist; a.l_deri = dev_deri;
    a.g = g; a.NB = B; a.G = 1; a.runs_per_row = 1;
    BeLocalTail t;
    memset(&t, 0, sizeof(t));
    t.cam = c->cam;
    t.sc.nterms = 3;                                       // loss = colour + beta_loc * loc + beta_smth * smth (local_training.py:47-52)
    t.sc.src[0] = 0; t.sc.src[1] = 5; t.sc.src[2] = 3;
    t.sc.scale[0] = 1.0 / (RR * Np); t.sc.scale[1] = 1.0 / (RR * Np); t.sc.scale[2] = 1.0 / (Ri2 * Np);
    t.sc.gamma[0] = 1.0f; t.sc.gamma[1] = (float)beta_bndry_loc; t.sc.gamma[2] = (float)beta_smthns;
    a.kc = (float)t.sc.scale[0]; a.kbl = (float)(beta_bndry_loc * t.sc.scale[1]); a.ks = (float)(beta_smthns * t.sc.scale[2]);
    t.est_wrapped = wrap_in_place ? dev_est : nullptr;
    t.ticket = reinterpret_cast<unsigned*>(c->lpart + (size_t)c->cfg.max_batch * 8);
    t.terms = dev_terms; t.loss = dev_loss;
    be_launch_loss(a, t, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// blocked (big-image) entry points: blurry_edges_test_big.py:135-190 without the unfolded full_* tensors
// ---------------------------------------------------------------------------------------------------
static int upload_blocks(be_ctx* c, const be_block* host_blocks, int n, cudaStream_t st) {
    const BeGeom& g = c->g;
    for (int i = 0; i < n; ++i) {
        const be_block& h = host_blocks[i];
        BE_REQUIRE(h.py0 >= 0 && h.py1 <= g.Hp && h.px0 >= 0 && h.px1 <= g.Wp && h.py0 <= h.py1 && h.px0 <= h.px1,
                   "block %d: patch window [%d,%d)x[%d,%d) outside the %dx%d patch grid", i, h.py0, h.py1, h.px0, h.px1, g.Hp, g.Wp);
        BE_REQUIRE(h.oy >= 0 && h.ox >= 0 && h.img >= 0, "block %d: negative origin or image index", i);
    }
    // The same descriptors as last time (a block-sharded rank renders the same blocks of every image): they are already on the
    // device - no copy, no host synchronisation, and the call can be captured into a CUDA graph.
    if (n == c->blk_n && c->blk_pin) {
        bool same = true;
        for (int i = 0; i < n && same; ++i) {
            const be_block& h = host_blocks[i];
            const BeBlock& d = c->blk_pin[i];
            same = d.img == h.img && d.oy == h.oy && d.ox == h.ox && d.py0 == h.py0 && d.py1 == h.py1 && d.px0 == h.px0 && d.px1 == h.px1;
        }
        if (same) return 0;
    }
    if (n > c->blk_cap) {
        if (c->blk_ev) BE_CUDA(cudaEventSynchronize(c->blk_ev));
        cudaFree(c->blk_dev); cudaFreeHost(c->blk_pin);
        c->blk_dev = nullptr; c->blk_pin = nullptr; c->blk_cap = 0;
        BE_CUDA(cudaMalloc(&c->blk_dev, (size_t)n * sizeof(BeBlock)));
        BE_CUDA(cudaMallocHost(&c->blk_pin, (size_t)n * sizeof(BeBlock)));
        c->blk_cap = n;
    }
    if (!c->blk_ev) BE_CUDA(cudaEventCreateWithFlags(&c->blk_ev, cudaEventDisableTiming));
    else BE_CUDA(cudaEventSynchronize(c->blk_ev));        // the previous upload must have left the pinned buffer
    for (int i = 0; i < n; ++i) {
        const be_block& h = host_blocks[i];
        BeBlock& d = c->blk_pin[i];
        d.img = h.img; d.oy = h.oy; d.ox = h.ox; d.py0 = h.py0; d.py1 = h.py1; d.px0 = h.px0; d.px1 = h.px1; d.pad = 0;
    }
    c->blk_n = n;
    BE_CUDA(cudaMemcpyAsync(c->blk_dev, c->blk_pin, (size_t)n * sizeof(BeBlock), cudaMemcpyHostToDevice, st));
    BE_CUDA(cudaEventRecord(c->blk_ev, st));
    return 0;
}

int be_colors_blocks_fwd(be_ctx* c, const float* dev_est, int32_t param_mode, const float* dev_img, const be_image_layout* layout,
                         const be_block* blocks, int32_t nitem, float* dev_colors, void* stream) {
    if (check_ctx(c) || check_layout(layout)) return 1;
    if (nitem == 0) return 0;
    BE_REQUIRE(dev_est && dev_img && dev_colors && blocks, "null pointer");
    BE_REQUIRE(param_mode == BE_PARAMS_LOCAL10 || param_mode == BE_PARAMS_LOCALRAW10, "be_colors_blocks_fwd takes 10-parameter patches");
    BE_REQUIRE(nitem > 0 && nitem <= 2 * c->cfg.max_batch, "nitem=%d exceeds 2*max_batch=%d", nitem, 2 * c->cfg.max_batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (upload_blocks(c, blocks, nitem, st)) return 1;
    const int L = c->g.Hp * c->g.Wp;
    be_launch_setup(dev_est, param_mode, nitem * L, c->cam, c->table, nullptr, st);
    BeRunArgs a;
    memset(&a, 0, sizeof(a));
    a.table = c->table; a.img = make_img(dev_img, layout); a.colors = dev_colors; a.blocks = c->blk_dev;
    a.g = c->g; a.cam = c->cam; a.NB = nitem; a.accH = c->g.H; a.accW = c->g.W;
    pick_runs(c->g, nitem, COLORS_CTAS, RUN_OVH, &a.G, &a.runs_per_row);
    launch_run(BE_RUN_COLORS, a, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}

int be_render_fold_blocks(be_ctx* c, const float* dev_est, int32_t param_mode, const float* dev_img, const be_image_layout* layout,
                          const be_block* blocks, int32_t nblk, int32_t densify_w, int32_t acc_y0, int32_t acc_H, int32_t acc_W, float* dev_acc,
                          void* stream) {
    if (check_ctx(c) || check_layout(layout)) return 1;
    if (nblk == 0) return 0;
    BE_REQUIRE(dev_est && dev_img && dev_acc && blocks, "null pointer");
    BE_REQUIRE(param_mode == BE_PARAMS_RESTORED12 || param_mode == BE_PARAMS_RAW12, "be_render_fold_blocks takes 12-parameter patches");
    BE_REQUIRE(nblk > 0 && nblk <= 2 * c->cfg.max_batch, "nblk=%d exceeds 2*max_batch=%d", nblk, 2 * c->cfg.max_batch);
    BE_REQUIRE(!c->deterministic, "be_render_fold_blocks has no deterministic (fixed-order) fold: blocks of one image overlap in the accumulator");
    const BeGeom& g = c->g;
    for (int i = 0; i < nblk; ++i) {   // the rows a block WRITES: those of its patch window [py0, py1)
        if (blocks[i].py1 <= blocks[i].py0) continue;
        const int r0 = blocks[i].oy + blocks[i].py0 * g.stride, r1 = blocks[i].oy + (blocks[i].py1 - 1) * g.stride + g.R;
        BE_REQUIRE(r0 >= acc_y0 && r1 <= acc_y0 + acc_H && blocks[i].ox + g.W <= acc_W,
                   "block %d (writes rows %d..%d) does not fit the accumulator (rows %d..%d, %d columns)", i, r0, r1, acc_y0, acc_y0 + acc_H, acc_W);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (upload_blocks(c, blocks, nblk, st)) return 1;
    const int L = g.Hp * g.Wp;
    be_launch_setup(dev_est, param_mode, nblk * L, c->cam, c->table, nullptr, st);
    BeRunArgs a;
    memset(&a, 0, sizeof(a));
    // the accumulator holds image rows [acc_y0, acc_y0 + acc_H): the kernel addresses it by image row
    a.table = c->table; a.img = make_img(dev_img, layout); a.acc = dev_acc - (size_t)acc_y0 * acc_W * BE_ACC; a.blocks = c->blk_dev;
    a.g = g; a.cam = c->cam; a.NB = nblk; a.densify_w = densify_w; a.accH = acc_H; a.accW = acc_W;
    pick_runs(g, nblk, RUN_CTAS, RUN_OVH, &a.G, &a.runs_per_row);
    launch_run(BE_RUN_INFER, a, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}

int be_fold_normalise_band(be_ctx* c, const float* dev_acc, int32_t B, int32_t y0, int32_t rows, int32_t full_H, int32_t acc_W, double thres,
                           float* dev_image, float* dev_sharp, float* dev_refoc, float* dev_bndry, float* dev_depth, float* dev_conf,
                           float* dev_depth_thr, void* stream) {
    if (check_ctx(c)) return 1;
    if (B == 0 || rows == 0) return 0;
    BE_REQUIRE(dev_acc && dev_image && dev_sharp && dev_refoc && dev_bndry && dev_depth && dev_conf, "null pointer");
    BE_REQUIRE(full_H >= c->g.R && acc_W >= c->g.R, "accumulator smaller than a patch");
    BE_REQUIRE(y0 >= 0 && rows > 0 && y0 + rows <= full_H, "row band [%d, %d) outside the %d rows of the image", y0, y0 + rows, full_H);
    BeGeom g = c->g;
    g.H = full_H; g.W = acc_W;
    g.Hp = (full_H - g.R) / g.stride + 1;
    g.Wp = (acc_W - g.R) / g.stride + 1;
    be_launch_normalise(dev_acc, g, B, rows, y0, (float)thres, dev_image, dev_sharp, dev_refoc, dev_bndry, dev_depth, dev_conf, dev_depth_thr,
                        (cudaStream_t)stream);
    BE_CUDA(cudaGetLastError());
    return 0;
}

int be_fold_normalise(be_ctx* c, const float* dev_acc, int32_t B, int32_t acc_H, int32_t acc_W, double thres, float* dev_image,
                      float* dev_sharp, float* dev_refoc, float* dev_bndry, float* dev_depth, float* dev_conf, float* dev_depth_thr,
                      void* stream) {
    return be_fold_normalise_band(c, dev_acc, B, 0, acc_H, acc_H, acc_W, thres, dev_image, dev_sharp, dev_refoc, dev_bndry, dev_depth, dev_conf,
                                  dev_depth_thr, stream);
}

// ---------------------------------------------------------------------------------------------------
// method-granularity entry points (the reference's helper-class METHODS, reference layouts)
// ---------------------------------------------------------------------------------------------------
#define BE_OP_PROLOGUE(...)                 \
    if (check_ctx(c)) return 1;             \
    if (n_ == 0) return 0;                  \
    BE_REQUIRE(__VA_ARGS__, "null pointer"); \
    cudaStream_t st = (cudaStream_t)stream;

int be_params2dists(be_ctx* c, const float* dev_params, int32_t K, int32_t B, int64_t Lsp, float* dev_dists, void* stream) {
    const int64_t n_ = (int64_t)B * Lsp;
    BE_OP_PROLOGUE(dev_params && dev_dists)
    BE_REQUIRE(K >= 8 && Lsp >= 1, "params need >= 8 channels");
    be_op_params2dists(dev_params, K, B, (size_t)Lsp, c->g.R, c->g.w, dev_dists, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_params2dists_bwd(be_ctx* c, const float* dev_params, int32_t K, const float* dev_grad_dists, int32_t B, int64_t Lsp,
                        float* dev_grad_params, void* stream) {
    const int64_t n_ = (int64_t)B * Lsp;
    BE_OP_PROLOGUE(dev_params && dev_grad_dists && dev_grad_params)
    be_op_params2dists_bwd(dev_params, K, dev_grad_dists, B, (size_t)Lsp, c->g.R, c->g.w, dev_grad_params, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_dists2indicators(be_ctx* c, const float* dev_dists, const float* dev_etas, int32_t B, int64_t Lsp, float* dev_wedges, void* stream) {
    const int64_t n_ = (int64_t)B * Lsp;
    BE_OP_PROLOGUE(dev_dists && dev_etas && dev_wedges)
    be_op_indicators(dev_dists, dev_etas, B, (size_t)Lsp, c->g.R * c->g.R, dev_wedges, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_dists2indicators_bwd(be_ctx* c, const float* dev_dists, const float* dev_etas, const float* dev_grad_wedges, int32_t B, int64_t Lsp,
                            float* dev_grad_dists, float* dev_grad_etas, void* stream) {
    const int64_t n_ = (int64_t)B * Lsp;
    BE_OP_PROLOGUE(dev_dists && dev_etas && dev_grad_wedges && dev_grad_dists && dev_grad_etas)
    be_op_indicators_bwd(dev_dists, dev_etas, dev_grad_wedges, B, (size_t)Lsp, c->g.R * c->g.R, dev_grad_dists, dev_grad_etas, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_elementwise(be_ctx* c, int32_t op, const float* dev_x, double p0, int64_t n_, float* dev_y, void* stream) {
    BE_OP_PROLOGUE(dev_x && dev_y)
    BE_REQUIRE(op >= 0 && op <= 2, "unknown elementwise op %d", op);
    be_op_unary(op, dev_x, (float)p0, c->cam, (size_t)n_, dev_y, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_elementwise_bwd(be_ctx* c, int32_t op, const float* dev_x, const float* dev_grad_y, double p0, int64_t n_, float* dev_grad_x,
                       void* stream) {
    BE_OP_PROLOGUE(dev_x && dev_grad_y && dev_grad_x)
    BE_REQUIRE(op >= 0 && op <= 2, "unknown elementwise op %d", op);
    be_op_unary_bwd(op, dev_x, dev_grad_y, (float)p0, c->cam, (size_t)n_, dev_grad_x, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_smish(be_ctx* c, const float* dev_x, int64_t n_, float* dev_y, void* stream) {
    BE_OP_PROLOGUE(dev_x && dev_y)
    be_op_smish(dev_x, (size_t)n_, dev_y, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_smish_bwd(be_ctx* c, const float* dev_x, const float* dev_grad_y, int64_t n_, float* dev_grad_x, void* stream) {
    BE_OP_PROLOGUE(dev_x && dev_grad_y && dev_grad_x)
    be_op_smish_bwd(dev_x, dev_grad_y, (size_t)n_, dev_grad_x, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_etas2depth(be_ctx* c, const float* dev_eta1, const float* dev_eta2, int64_t n_, float* dev_z, void* stream) {
    BE_OP_PROLOGUE(dev_eta1 && dev_eta2 && dev_z)
    be_op_depth(dev_eta1, dev_eta2, c->cam, (size_t)n_, dev_z, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_etas2depth_bwd(be_ctx* c, const float* dev_eta1, const float* dev_eta2, const float* dev_grad_z, int64_t n_, float* dev_g1,
                      float* dev_g2, void* stream) {
    BE_OP_PROLOGUE(dev_eta1 && dev_eta2 && dev_grad_z && dev_g1 && dev_g2)
    be_op_depth_bwd(dev_eta1, dev_eta2, dev_grad_z, c->cam, (size_t)n_, dev_g1, dev_g2, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_inverse_3by3(be_ctx* c, const float* dev_A, int64_t n_, float* dev_inv, void* stream) {
    BE_OP_PROLOGUE(dev_A && dev_inv)
    be_op_inverse3(dev_A, (size_t)n_, dev_inv, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_inverse_3by3_bwd(be_ctx* c, const float* dev_inv, const float* dev_grad_inv, int64_t n_, float* dev_grad_A, void* stream) {
    BE_OP_PROLOGUE(dev_inv && dev_grad_inv && dev_grad_A)
    be_op_inverse3_bwd(dev_inv, dev_grad_inv, (size_t)n_, dev_grad_A, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_image_derivative(be_ctx* c, const float* dev_img, int64_t n_, int32_t H, int32_t W, float* dev_out, void* stream) {
    BE_OP_PROLOGUE(dev_img && dev_out)
    BE_REQUIRE(H >= 3 && W >= 3, "image smaller than the 3x3 Sobel kernel");
    be_op_sobel(dev_img, (size_t)n_, H, W, dev_out, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_image_derivative_bwd(be_ctx* c, const float* dev_img, const float* dev_grad_out, int64_t n_, int32_t H, int32_t W,
                            float* dev_grad_img, void* stream) {
    BE_OP_PROLOGUE(dev_img && dev_grad_out && dev_grad_img)
    be_op_sobel_bwd(dev_img, dev_grad_out, (size_t)n_, H, W, dev_grad_img, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_fold(be_ctx* c, const float* dev_patches, int64_t n_, int32_t mode, float* dev_out, void* stream) {
    BE_OP_PROLOGUE(dev_patches && dev_out)
    be_op_fold(dev_patches, (size_t)n_, c->g, mode, dev_out, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_fold_depth(be_ctx* c, const float* dev_depth_map, const int32_t* dev_depth_mask, int64_t n_, float* dev_depth, float* dev_conf,
                  void* stream) {
    BE_OP_PROLOGUE(dev_depth_map && dev_depth_mask && dev_depth && dev_conf)
    be_op_fold_depth(dev_depth_map, dev_depth_mask, (size_t)n_, c->g, dev_depth, dev_conf, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_unfold(be_ctx* c, const float* dev_img, int64_t n_, int32_t mode, float* dev_patches, void* stream) {
    BE_OP_PROLOGUE(dev_img && dev_patches)
    be_op_unfold(dev_img, (size_t)n_, c->g, mode, dev_patches, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}

int be_patch_gather(be_ctx* c, const float* dev_img, int64_t n_, float* dev_vec, void* stream) {
    BE_OP_PROLOGUE(dev_img && dev_vec)
    be_op_patch_gather(dev_img, (size_t)n_, c->g, dev_vec, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_assemble_pm(be_ctx* c, const float* dev_params, const float* dev_colors, int64_t n_, float* dev_pm, void* stream) {
    BE_OP_PROLOGUE(dev_params && dev_colors && dev_pm)
    be_op_assemble_pm(dev_params, dev_colors, (size_t)n_, (size_t)c->g.Hp * c->g.Wp, dev_pm, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}
int be_eval_depth(be_ctx* c, const float* dev_depth, const float* dev_gt, int64_t n_, int32_t H, int32_t W, int32_t crop, double* dev_sums6,
                  void* stream) {
    BE_OP_PROLOGUE(dev_depth && dev_gt && dev_sums6)
    BE_REQUIRE(crop >= 0 && 2 * crop < H && 2 * crop < W, "crop %d too large for %dx%d", crop, H, W);
    BE_CUDA(cudaMemsetAsync(dev_sums6, 0, (size_t)n_ * 6 * sizeof(double), st));
    be_op_eval_depth(dev_depth, dev_gt, (size_t)n_, H, W, crop, dev_sums6, st);
    BE_CUDA(cudaGetLastError());
    return 0;
}

int be_host_render_fold(be_ctx* c, const float* est, int32_t param_mode, const float* img, const be_image_layout* layout,
                        int32_t B, int32_t densify_w, float* image, float* sharp, float* refoc, float* bndry, float* depth,
                        float* conf, float* depth_thr) {
    if (check_ctx(c) || check_layout(layout)) return 1;
    if (B == 0) return 0;
    BE_REQUIRE(est && img && image && sharp && refoc && bndry && depth && conf, "null pointer");
    BE_REQUIRE(B > 0 && B <= c->cfg.max_batch, "B=%d exceeds max_batch=%d", B, c->cfg.max_batch);
    BE_REQUIRE(param_mode == BE_PARAMS_RESTORED12 || param_mode == BE_PARAMS_RAW12, "be_host_render_fold takes 12-parameter patches");
    BE_REQUIRE(layout->sb == 6LL * c->g.H * c->g.W, "host entry point needs pairs stored contiguously (layout.sb = 6*H*W)");
    const BeGeom& g = c->g;
    const size_t HW = (size_t)g.H * g.W, L = (size_t)g.Hp * g.Wp;
    const size_t mb = (size_t)c->cfg.max_batch;
    if (!c->st_est) {
        if (!c->st_streams[0])
            for (int i = 0; i < 3; ++i) BE_CUDA(cudaStreamCreateWithFlags(&c->st_streams[i], cudaStreamNonBlocking));
        if (!c->st_events[0])
            for (int i = 0; i < 2 * BE_HOST_CHUNKS; ++i) BE_CUDA(cudaEventCreateWithFlags(&c->st_events[i], cudaEventDisableTiming));
        BE_CUDA(cudaMalloc(&c->st_est, mb * L * 12 * sizeof(float)));
        BE_CUDA(cudaMalloc(&c->st_img, mb * 6 * HW * sizeof(float)));
        BE_CUDA(cudaMalloc(&c->st_out, mb * 16 * HW * sizeof(float)));
        c->st_bytes += mb * (L * 12 + 22 * HW) * sizeof(float);
    }
    // Software pipeline over chunks of pairs: H2D(chunk i+1) | kernels(chunk i) | D2H(chunk i-1) on three streams, so that the
    // PCIe transfers hide behind the renderer.  Measured on the B200 box (tools/microbench/pcie.py, BE_HOST_TRACE=1): the call is
    // bound by the link, not by the kernels - D2H runs at 50 GB/s alone but at 21-32 GB/s while H2D is active.  Alternatives tried
    // and rejected: two alternating kernel streams (no gain), an export kernel writing the maps straight into the pinned host
    // arrays (52 GB/s alone, tools/microbench/zerocopy.cu, but 20 % slower in the pipeline than the copy engine).
    if (ensure_acc(c)) return 1;
    cudaStream_t s_in = c->st_streams[0], s_out = c->st_streams[2];
    // Chunk sizes ramp up and down again (1:2:3:4:3:2:1): a small first chunk starts the kernels early, a small last chunk
    // leaves little D2H after the last kernel, the large middle chunks run the renderer at full-wave efficiency.
    static const int ramp_on = [] { const char* e = getenv("BE_HOST_RAMP"); return e ? atoi(e) : 1; }();
    static const int want = [] { const char* e = getenv("BE_HOST_CHUNKS"); const int v = e ? atoi(e) : 8; return v < 1 ? 1 : (v > BE_HOST_CHUNKS ? BE_HOST_CHUNKS : v); }();
    int bounds[BE_HOST_CHUNKS + 1];
    int nchunk = B < want ? B : want;
    if (ramp_on && B >= 16) {
        static int wgt[BE_HOST_CHUNKS] = {1, 2, 3, 4, 3, 2, 1};
        static int nw = 7;
        static const bool parsed = [] {             // BE_HOST_WGT="1,2,3,..." overrides the ramp (tuning aid)
            const char* e = getenv("BE_HOST_WGT");
            if (e) {
                int k = 0;
                while (*e && k < BE_HOST_CHUNKS) { const int v = atoi(e); wgt[k++] = v < 1 ? 1 : v; while (*e && *e != ',') ++e; if (*e == ',') ++e; }
                if (k > 0) nw = k;
            }
            return true;
        }();
        (void)parsed;
        int tot_w = 0;
        for (int i = 0; i < nw; ++i) tot_w += wgt[i];
        nchunk = nw;
        int acc_w = 0;
        bounds[0] = 0;
        for (int i = 0; i < nw; ++i) { acc_w += wgt[i]; bounds[i + 1] = (int)((long long)B * acc_w / tot_w); }
    } else {
        for (int i = 0; i <= nchunk; ++i) bounds[i] = (int)((long long)B * i / nchunk);
    }
    const size_t f = sizeof(float);
    float* o = c->st_out;
    float* d_map[7] = {o, o + (size_t)B * 6 * HW, o + (size_t)B * 9 * HW, o + (size_t)B * 12 * HW, o + (size_t)B * 13 * HW,
                       o + (size_t)B * 14 * HW, o + (size_t)B * 15 * HW};
    float* h_map[7] = {image, sharp, refoc, bndry, depth, conf, depth_thr};
    const size_t per[7] = {6 * HW, 3 * HW, 3 * HW, HW, HW, HW, HW};
    static const bool trace = getenv("BE_HOST_TRACE") != nullptr;
    static cudaEvent_t tev[3 * BE_HOST_CHUNKS + 1];
    if (trace && !tev[0]) for (auto& e : tev) cudaEventCreate(&e);
    if (trace) cudaEventRecord(tev[3 * BE_HOST_CHUNKS], s_in);
    for (int i = 0; i < nchunk; ++i) {
        const int b0 = bounds[i], b1 = bounds[i + 1], nb = b1 - b0;
        if (nb == 0) continue;                      // more chunks than pairs: nothing to copy or launch
        cudaStream_t s_k = c->st_streams[1];
        BE_CUDA(cudaMemcpyAsync(c->st_est + (size_t)b0 * L * 12, est + (size_t)b0 * L * 12, (size_t)nb * L * 12 * f, cudaMemcpyHostToDevice, s_in));
        BE_CUDA(cudaMemcpyAsync(c->st_img + (size_t)b0 * 6 * HW, img + (size_t)b0 * 6 * HW, (size_t)nb * 6 * HW * f, cudaMemcpyHostToDevice, s_in));
        BE_CUDA(cudaEventRecord(c->st_events[2 * i], s_in));
        if (trace) cudaEventRecord(tev[3 * i], s_in);
        BE_CUDA(cudaStreamWaitEvent(s_k, c->st_events[2 * i], 0));
        if (render_fold_range(c, c->st_est + (size_t)b0 * L * 12, param_mode, c->st_img + (size_t)b0 * 6 * HW, layout, b0, nb, densify_w,
                              d_map[0] + b0 * per[0], d_map[1] + b0 * per[1], d_map[2] + b0 * per[2], d_map[3] + b0 * per[3],
                              d_map[4] + b0 * per[4], d_map[5] + b0 * per[5], d_map[6] + b0 * per[6], s_k, false))
            return 1;
        BE_CUDA(cudaEventRecord(c->st_events[2 * i + 1], s_k));
        if (trace) cudaEventRecord(tev[3 * i + 1], s_k);
        BE_CUDA(cudaStreamWaitEvent(s_out, c->st_events[2 * i + 1], 0));
        for (int m = 0; m < 7; ++m) {
            if (!h_map[m]) continue;
            BE_CUDA(cudaMemcpyAsync(h_map[m] + b0 * per[m], d_map[m] + b0 * per[m], nb * per[m] * f, cudaMemcpyDeviceToHost, s_out));
        }
        if (trace) cudaEventRecord(tev[3 * i + 2], s_out);
    }
    BE_CUDA(cudaGetLastError());
    BE_CUDA(cudaStreamSynchronize(s_out));
    if (trace) {
        for (int i = 0; i < nchunk; ++i) {
            float a, b2, d;
            cudaEventElapsedTime(&a, tev[3 * BE_HOST_CHUNKS], tev[3 * i]); cudaEventElapsedTime(&b2, tev[3 * BE_HOST_CHUNKS], tev[3 * i + 1]); cudaEventElapsedTime(&d, tev[3 * BE_HOST_CHUNKS], tev[3 * i + 2]);
            fprintf(stderr, "chunk %d: h2d done %.3f  kernels done %.3f  d2h done %.3f\n", i, a, b2, d);
        }
    }
    return 0;
}

}  // extern "C"

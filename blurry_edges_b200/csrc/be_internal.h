// Internal declarations shared by be_kernels.cu (device code + launchers) and be_capi.cu (C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "be_math.cuh"

constexpr int BE_MAX_R = 21;          // 2 slots/thread x 224 threads cover R*R <= 448 pixels
constexpr int BE_THREADS = 224;       // 7 warps
constexpr int BE_WARPS = BE_THREADS / 32;
constexpr int BE_REC = 32;            // floats per patch-table record (128 B)
constexpr int BE_ACC = 16;            // floats per pixel of the fold accumulator (15 used)
constexpr int BE_GREC = 12;           // floats per patch of the backward chain-rule record
constexpr int BE_CREC = 16;           // floats per patch of the colour record of the TRAINFWD pass: C[9], M^-1[6] (packed 00,01,02,11,12,22)
constexpr int BE_TW = 33;             // floats per pixel of the packed training targets: 8 float4 planes + 1 scalar plane (be_train.cu)

enum BeRunMode { BE_RUN_COLORS = 0, BE_RUN_INFER = 1, BE_RUN_TRAINFWD = 2 };

struct BeGeom {
    int R, stride, H, W, Hp, Wp;
    float w, lam;
};

struct BeImg {                        // strided view of an image tensor (element strides)
    const float* p;
    long long sb, sm, sc, sy, sx;
};

struct BeBlock {                      // one work item of a blocked (big-image) launch: a 147x147 block of a larger image
    int img;                          // image / pair index used for pixel loads and for the accumulator
    int oy, ox;                       // pixel origin of the block inside the larger image
    int py0, py1, px0, px1;           // patch window of the block that is rendered (blurry_edges_test_big.py:166-177)
    int pad;
};

struct BeRunArgs {
    const float* table;               // [NB*L][BE_REC] patch records (be_setup_kernel)
    BeImg img;
    float* acc;                       // [NB][H][W][BE_ACC] fold accumulator (INFER)
    float* colors;                    // [NB][3][3][Hp][Wp]               (COLORS)
    const float* zgt;                 // [NB][H][W] ground-truth boundary depth (TRAINFWD)
    unsigned long long* mask_count;   // sum over the batch of [z_gt != 0 and mask != 0] (TRAINFWD)
    BeGeom g;
    BeCam cam;
    int NB;                           // pairs (INFER) or single images (COLORS)
    int G, runs_per_row;              // patches per CTA run, runs per patch row
    int densify_w;
    float* crec;                      // [NB*L][BE_CREC] colours + inverse normal matrix per patch (TRAINFWD, optional)
    const BeBlock* blocks;            // nullptr: item b is pair/image b at origin (0,0), all patches
    int accH, accW;                   // accumulator plane size (= H, W unless blocked)
    float* stage;                     // deterministic fold: [NB][Hp][R][W][ACCW] per-patch-row slabs written with plain stores instead of the
                                      // atomics on `acc` (needs runs_per_row == 1 and no blocks); nullptr = atomic fold
};

struct BeLossArgs {
    const float* table;               // [N][BE_REC]
    const float* gtable;              // [N][BE_GREC]
    const float* crec;                // [N][BE_CREC] (global loss: written by the TRAINFWD pass)
    const float* T;                   // packed targets of the global loss: 8 float4 planes + 1 scalar plane over [NB][H][W] (be_train.cu)
    int same_gt;                      // img_gt is img_ny (the training call, global_training.py:210): the loss kernel never touches the GT planes
    const float *l_ny, *l_gt, *l_bd, *l_deri;   // local loss: [NB,R,R,3], [NB,R,R,3], [NB,R,R], [NB,R-2,R-2,3]
    const float* l_est;               // local loss: raw LocalStage output [NB,10]; the kernel builds the patch record itself (table may be null)
    float* grad;                      // [N][12|10] or nullptr
    float* grad_depth;                // [N][4] or nullptr: the depth term's share of the eta gradients, NOT divided by the mask count
                                      // (deferred normaliser: the count of the whole batch is still being all-reduced)
    int defer_depth;                  // 1: mask_count is not read by the loss kernel (it may be NULL / not final yet)
    float* partials;                  // [grid][8]
    const unsigned long long* mask_count;
    BeGeom g;
    int NB, G, runs_per_row;
    int b0, NBT;                      // global loss: the launch covers pairs [b0, b0 + NB) of targets T laid out for NBT pairs (table, gtable,
                                      // crec, grad, grad_depth, partials are already offset to pair b0 by the caller)
    float kc, kcc, kbc, ks, ksc, kbl, gamma_d;   // gamma_k / (normaliser_k * N_patches); depth: gamma_d / mask_count
};

struct BeLossScale {                  // partial sums -> reported terms
    int nterms;
    int src[7];
    double scale[7];
    float gamma[7];
    int masked[7];
};

struct BeLocalTail {                  // single-launch local loss: in-kernel setup + last-CTA reduction (be_train.cu)
    BeCam cam;
    BeLossScale sc;
    float* est_wrapped;               // nullable: the angles of est are written back wrapped to [0, 2 pi) (local_training.py:33)
    unsigned* ticket;                 // device counter, zero between launches
    float* terms;
    float* loss;
};

// launchers (be_kernels.cu / be_train.cu); all asynchronous on `st`
void be_launch_setup(const float* est, int param_mode, int npatch, const BeCam& cam, float* table, float* gtable, cudaStream_t st);
// pairs [b0, b0 + nb) of whole-batch arrays laid out for Btot pairs
void be_launch_train_targets(const float* acc, const BeGeom& g, int b0, int nb, int Btot, const float* img_ny, const float* img_gt,
                             const float* bndry_dist, const float* deri, const float* bndry_depth, float* T, float* gimg, float* gbnd,
                             cudaStream_t st);
void be_launch_loss(const BeLossArgs& a, const BeLocalTail& tail, cudaStream_t st);   // local-stage loss, one launch (be_train.cu)
void be_launch_loss2(const BeLossArgs& a, cudaStream_t st);              // global-stage loss (needs a.crec)
void be_launch_grad_depth_fixup(float* grad, const float* grad_depth, const unsigned long long* mask_count, const unsigned long long* true_patches,
                                double assumed_patches, size_t npatch, cudaStream_t st);
void be_launch_loss_reduce(const float* partials, int nblocks, const BeLossScale& sc, const unsigned long long* mask_count,
                           const unsigned long long* true_patches, double assumed_patches, float* terms, float* loss, cudaStream_t st);
void be_launch_run3(int mode, const BeRunArgs& a, cudaStream_t st);   // renderer + fused fold (be_run3.cu)
void be_launch_normalise(const float* acc, const BeGeom& g, int B, int rows, int y_first, float thres, float* image, float* sharp, float* refoc,
                         float* bndry, float* depth, float* conf, float* depth_thr, cudaStream_t st);
void be_launch_stage_reduce(const float* stage, const BeGeom& g, int B, int accw, float* acc, cudaStream_t st);   // fixed-order fold of the slabs
void be_launch_refold(const float* unfolded, const BeGeom& g, int M, float* image, cudaStream_t st);
void be_launch_cover_count(const BeGeom& g, float* out, cudaStream_t st);

extern long long g_be_launches;

// Opt-in to > 48 KB of dynamic shared memory.  The attribute belongs to the (function, device) pair, so it is remembered per
// device: a process that drives several GPUs (one context each) configures every one of them on its first launch there.
constexpr int BE_MAX_DEVICES = 64;
template <typename Kernel>
inline void be_opt_in_smem(Kernel kernel, size_t bytes, bool (&done)[BE_MAX_DEVICES]) {
    int dev = 0;
    cudaGetDevice(&dev);
    const bool known = dev >= 0 && dev < BE_MAX_DEVICES;
    if (known && done[dev]) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (known) done[dev] = true;
}

// method-granularity ops (be_ops.cu)
void be_op_params2dists(const float* params, int K, int B, size_t Lsp, int R, float w, float* dists, cudaStream_t st);
void be_op_params2dists_bwd(const float* params, int K, const float* gd, int B, size_t Lsp, int R, float w, float* gp, cudaStream_t st);
void be_op_indicators(const float* dists, const float* etas, int B, size_t Lsp, int RR, float* wedges, cudaStream_t st);
void be_op_indicators_bwd(const float* dists, const float* etas, const float* gw, int B, size_t Lsp, int RR, float* gd, float* ge,
                          cudaStream_t st);
void be_op_unary(int op, const float* x, float p0, const BeCam& cam, size_t n, float* y, cudaStream_t st);
void be_op_unary_bwd(int op, const float* x, const float* gy, float p0, const BeCam& cam, size_t n, float* gx, cudaStream_t st);
void be_op_smish(const float* x, size_t n, float* y, cudaStream_t st);
void be_op_smish_bwd(const float* x, const float* gy, size_t n, float* gx, cudaStream_t st);
void be_op_depth(const float* e1, const float* e2, const BeCam& cam, size_t n, float* z, cudaStream_t st);
void be_op_depth_bwd(const float* e1, const float* e2, const float* gz, const BeCam& cam, size_t n, float* g1, float* g2, cudaStream_t st);
void be_op_inverse3(const float* A, size_t n, float* out, cudaStream_t st);
void be_op_inverse3_bwd(const float* inv, const float* g, size_t n, float* gA, cudaStream_t st);
void be_op_sobel(const float* img, size_t N, int H, int W, float* out, cudaStream_t st);
void be_op_sobel_bwd(const float* img, const float* gout, size_t N, int H, int W, float* gimg, cudaStream_t st);
void be_op_fold(const float* patches, size_t P, const BeGeom& g, int mode, float* out, cudaStream_t st);
void be_op_fold_depth(const float* dmap, const int* dmask, size_t B, const BeGeom& g, float* depth, float* conf, cudaStream_t st);
void be_op_unfold(const float* img, size_t P, const BeGeom& g, int mode, float* patches, cudaStream_t st);
void be_op_patch_gather(const float* img, size_t M, const BeGeom& g, float* vec, cudaStream_t st);
void be_op_assemble_pm(const float* params, const float* colors, size_t B, size_t L, float* pm, cudaStream_t st);
void be_op_eval_depth(const float* pred, const float* gt, size_t N, int H, int W, int crop, double* out6, cudaStream_t st);

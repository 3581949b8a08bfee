/* blurry_edges_b200 - C ABI of the B200-native (sm_100a) Blurry-Edges render -> fold -> depth path.
 *
 * The reference (guo-research-group/Blurry-Edges) is pure Python and has no FFI; its boundary for
 * this path is the set of Python classes the scripts subclass.  Each entry
 * point below names the reference code it replaces (paths relative to the reference repo).  The
 * Python mirror in blurry_edges_b200/ binds these with ctypes; INTEGRATION.md shows the binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; be_last_error() gives the message
 *   - all tensors are fp32, dense, row-major; "dev" pointers live in HBM of the current device
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls are asynchronous
 *     on it unless their name starts with be_host_
 *   - inputs are borrowed and never written; outputs are caller-allocated
 *   - patch index l = py * Wp + px;  Hp = (H-R)/stride + 1, Wp likewise; R <= 21
 */
#ifndef BLURRY_EDGES_B200_H
#define BLURRY_EDGES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BE_ABI_VERSION 2

/* how a patch's parameter vector is encoded (see be_math.cuh) */
#define BE_PARAMS_RESTORED12 0   /* xy, wrapped angles, eta coefficients: what blurry_edges_test.py:135-138 builds */
#define BE_PARAMS_RAW12 1        /* raw GlobalStage output: restore = global_training.py:141-145 */
#define BE_PARAMS_LOCAL10 2      /* xy, wrapped angles, 2 eta coefficients: blurry_edges_test.py:123-127 */
#define BE_PARAMS_LOCALRAW10 3   /* raw LocalStage output, angles wrapped inside: local_training.py:33 */

typedef struct be_config {
    int32_t R, stride, H, W;          /* utils/args.py:9-11,40  (defaults 21, 2, 147, 147) */
    double w, alpha_lambda;           /* utils/args.py:12-13 (python floats) */
    /* utils/args.py:14-15 camera (DepthEtas.__init__, utils/depth_etas.py:4-21) */
    double cam_s, cam_rho_1, cam_rho_2, cam_sigma_cam, cam_pixel_pitch, cam_mag;
    double rho_prime;                 /* utils/args.py:81 */
    int32_t max_batch;                /* largest number of image PAIRS per call; sizes the workspace */
} be_config;

/* element strides of an image tensor: value(b, m, c, y, x) = base[b*sb + m*sm + c*sc + y*sy + x*sx]
 * (b = pair or image index, m = image within the pair).  Planar [B,2,3,H,W]: {6HW, 3HW, HW, W, 1};
 * dataset-native [B,2,H,W,3] (data/dataset.py:63): {6HW, 3HW, 1, 3W, 3}. */
typedef struct be_image_layout { int64_t sb, sm, sc, sy, sx; } be_image_layout;

typedef struct be_ctx be_ctx;

int be_abi_version(void);
const char* be_last_error(void);

/* PostProcessGlobalBase.__init__ + DepthEtas.__init__ (utils/postprocessing_loss.py:8-20,131-143,
 * utils/depth_etas.py:4-21): validates geometry, derives constants, allocates the HBM workspace. */
int be_ctx_create(be_ctx** out, const be_config* cfg);
int be_ctx_destroy(be_ctx* ctx);
int64_t be_ctx_workspace_bytes(const be_ctx* ctx);
/* derived constants, for the Python mirror's attributes: out[0..7] = numerator, denominator_constant,
 * denominator_factor_root, denominator_factor, intercept, lambda_ridge, Hp, Wp */
int be_ctx_constants(const be_ctx* ctx, double* out8);
/* same constants from a config alone; pure host arithmetic, usable without a GPU */
int be_derive_constants(const be_config* cfg, double* out8);

/* Deterministic fold (torch.use_deterministic_algorithms(True), which global_training.py:177 asks for through
 * utils/util_func.py:17-19).  By default the fused fold adds overlapping patches with fp32 atomics, so the sums of a pixel arrive in a
 * run-dependent order (repeated launches agree to ~1e-7, not bit for bit).  With enable != 0, be_render_fold_fwd,
 * be_host_render_fold and be_global_loss_stage1 / be_host_global_loss* give every patch row ONE thread block, which writes its
 * overlap sums into a private slab with plain stores; a second kernel adds the <= ceil(R/stride) slabs of every pixel in ascending
 * patch-row order.  Results are bit-identical from launch to launch; cost: one extra pass over B*Hp*R*W*{16|8} floats.
 * be_render_fold_blocks refuses to run in this mode. */
int be_ctx_set_deterministic(be_ctx* ctx, int32_t enable);

/* num_patches of PostProcessGlobalBase (utils/postprocessing_loss.py:139-143), closed form. out [H,W]. */
int be_cover_count(be_ctx* ctx, float* dev_out, void* stream);

/* Undo nn.Unfold (blurry_edges_test.py:119-120): unfolded [M,3,R,R,Hp,Wp] -> image [M,3,H,W]. */
int be_refold_image(be_ctx* ctx, const float* dev_unfolded, int32_t M, float* dev_image, void* stream);

/* Pass A: PostProcess.forward(..., colors_only=True) (blurry_edges_test.py:81-92, get_colors :19-28;
 * global_data_pre_cal.py:39-47).  est [M,L,10] (param_mode LOCAL10 or LOCALRAW10), one image per est row
 * (layout.sm unused) -> colors [M,3(channel),3(wedge),Hp,Wp]. */
int be_colors_fwd(be_ctx* ctx, const float* dev_est, int32_t param_mode, const float* dev_img,
                  const be_image_layout* layout, int32_t M, float* dev_colors, void* stream);

/* Pass B: PostProcess.forward(..., colors_only=False) (blurry_edges_test.py:30-100) for B pairs:
 * render both images with shared ridge colours, sharpened + refocused renders, boundary, depth mask/map,
 * and the five folds of utils/postprocessing_loss.py:151-173 - without materialising unfolded tensors.
 * est [B,L,12]; outputs image [B,2,3,H,W], sharp [B,3,H,W], refoc [B,3,H,W], bndry [B,1,H,W],
 * depth [B,H,W], conf [B,H,W].  densify_w != 0 selects the `--densify w` mask rule (:47-50).
 * depth_thresholded (may be NULL): where(conf > thres, depth, 0) with thres = 0 ('w') / 0.05 (:109-112,144). */
int be_render_fold_fwd(be_ctx* ctx, const float* dev_est, int32_t param_mode, const float* dev_img,
                       const be_image_layout* layout, int32_t B, int32_t densify_w,
                       float* dev_image, float* dev_sharp, float* dev_refoc, float* dev_bndry,
                       float* dev_depth, float* dev_conf, float* dev_depth_thresholded, void* stream);

/* Host-buffer form of pass B (what a ctypes binding on the reference side calls with numpy arrays):
 * copies est/img to HBM, runs be_render_fold_fwd, copies the six maps (+thresholded depth) back, and
 * synchronises.  Same shapes as above; img is planar [B,2,3,H,W] or dataset-native via `layout`. */
int be_host_render_fold(be_ctx* ctx, const float* est, int32_t param_mode, const float* img,
                        const be_image_layout* layout, int32_t B, int32_t densify_w,
                        float* image, float* sharp, float* refoc, float* bndry, float* depth, float* conf,
                        float* depth_thresholded);

/* Blocked launches for images larger than the 147x147 network window (blurry_edges_test_big.py:113-190).  A block is a
 * window of the larger image at pixel origin (oy, ox); only the patches [py0,py1) x [px0,px1) of its patch grid are
 * rendered (the 10-patch margins are dropped except at the image border, :166-177).
 * be_colors_blocks_fwd: pass A for nitem (block, image) items; est [nitem,L,10] -> colours [nitem,3,3,Hp,Wp].
 * be_render_fold_blocks: pass B of nblk blocks (est [nblk,L,12]) ADDED into a caller-owned, caller-zeroed accumulator
 *   [*,acc_H,acc_W,16] that holds image rows [acc_y0, acc_y0 + acc_H), at the blocks' origins - the per-block patch grids are never
 *   stitched or unfolded.  Ranks of a multi-GPU job render contiguous bands of blocks into accumulators that cover only the rows
 *   their blocks touch, then exchange row bands (each rank OWNS a band of image rows: it receives the other ranks' partial sums
 *   for those rows, adds them and normalises its band).
 * be_fold_normalise: accumulator [B,acc_H,acc_W,16] -> the six maps (+ thresholded depth) at that size (:185-190).
 * be_fold_normalise_band: the same for rows [y0, y0 + rows) of an image of full_H rows (accumulator and outputs hold the band). */
typedef struct be_block { int32_t img, oy, ox, py0, py1, px0, px1; } be_block;
int be_colors_blocks_fwd(be_ctx* ctx, const float* dev_est, int32_t param_mode, const float* dev_img,
                         const be_image_layout* layout, const be_block* blocks, int32_t nitem, float* dev_colors, void* stream);
int be_render_fold_blocks(be_ctx* ctx, const float* dev_est, int32_t param_mode, const float* dev_img,
                          const be_image_layout* layout, const be_block* blocks, int32_t nblk, int32_t densify_w,
                          int32_t acc_y0, int32_t acc_H, int32_t acc_W, float* dev_acc, void* stream);
int be_fold_normalise(be_ctx* ctx, const float* dev_acc, int32_t B, int32_t acc_H, int32_t acc_W, double thres,
                      float* dev_image, float* dev_sharp, float* dev_refoc, float* dev_bndry, float* dev_depth, float* dev_conf,
                      float* dev_depth_thresholded, void* stream);
int be_fold_normalise_band(be_ctx* ctx, const float* dev_acc, int32_t B, int32_t y0, int32_t rows, int32_t full_H, int32_t acc_W, double thres,
                           float* dev_image, float* dev_sharp, float* dev_refoc, float* dev_bndry, float* dev_depth, float* dev_conf,
                           float* dev_depth_thresholded, void* stream);

/* GlobalLoss.forward + backward (global_training.py:62-157), in two stages so that a data-parallel caller can all-reduce
 * the depth-term normaliser between them (the depth term divides by the mask count of the WHOLE batch, :127).
 * Layouts are the dataset's (data/dataset.py:50-56): raw [B,L,12] network output, img_ny/img_gt [B,2,H,W,3],
 * bndry_dist / bndry_depth [B,H,W], deri [B,2,H-2,W-2,3].
 * stage1: split_restore_params, render both images with shared colours + boundary, fold -> global_image [B,2,3,H,W] and
 *         global_bndry [B,1,H,W] (may be NULL; they are detached targets), mask count -> dev_mask_count (one int64).
 * stage2: the seven loss terms (unweighted, [7]) and loss = sum gamma_k term_k ([1]); if dev_grad != NULL also
 *         d loss / d raw [B,L,12], computed analytically per patch (every folded target is detached in the reference).
 *         global_patches = (global batch) * L; dev_mask_count may have been summed over ranks by the caller.
 * dev_img_gt may be the SAME pointer as dev_img_ny (the training loop of the reference passes the clean image twice,
 * global_training.py:210); stage 2 then runs a kernel variant that never reads the duplicate values (smaller L1 footprint). */
int be_global_loss_stage1(be_ctx* ctx, const float* dev_raw, const float* dev_img_ny, const float* dev_img_gt,
                          const float* dev_bndry_dist, const float* dev_deri, const float* dev_bndry_depth, int32_t B,
                          float* dev_global_image, float* dev_global_bndry, int64_t* dev_mask_count, void* stream);
/* Stage 1 in two calls (same arguments): `render` ends with the kernel that counts the depth mask, `targets` builds the global maps
 * and the packed targets.  A data-parallel caller starts the all-reduce of the count between the two: the collective's kernel then
 * becomes resident while the small target kernels run, instead of queueing behind the loss kernel, which fills every SM. */
int be_global_loss_stage1_render(be_ctx* ctx, const float* dev_raw, const float* dev_img_ny, const float* dev_img_gt,
                                 const float* dev_bndry_dist, const float* dev_deri, const float* dev_bndry_depth, int32_t B,
                                 int64_t* dev_mask_count, void* stream);
int be_global_loss_stage1_targets(be_ctx* ctx, const float* dev_raw, const float* dev_img_ny, const float* dev_img_gt,
                                  const float* dev_bndry_dist, const float* dev_deri, const float* dev_bndry_depth, int32_t B,
                                  float* dev_global_image, float* dev_global_bndry, int64_t* dev_mask_count, void* stream);
int be_global_loss_stage2(be_ctx* ctx, int32_t B, const double* gammas7, int64_t global_patches, const int64_t* dev_mask_count,
                          float* dev_terms, float* dev_loss, float* dev_grad, void* stream);
/* Stage 2 in two calls, for data-parallel callers: `launch` starts the loss kernel without the mask count (the depth term's share
 * of the eta gradients goes, un-normalised, to dev_grad_depth [B,L,4]; dev_grad and dev_grad_depth are both NULL or both set), so
 * the all-reduce of the count over the ranks overlaps the kernel; `finish`, ordered after the all-reduce, produces terms and loss
 * and adds dev_grad_depth / count to dev_grad[:, :, 8:12].  launch + finish == be_global_loss_stage2 up to fp32 rounding of that
 * last addition.  The same gammas7 / global_patches must be passed to both.  Uneven shards: pass global_patches = (local patches) x
 * (ranks) to both and the all-reduced TRUE patch count of the global batch as dev_true_patches (device int64, may be NULL) to
 * `finish`, which rescales terms, loss and gradient by assumed / true - the two counts travel in the same 16-byte all-reduce as
 * the mask count and the loss kernel never waits for it. */
int be_global_loss_stage2_launch(be_ctx* ctx, int32_t B, const double* gammas7, int64_t global_patches, float* dev_grad,
                                 float* dev_grad_depth, void* stream);
int be_global_loss_stage2_finish(be_ctx* ctx, int32_t B, const double* gammas7, int64_t global_patches, const int64_t* dev_mask_count,
                                 const int64_t* dev_true_patches, float* dev_terms, float* dev_loss, float* dev_grad, float* dev_grad_depth,
                                 void* stream);

/* Host-buffer form of the training step (global_training.py:208-211 without the network: criteria(est, ...) + backward to est), what
 * a ctypes binding on the reference side calls with numpy arrays.  All pointers are HOST memory (pinned memory makes the copies
 * asynchronous) in the dataset layouts of be_global_loss_stage1; img_gt may be the same pointer as img_ny.  The batch is cut into
 * chunks of pairs: the H2D copy of chunk i+1 overlaps stage 1 and the loss kernel of chunk i (depth normaliser deferred), then one
 * reduce + depth fix-up and the D2H copy of terms [7], loss [1] and grad [B,L,12] (grad may be NULL).  Synchronous.
 *   be_host_global_loss            one call, one GPU.
 *   be_host_global_loss_begin/_end the two halves for data-parallel callers: `begin` leaves the local mask count in dev_mask_count
 *                                  (DEVICE int64, caller-owned) and issues its kernels on `stream`; the caller all-reduces the count
 *                                  over the ranks on that stream.  global_patches = (local patches) x (ranks) in both halves;
 *                                  dev_true_patches (DEVICE int64 or NULL): the all-reduced true patch count, for uneven shards
 *                                  (same correction as be_global_loss_stage2_finish). */
int be_host_global_loss(be_ctx* ctx, const float* raw, const float* img_ny, const float* img_gt, const float* bndry_dist,
                        const float* deri, const float* bndry_depth, int32_t B, const double* gammas7, float* terms7, float* loss1,
                        float* grad);
int be_host_global_loss_begin(be_ctx* ctx, const float* raw, const float* img_ny, const float* img_gt, const float* bndry_dist,
                              const float* deri, const float* bndry_depth, int32_t B, const double* gammas7, int64_t global_patches,
                              int32_t want_grad, int64_t* dev_mask_count, void* stream);
int be_host_global_loss_end(be_ctx* ctx, int32_t B, const double* gammas7, int64_t global_patches, const int64_t* dev_mask_count,
                            const int64_t* dev_true_patches, float* terms7, float* loss1, float* grad, void* stream);

/* LocalLoss.forward + backward (local_training.py:32-52) in ONE kernel launch: est [B,10] raw LocalStage output, img_ny / img_gt
 * [B,R,R,3], bndry_dist [B,R,R], deri [B,R-2,R-2,3] -> terms [3] = (colour, boundary localisation, smoothness),
 * loss [1] = terms[0] + beta_bndry_loc * terms[1] + beta_smthns * terms[2], grad [B,10] (may be NULL).  Like the reference (:33) the
 * angles est[:, 4:8] are wrapped to [0, 2 pi) IN PLACE when wrap_in_place != 0 - the one input this library writes.
 * Needs a context created with H = W = R. */
int be_local_loss(be_ctx* ctx, float* dev_est, const float* dev_img_ny, const float* dev_img_gt, const float* dev_bndry_dist,
                  const float* dev_deri, int32_t B, double beta_bndry_loc, double beta_smthns, int32_t wrap_in_place, float* dev_terms,
                  float* dev_loss, float* dev_grad, void* stream);

/* Method-granularity entry points: one per METHOD of PostProcessBase / PostProcessGlobalBase / DepthEtas, on the
 * reference's own tensor layouts, each with its backward (autograd of the reference).  Lsp = Hp*Wp for the global layout
 * ([B,K,Hp,Wp], [B,2,R,R,Hp,Wp], ...) and 1 for the local layout ([B,K], [B,2,R,R], ...).  The fused entry points above
 * never call these; they exist so that subclasses written against the reference's base classes run on this library.
 *   be_params2dists        utils/postprocessing_loss.py:43-86   params [B,K>=8,Lsp] (first 8 channels) -> dists [B,2,R,R,Lsp]
 *   be_dists2indicators    :91-95     dists [B,2,R,R,Lsp], etas [B,2,Lsp] -> wedges [B,3,R,R,Lsp]
 *   be_elementwise         op 0 params2etas :88-89, op 1 normalized_gaussian(x, delta=p0) :97-98,
 *                          op 2 DepthEtas.depth2sigma(depth, rho_prime=p0) utils/depth_etas.py:36-37
 *   be_etas2depth          utils/depth_etas.py:23-34
 *   be_inverse_3by3        :104-112   n matrices [n,3,3]
 *   be_image_derivative    :114-117   n planes [n,H,W] -> [n,H-2,W-2]
 *   be_fold                :151-164   n planes of patches [n,R,R,Hp,Wp] -> [n,H,W]; mode 0 divides by num_patches, 1 = plain sum
 *   be_fold_depth          :166-173   depth_map fp32 + depth_mask int32 [n,R,R,Hp,Wp] -> depth, confidence [n,H,W]
 *   be_unfold              nn.Unfold  n planes [n,H,W] -> [n,R,R,Hp,Wp]; mode 0 = adjoint of be_fold mode 0, 1 = plain
 *   be_smish               models/local_stage.py:4-6  Smish activation x tanh(log(1 + sigmoid(x))) of the LocalStage CNN, n elements
 *                          (the elementwise op next to the hot path) */
int be_params2dists(be_ctx* ctx, const float* dev_params, int32_t K, int32_t B, int64_t Lsp, float* dev_dists, void* stream);
int be_params2dists_bwd(be_ctx* ctx, const float* dev_params, int32_t K, const float* dev_grad_dists, int32_t B, int64_t Lsp,
                        float* dev_grad_params, void* stream);
int be_dists2indicators(be_ctx* ctx, const float* dev_dists, const float* dev_etas, int32_t B, int64_t Lsp, float* dev_wedges, void* stream);
int be_dists2indicators_bwd(be_ctx* ctx, const float* dev_dists, const float* dev_etas, const float* dev_grad_wedges, int32_t B,
                            int64_t Lsp, float* dev_grad_dists, float* dev_grad_etas, void* stream);
int be_elementwise(be_ctx* ctx, int32_t op, const float* dev_x, double p0, int64_t n, float* dev_y, void* stream);
int be_elementwise_bwd(be_ctx* ctx, int32_t op, const float* dev_x, const float* dev_grad_y, double p0, int64_t n, float* dev_grad_x,
                       void* stream);
int be_smish(be_ctx* ctx, const float* dev_x, int64_t n, float* dev_y, void* stream);
int be_smish_bwd(be_ctx* ctx, const float* dev_x, const float* dev_grad_y, int64_t n, float* dev_grad_x, void* stream);
int be_etas2depth(be_ctx* ctx, const float* dev_eta1, const float* dev_eta2, int64_t n, float* dev_z, void* stream);
int be_etas2depth_bwd(be_ctx* ctx, const float* dev_eta1, const float* dev_eta2, const float* dev_grad_z, int64_t n, float* dev_g1,
                      float* dev_g2, void* stream);
int be_inverse_3by3(be_ctx* ctx, const float* dev_A, int64_t n, float* dev_inv, void* stream);
int be_inverse_3by3_bwd(be_ctx* ctx, const float* dev_inv, const float* dev_grad_inv, int64_t n, float* dev_grad_A, void* stream);
int be_image_derivative(be_ctx* ctx, const float* dev_img, int64_t n, int32_t H, int32_t W, float* dev_out, void* stream);
int be_image_derivative_bwd(be_ctx* ctx, const float* dev_img, const float* dev_grad_out, int64_t n, int32_t H, int32_t W,
                            float* dev_grad_img, void* stream);
int be_fold(be_ctx* ctx, const float* dev_patches, int64_t n, int32_t mode, float* dev_out, void* stream);
int be_fold_depth(be_ctx* ctx, const float* dev_depth_map, const int32_t* dev_depth_mask, int64_t n, float* dev_depth, float* dev_conf,
                  void* stream);
int be_unfold(be_ctx* ctx, const float* dev_img, int64_t n, int32_t mode, float* dev_patches, void* stream);

/* Glue of the inference driver around passes A and B (blurry_edges_test.py:119-148):
 *   be_patch_gather   n images [n,3,H,W] -> vec [n*L,3,R,R], the LocalStage input (:119-121, Unfold + permute in one pass)
 *   be_assemble_pm    n pairs: raw LocalStage params [2n,L,10] + colours [2n,3,3,Hp,Wp] -> GlobalStage input pm [n,L,38] (:123-132)
 *   be_eval_depth     n images: thresholded depth + ground truth [n,H,W] -> per-image sums [n,6] (fp64): #valid, #acc<1.25,
 *                     #acc<1.25^2, #acc<1.25^3, sum err^2, sum err/gt  (utils/metrics.py:3-21 with msk = depth>0, z in [0.75,1.18]) */
int be_patch_gather(be_ctx* ctx, const float* dev_img, int64_t n, float* dev_vec, void* stream);
int be_assemble_pm(be_ctx* ctx, const float* dev_params, const float* dev_colors, int64_t n, float* dev_pm, void* stream);
int be_eval_depth(be_ctx* ctx, const float* dev_depth, const float* dev_gt, int64_t n, int32_t H, int32_t W, int32_t crop,
                  double* dev_sums6, void* stream);

/* Measurement hook: with timing enabled, be_render_fold_fwd brackets each of its four device operations with CUDA
 * events on the caller's stream; be_ctx_last_timing waits for the last call and returns their durations in ms:
 * ms4 = {accumulator memset, be_setup_kernel, be_run3_kernel, be_normalise_kernel}. */
int be_ctx_set_timing(be_ctx* ctx, int32_t enable);
int be_ctx_last_timing(be_ctx* ctx, float* ms4);
/* The same for the training step (be_global_loss_stage1 + stage2, or stage2_launch + stage2_finish): ms7 = {accumulator memset,
 * be_setup_kernel, be_run3_kernel<TRAINFWD>, ~0 (slot of the former separate normalise launch), be_train_targets_kernel (normalise + target
 * packing in one launch), be_loss2_kernel, reduce (+ depth fix-up)}. */
int be_ctx_last_train_timing(be_ctx* ctx, float* ms7);
/* the same for the step `steps_back` steps before the last one (a ring of 64 event sets: a benchmark loop reads the times of all
 * its steps after the loop, without synchronising the host with the device between steps) */
int be_ctx_train_timing_at(be_ctx* ctx, int32_t steps_back, float* ms7);

/* Number of kernel launches issued by this library since load (for bench.py's gpu_launches). */
int64_t be_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif

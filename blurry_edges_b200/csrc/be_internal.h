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

enum BeRunMode { BE_RUN_COLORS = 0, BE_RUN_INFER = 1 };

struct BeGeom {
    int R, stride, H, W, Hp, Wp;
    float w, lam;
};

struct BeImg {                        // strided view of an image tensor (element strides)
    const float* p;
    long long sb, sm, sc, sy, sx;
};

struct BeRunArgs {
    const float* table;               // [NB*L][BE_REC] patch records (be_setup_kernel)
    BeImg img;
    float* acc;                       // [NB][H][W][BE_ACC] fold accumulator (INFER)
    float* colors;                    // [NB][3][3][Hp][Wp]               (COLORS)
    BeGeom g;
    BeCam cam;
    int NB;                           // pairs (INFER) or single images (COLORS)
    int G, runs_per_row;              // patches per CTA run, runs per patch row
    int densify_w;
};

// launchers (be_kernels.cu); all asynchronous on `st`
void be_launch_setup(const float* est, int param_mode, int npatch, const BeCam& cam, float* table, cudaStream_t st);
void be_launch_run(int mode, const BeRunArgs& a, cudaStream_t st);
void be_launch_normalise(const float* acc, const BeGeom& g, int B, float thres, float* image, float* sharp, float* refoc,
                         float* bndry, float* depth, float* conf, float* depth_thr, cudaStream_t st);
void be_launch_refold(const float* unfolded, const BeGeom& g, int M, float* image, cudaStream_t st);
void be_launch_cover_count(const BeGeom& g, float* out, cudaStream_t st);

extern long long g_be_launches;

// Host launchers for the bandwidth-bound pose kernels (pose_kernels.cu). Return 0 on success; enqueue on `st`.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace stl {

constexpr int kMaxJoints = 64;

int flip_avg(const float* heat, const float* heat_f, float* out, int B, int J, int h, int w, const int* pairs,
             int n_pairs, cudaStream_t st);
int flip_back(const float* in, float* out, int B, int J, int h, int w, const int* pairs, int n_pairs,
              cudaStream_t st);
// heat_f may be null (plain decode). avg_out (may be null) receives the flip-averaged heatmaps when heat_f is given.
// preds/center/scale may be null (heatmap-space decode only); refine=0 gives get_max_preds_hrnet's raw argmax.
int decode(const float* heat, const float* heat_f, const float* center, const float* scale, int B, int J, int h,
           int w, const int* pairs, int n_pairs, int refine, float* avg_out, float* preds, float* maxvals,
           float* coords, cudaStream_t st);
int oks_nms(const float* kpts, const double* area, const double* box_score, const int* offsets, int n_images,
            int max_persons, int J, const double* vars, float in_vis_thr, double oks_thr, float nms_vis_thr, int rescore,
            double* score_out, int* keep_rank, cudaStream_t st);
int generate_target(const double* joints, const double* joints_vis, const float* joints_weight, int B, int J, int h, int w,
                    int image_h, int image_w, int sigma, float* target, float* weight, cudaStream_t st);
int upsampled_argmax(const float* heat, int B, int J, int h, int w, int out_h, int out_w, float* coords, float* maxvals,
                     cudaStream_t st);
int warp_affine_crops(const uint8_t* img, int ih, int iw, const double* minv, int N, int out_h, int out_w,
                      uint8_t* out_u8, float* out_f, const float* mean3, const float* std3, cudaStream_t st);
int warp_affine_crops_f32(const float* img, int ih, int iw, const double* minv, int N, int out_h, int out_w, float* out,
                          cudaStream_t st);
int pck_accuracy(const float* pred, const float* tgt, int B, int J, int h, int w, float thr, float* acc, float* avg_acc,
                 int* cnt, cudaStream_t st);
int scale_inplace(float* x, const float* scale_dev, long long n, cudaStream_t st);
size_t mse_workspace_bytes();
int mse_loss(const float* out, const float* tgt, const float* tw, int B, int J, int hw, float* loss, float* grad,
             void* workspace, cudaStream_t st);

}  // namespace stl

// Host launchers for the auxiliary kernels (aux_kernels.cu). All return 0 on success and enqueue on `st`.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace stl {

int nchw_to_padded(const float* x, __nv_bfloat16* y, int N, int C, int H, int W, int Cp, cudaStream_t st);
int padded_to_nchw(const __nv_bfloat16* y, float* x, int N, int C, int H, int W, int Cp, cudaStream_t st);
int pack_weights(const float* w, const float* gamma, const float* beta, const float* mean, const float* var,
                 const float* cbias, float eps, int Cout, int Cin, int k, int Cout_pad, int Cin_pad,
                 __nv_bfloat16* wp, float* bias_out, cudaStream_t st);
// network input (fp32 NCHW, 3 channels) -> padded-linear NHWC bf16 with 16 channels; images >= n_plain are mirrored
struct PackItem { const float* w; void* wp; int Cout, Cin, k, rows_pad, cols_pad, dgrad; };
int pack_weights_batched(const PackItem* items, const int* block_offsets, int n_items, int total_blocks, cudaStream_t st);
// one torch.optim.SGD step (dampening 0) over every parameter tensor in one launch; buf may be null (no momentum)
struct SgdItem { float* p; const float* g; float* buf; int numel; int pad_; };
int sgd_step_batched(const SgdItem* items, const int* block_offsets, int n_items, int total_blocks, float lr,
                     float momentum, float weight_decay, int nesterov, cudaStream_t st);
int pack_weights_dgrad(const float* w, int Cout, int Cin, int k, int Rows_pad, int K_pad, __nv_bfloat16* wp,
                       float* bias_out, cudaStream_t st);
int add_f32(const float* a, const float* b, float* out, int n, cudaStream_t st);   // out = a + b
int stem_im2col(const float* x, __nv_bfloat16* y, int n_total, int n_plain, int H, int W, cudaStream_t st);
int stem_pack_input(const float* x, __nv_bfloat16* y, int n_total, int n_plain, int H, int W, cudaStream_t st);
int fuse_sum(const __nv_bfloat16* x, const __nv_bfloat16* const* z, const int* shift, int n_up, __nv_bfloat16* y,
             int N, int H, int W, int C, cudaStream_t st);

// y = relu(x + sum up(z)) (C = 32) and the 1x1 heatmap head on it in one pass: heat fp32 [N][J][H][W]; y is not written
int fuse_head(const __nv_bfloat16* x, const __nv_bfloat16* const* z, const int* shift, int n_up, const __nv_bfloat16* w,
              const float* bias, float* heat, int N, int H, int W, int C, int J, cudaStream_t st);

}  // namespace stl

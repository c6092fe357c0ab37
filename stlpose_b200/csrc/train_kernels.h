// Host launchers of the training-path kernels (train_kernels.cu). Return 0 on success; enqueue on `st`.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace stl {

size_t bn_workspace_floats(int C);
int bn_train_forward(const __nv_bfloat16* z, const float* gamma, const float* beta, const __nv_bfloat16* residual,
                     int relu, float eps, float momentum, int N, int H, int W, int C, __nv_bfloat16* y, float* sums,
                     float* mean, float* rstd, float* run_mean, float* run_var, unsigned* ticket, cudaStream_t st);
// statistics + normalisation (backward: both passes) in ONE cooperative launch; `sync`: two device words, zero between
// launches; falls back to the two separate launches on a device without cooperative launch
int bn_train_forward_coop(const __nv_bfloat16* z, const float* gamma, const float* beta, const __nv_bfloat16* residual,
                          int relu, float eps, float momentum, int N, int H, int W, int C, __nv_bfloat16* y, float* sums,
                          float* mean, float* rstd, float* run_mean, float* run_var, unsigned* ticket, unsigned* sync,
                          cudaStream_t st);
int bn_train_backward_coop(const __nv_bfloat16* dy, const __nv_bfloat16* y, const __nv_bfloat16* z, const float* mean,
                           const float* rstd, const float* gamma, const float* beta, int relu, int N, int H, int W, int C,
                           __nv_bfloat16* dz, __nv_bfloat16* dres, float* sums, float* partial, unsigned* ticket,
                           unsigned* sync, cudaStream_t st);
// statistics from the rows a convolution launched with ConvSpec::stats left behind, then the normalisation
int bn_train_forward_fused(const __nv_bfloat16* z, const float* stat_rows, int rows, int c_pad, const float* gamma,
                           const float* beta, const __nv_bfloat16* residual, int relu, float eps, float momentum, int N,
                           int H, int W, int C, __nv_bfloat16* y, float* mean, float* rstd, float* run_mean, float* run_var,
                           cudaStream_t st);
// normalisation alone (mean / rstd already final, e.g. from stl_conv2d_bn)
int bn_apply(const __nv_bfloat16* z, const float* mean, const float* rstd, const float* gamma, const float* beta,
             const __nv_bfloat16* residual, int relu, int N, int H, int W, int C, __nv_bfloat16* y, cudaStream_t st);
// relu: 0 none, 1 mask from the stored output y, 2 mask recomputed from z (needs beta, units without residual)
int bn_train_backward(const __nv_bfloat16* dy, const __nv_bfloat16* y, const __nv_bfloat16* z, const float* mean,
                      const float* rstd, const float* gamma, const float* beta, int relu, int N, int H, int W, int C,
                      __nv_bfloat16* dz, __nv_bfloat16* dres, float* sums, float* partial, unsigned* ticket,
                      cudaStream_t st);
int sum_relu_forward(const __nv_bfloat16* const* same, int n_same, const __nv_bfloat16* const* up, const int* shift,
                     int n_up, __nv_bfloat16* y, int N, int H, int W, int C, cudaStream_t st);
int relu_mask(const __nv_bfloat16* dy, const __nv_bfloat16* y, __nv_bfloat16* g, long long elems, cudaStream_t st);
int upsample_backward(const __nv_bfloat16* g, __nv_bfloat16* dlow, int N, int H, int W, int C, int shift,
                      cudaStream_t st);
int zero_stuff(const __nv_bfloat16* dz, __nv_bfloat16* u, int N, int H, int W, int C, cudaStream_t st);
int conv_dgrad_naive(const __nv_bfloat16* dz, const __nv_bfloat16* w_packed, __nv_bfloat16* dx, int N, int Hi, int Wi,
                     int Cin, int Cout, int k, int stride, cudaStream_t st);
// CUDA-core weight gradient: per-block slabs of partial sums in `workspace`, added in a fixed order (deterministic)
size_t conv_wgrad_naive_workspace_bytes(int N, int Hi, int Wi, int Cin, int Cout, int k, int stride, int cin_real);
int conv_wgrad_naive(const __nv_bfloat16* x, const __nv_bfloat16* dz, float* dw, int N, int Hi, int Wi, int Cin,
                     int Cout, int k, int stride, int cin_real, void* workspace, size_t workspace_bytes,
                     cudaStream_t st);

// tcgen05 weight gradient for stride-1 convolutions (wgrad_tc.cu)
bool wgrad_tc_supported(int W, int cin, int cout, int cin_real, int k, int stride);
size_t wgrad_tc_workspace_bytes(int N, int H, int W, int cin, int cout, int k, int cin_real);
int wgrad_tc_launch(const __nv_bfloat16* x, const __nv_bfloat16* dz, float* dw, int N, int H, int W, int cin, int cout,
                    int k, int cin_real, void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace stl

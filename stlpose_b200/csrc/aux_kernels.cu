// Layout conversion, weight packing (BatchNorm folding), the stem convolution, the fuse-layer sum and a
// CUDA-core reference convolution used to validate the tensor-core kernel on the device.
#include <stdio.h>

#include "aux_kernels.h"
#include "conv.h"

namespace stl {

namespace {

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// ------------------------------------------------------------------ fp32 NCHW -> padded-linear NHWC bf16
__global__ void nchw_to_padded_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int C,
                                      int H, int W, int Cp) {
  const long long total = (long long)N * (H + 1) * (W + 1) * Cp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    long long q = i / Cp;
    const int w = (int)(q % (W + 1));
    q /= (W + 1);
    const int h = (int)(q % (H + 1));
    const int n = (int)(q / (H + 1));
    float v = 0.f;
    if (c < C && h < H && w < W) v = x[(((size_t)n * C + c) * H + h) * W + w];
    y[i] = __float2bfloat16(v);
  }
}

__global__ void padded_to_nchw_kernel(const __nv_bfloat16* __restrict__ y, float* __restrict__ x, int N, int C,
                                      int H, int W, int Cp) {
  const long long total = (long long)N * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    long long t = i / W;
    const int h = (int)(t % H);
    t /= H;
    const int c = (int)(t % C);
    const int n = (int)(t / C);
    x[i] = __bfloat162float(y[(((size_t)n * (H + 1) + h) * (W + 1) + w) * Cp + c]);
  }
}

// ------------------------------------------------------------------ weight packing with BN folding
// w_oihw [Cout][Cin][k][k] fp32  ->  wp [k*k][Cout_pad][Cin_pad] bf16 scaled by gamma/sqrt(var+eps);
// bias_out[c] = beta - mean*scale (+ conv bias * scale).  Padding rows/columns are zero.
__global__ void pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean,
                                    const float* __restrict__ var, const float* __restrict__ cbias, float eps,
                                    int Cout, int Cin, int k, int Cout_pad, int Cin_pad,
                                    __nv_bfloat16* __restrict__ wp, float* __restrict__ bias_out) {
  const int taps = k * k;
  const long long total = (long long)taps * Cout_pad * Cin_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin_pad);
    const int co = (int)((i / Cin_pad) % Cout_pad);
    const int t = (int)(i / ((long long)Cin_pad * Cout_pad));
    float v = 0.f;
    if (co < Cout && ci < Cin) {
      const float scale = gamma ? gamma[co] / sqrtf(var[co] + eps) : 1.f;
      v = w[((size_t)co * Cin + ci) * taps + t] * scale;
    }
    wp[i] = __float2bfloat16(v);
  }
  for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < Cout_pad; co += gridDim.x * blockDim.x) {
    float b = 0.f;
    if (co < Cout) {
      const float scale = gamma ? gamma[co] / sqrtf(var[co] + eps) : 1.f;
      b = (cbias ? cbias[co] * scale : 0.f) + (gamma ? beta[co] - mean[co] * scale : 0.f);
    }
    bias_out[co] = b;
  }
}

// Weights of the convolution that computes a stride-1 dgrad: dx = conv(dz, W') with W'[ci][co][kh][kw] =
// W[co][ci][k-1-kh][k-1-kw].  w_oihw [Cout][Cin][k][k] fp32 -> wp [k*k][Rows_pad][K_pad] bf16, rows = input channels of
// the forward layer, columns = its output channels, taps reversed.  bias_out (Rows_pad) is zeroed.
__global__ void pack_weights_dgrad_kernel(const float* __restrict__ w, int Cout, int Cin, int k, int Rows_pad,
                                          int K_pad, __nv_bfloat16* __restrict__ wp, float* __restrict__ bias_out) {
  const int taps = k * k;
  const long long total = (long long)taps * Rows_pad * K_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % K_pad);
    const int ci = (int)((i / K_pad) % Rows_pad);
    const int t = (int)(i / ((long long)K_pad * Rows_pad));
    float v = 0.f;
    if (co < Cout && ci < Cin) v = w[((size_t)co * Cin + ci) * taps + (taps - 1 - t)];
    wp[i] = __float2bfloat16(v);
  }
  if (bias_out)
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < Rows_pad; r += gridDim.x * blockDim.x) bias_out[r] = 0.f;
}

// ------------------------------------------------------------------ network input -> tensor-core layout
// fp32 NCHW [B][3][H][W] -> padded-linear NHWC bf16 with the 3 channels padded to 16 (one UMMA K-step), so that the
// stem's 3->64 stride-2 conv runs on the same tcgen05 kernel as every other layer.  Images n >= n_plain are image
// n - n_plain mirrored in W: the flip-test pass of lib/inference.py:21 costs no extra input traffic.
constexpr int kStemCin = 16;
__global__ void __launch_bounds__(256) stem_pack_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                        int n_total, int n_plain, int H, int W) {
  const long long total = (long long)n_total * H * W;
  const size_t plane = (size_t)H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const long long t = i / W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    const bool flip = n >= n_plain;
    const float* src = x + (size_t)(flip ? n - n_plain : n) * 3 * plane + (size_t)h * W + (flip ? W - 1 - w : w);
    const float r = __ldcs(src), g = __ldcs(src + plane), b = __ldcs(src + 2 * plane);
    uint4* dst = reinterpret_cast<uint4*>(y + (((size_t)n * (H + 1) + h) * (W + 1) + w) * kStemCin);
    dst[0] = make_uint4(pack_bf16(r, g), pack_bf16(b, 0.f), 0u, 0u);
    dst[1] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// Stem as a GEMM: the first convolution (HRnet.py:286, 3 -> 64 channels, 3x3, stride 2, pad 1) has K = 27.  Packing the
// network input as its im2col matrix - one 32-channel bf16 row per OUTPUT pixel, k = ci*9 + kh*3 + kw (the OIHW order of
// the filter), rows 27..31 zero - makes it a 1x1 convolution for the tensor-core kernel (flat tiles, staged epilogue)
// and halves the bytes the 16-channel-per-input-pixel packing wrote.  y: padded-linear [n_total][H/2+1][W/2+1][32]
// (zero cells untouched); images n >= n_plain are image n - n_plain mirrored in W (flip test, lib/inference.py:21).
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                          int n_total, int n_plain, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)n_total * Ho * Wo;
  const size_t plane = (size_t)H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int wo = (int)(i % Wo);
    const long long t = i / Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const bool flip = n >= n_plain;
    const float* src = x + (size_t)(flip ? n - n_plain : n) * 3 * plane;
    float v[32];
#pragma unroll
    for (int k = 27; k < 32; ++k) v[k] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = 2 * ho + kh - 1;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int w = 2 * wo + kw - 1;
        const bool ok = h >= 0 && h < H && w >= 0 && w < W;
        const size_t off = (size_t)(ok ? h : 0) * W + (ok ? (flip ? W - 1 - w : w) : 0);
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) v[ci * 9 + kh * 3 + kw] = ok ? __ldg(src + ci * plane + off) : 0.f;
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(y + (((size_t)n * (Ho + 1) + ho) * (Wo + 1) + wo) * 32);
#pragma unroll
    for (int g = 0; g < 4; ++g)
      dst[g] = make_uint4(pack_bf16(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16(v[g * 8 + 2], v[g * 8 + 3]),
                          pack_bf16(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16(v[g * 8 + 6], v[g * 8 + 7]));
  }
}

// ------------------------------------------------------------------ fuse-layer sum at the highest resolution
// y = relu(x + sum_u upsample_nearest(z_u))   (HRnet.py:258-264, row i = 0), padded-linear layout, 8 ch / thread
struct FuseArgs {
  const __nv_bfloat16* x;
  const __nv_bfloat16* z[kMaxUp];
  int shift[kMaxUp];
  int n_up;
  __nv_bfloat16* y;
  int N, H, W, C;
  FastDiv fd_c8, fd_wp, fd_hp;   // multiply-shift division of the flat item index (items < 2^31)
};

// NUP = number of upsampled addends (compile time: their gathers are all issued before the first one is consumed)
template <int NUP>
__global__ void __launch_bounds__(256) fuse_sum_kernel(const FuseArgs a) {
  // flat walk over (padded pixel, 8-channel group) items, 4 independent items per thread per iteration so that
  // four 16-byte loads are in flight before the first one is consumed
  const int c8n = a.C / 8;
  const int Wp = a.W + 1, Hp = a.H + 1;
  const uint32_t total = (uint32_t)a.N * (uint32_t)(Hp * Wp * c8n);
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    uint4 u[4];
    uint32_t idx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      idx[k] = i0 + k * stride;
      u[k] = idx[k] < total ? __ldcs(reinterpret_cast<const uint4*>(a.x) + idx[k]) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (idx[k] >= total) continue;
      const uint32_t q = a.fd_c8.div(idx[k]);
      const int c8 = (int)(idx[k] - q * (uint32_t)c8n);
      const uint32_t t = a.fd_wp.div(q);
      const int w = (int)(q - t * (uint32_t)Wp);
      const int n = (int)a.fd_hp.div(t);
      const int h = (int)t - n * Hp;
      uint4 out = make_uint4(0, 0, 0, 0);
      if (h < a.H && w < a.W) {
        float f[8] = {bf16_lo(u[k].x), bf16_hi(u[k].x), bf16_lo(u[k].y), bf16_hi(u[k].y),
                      bf16_lo(u[k].z), bf16_hi(u[k].z), bf16_lo(u[k].w), bf16_hi(u[k].w)};
        uint4 zz[NUP];
#pragma unroll
        for (int j = 0; j < NUP; ++j) {
          const int s = a.shift[j];
          const size_t qs = ((size_t)n * ((a.H >> s) + 1) + (h >> s)) * ((a.W >> s) + 1) + (w >> s);
          zz[j] = __ldg(reinterpret_cast<const uint4*>(a.z[j] + qs * a.C + c8 * 8));
        }
#pragma unroll
        for (int j = 0; j < NUP; ++j) {
          const uint4 z = zz[j];
          f[0] += bf16_lo(z.x); f[1] += bf16_hi(z.x); f[2] += bf16_lo(z.y); f[3] += bf16_hi(z.y);
          f[4] += bf16_lo(z.z); f[5] += bf16_hi(z.z); f[6] += bf16_lo(z.w); f[7] += bf16_hi(z.w);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
        out = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
      }
      __stcs(reinterpret_cast<uint4*>(a.y) + idx[k], out);
    }
  }
}

// The last fuse row of the network and the heatmap head in one pass (HRnet.py:255-264 with n_out = 1, then :466):
//   y = relu(x + sum up(z))  (rounded to bf16 like the stored tensor it replaces)  ->  heat[j] = sum_c W[j][c] y[c] + b[j]
// C = 32 channels, J <= 32 joints.  One thread per padded pixel: 64 B of x, the gathers, 17 x 32 FMAs against the weights
// in shared memory (broadcast reads), fp32 NCHW stores that are coalesced along w.  y is never written: two passes over
// the 32-channel map (209 MB each per 1 024 images at 64x48) and one launch less.
struct FuseHeadArgs {
  const __nv_bfloat16* x;
  const __nv_bfloat16* z[kMaxUp];
  int shift[kMaxUp];
  const __nv_bfloat16* w;      // packed [J_pad][32] bf16
  const float* bias;           // [J_pad]
  float* heat;                 // fp32 [N][J][H][W]
  int N, H, W, J;
  FastDiv fd_wp, fd_hp;
};

template <int NUP>
__global__ void __launch_bounds__(256) fuse_head_kernel(const FuseHeadArgs a) {
  __shared__ __align__(16) float sw[32 * 32];
  __shared__ float sb[32];
  for (int i = threadIdx.x; i < a.J * 32; i += blockDim.x) sw[i] = __bfloat162float(a.w[i]);
  if (threadIdx.x < a.J) sb[threadIdx.x] = a.bias[threadIdx.x];
  __syncthreads();
  const int Wp = a.W + 1, Hp = a.H + 1;
  const uint32_t total = (uint32_t)a.N * (uint32_t)(Hp * Wp);
  const size_t plane = (size_t)a.H * a.W;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
    const uint32_t t = a.fd_wp.div(q);
    const int w = (int)(q - t * (uint32_t)Wp);
    const int n = (int)a.fd_hp.div(t);
    const int h = (int)t - n * Hp;
    if (h >= a.H || w >= a.W) continue;
    const uint4* xp = reinterpret_cast<const uint4*>(a.x) + (size_t)q * 4;
    uint4 u[4], zz[NUP][4];
#pragma unroll
    for (int g = 0; g < 4; ++g) u[g] = __ldcs(xp + g);
#pragma unroll
    for (int j = 0; j < NUP; ++j) {
      const int s = a.shift[j];
      const size_t qs = ((size_t)n * ((a.H >> s) + 1) + (h >> s)) * ((a.W >> s) + 1) + (w >> s);
      const uint4* zp = reinterpret_cast<const uint4*>(a.z[j]) + qs * 4;
#pragma unroll
      for (int g = 0; g < 4; ++g) zz[j][g] = __ldg(zp + g);
    }
    float f[32];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      f[g * 8 + 0] = bf16_lo(u[g].x); f[g * 8 + 1] = bf16_hi(u[g].x); f[g * 8 + 2] = bf16_lo(u[g].y); f[g * 8 + 3] = bf16_hi(u[g].y);
      f[g * 8 + 4] = bf16_lo(u[g].z); f[g * 8 + 5] = bf16_hi(u[g].z); f[g * 8 + 6] = bf16_lo(u[g].w); f[g * 8 + 7] = bf16_hi(u[g].w);
    }
#pragma unroll
    for (int j = 0; j < NUP; ++j) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 z = zz[j][g];
        f[g * 8 + 0] += bf16_lo(z.x); f[g * 8 + 1] += bf16_hi(z.x); f[g * 8 + 2] += bf16_lo(z.y); f[g * 8 + 3] += bf16_hi(z.y);
        f[g * 8 + 4] += bf16_lo(z.z); f[g * 8 + 5] += bf16_hi(z.z); f[g * 8 + 6] += bf16_lo(z.w); f[g * 8 + 7] += bf16_hi(z.w);
      }
    }
    // ReLU, then the bf16 rounding of the tensor this replaces (the head convolution read y from memory)
#pragma unroll
    for (int c = 0; c < 32; ++c) f[c] = __bfloat162float(__float2bfloat16_rn(fmaxf(f[c], 0.f)));
    float* out = a.heat + ((size_t)n * a.J * a.H + h) * a.W + w;
#pragma unroll 1
    for (int j = 0; j < a.J; ++j) {
      const float4* wr = reinterpret_cast<const float4*>(sw + j * 32);
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 ww = wr[c4];
        acc0 = fmaf(f[c4 * 4 + 0], ww.x, acc0); acc1 = fmaf(f[c4 * 4 + 1], ww.y, acc1);
        acc2 = fmaf(f[c4 * 4 + 2], ww.z, acc2); acc3 = fmaf(f[c4 * 4 + 3], ww.w, acc3);
      }
      __stcs(out + (size_t)j * plane, (acc0 + acc1) + (acc2 + acc3) + sb[j]);
    }
  }
}

// Every raw weight repack of a training step in one launch (a step re-packs 293 forward + 292 dgrad layouts; as separate
// launches they are 585 graph nodes of a few microseconds each).  Block -> item by binary search over the block offsets.
__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const PackItem* __restrict__ items,
                                                                   const int* __restrict__ block_offsets, int n_items) {
  int lo = 0, hi = n_items - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (block_offsets[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PackItem it = items[lo];
  const int taps = it.k * it.k;
  const int total = taps * it.rows_pad * it.cols_pad;
  const int base = ((int)blockIdx.x - block_offsets[lo]) * 1024;
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(it.wp);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = base + j * 256 + (int)threadIdx.x;
    if (i >= total) break;
    const int col = i % it.cols_pad;
    const int row = (i / it.cols_pad) % it.rows_pad;
    const int t = i / (it.cols_pad * it.rows_pad);
    const int co = it.dgrad ? col : row, ci = it.dgrad ? row : col;
    float v = 0.f;
    if (co < it.Cout && ci < it.Cin) v = it.w[((size_t)co * it.Cin + ci) * taps + (it.dgrad ? taps - 1 - t : t)];
    wp[i] = __float2bfloat16(v);
  }
}

// ------------------------------------------------------------------ optimizer step of the fine-tuning loop
// torch.optim.SGD.step() (02_train.py:218 with lib/model_setup.py:138-139: momentum 0.9, weight decay 5e-4; dampening 0)
// for ALL parameter tensors in one launch.  torch's foreach implementation needs ~47 multi-tensor launches of ~21 us for
// the 878 tensors of HRNet-W32 (1.0 ms per step, 5 % of a 32-crop step); here block -> tensor by binary search over the
// block offsets, as in pack_weights_batched_kernel.  Same operations in the same order as torch:
//   g = grad (+ weight_decay * p);  buf = momentum * buf + g  (buf starts at zero, so the first step gives buf = g);
//   g = nesterov ? g + momentum * buf : buf;  p -= lr * g
__global__ void __launch_bounds__(256) sgd_batched_kernel(const SgdItem* __restrict__ items,
                                                          const int* __restrict__ block_offsets, int n_items, float lr,
                                                          float momentum, float weight_decay, int nesterov) {
  int lo = 0, hi = n_items - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (block_offsets[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const SgdItem it = items[lo];
  const int base = ((int)blockIdx.x - block_offsets[lo]) * 1024;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = base + j * 256 + (int)threadIdx.x;
    if (i >= it.numel) break;
    const float pv = it.p[i];
    float g = it.g[i];
    if (weight_decay != 0.f) g = __fmaf_rn(weight_decay, pv, g);
    if (it.buf) {
      const float b = __fadd_rn(__fmul_rn(momentum, it.buf[i]), g);   // torch: buf.mul_(momentum).add_(g): two roundings
      it.buf[i] = b;
      g = nesterov ? __fmaf_rn(momentum, b, g) : b;
    }
    it.p[i] = __fmaf_rn(-lr, g, pv);
  }
}

// ------------------------------------------------------------------ CUDA-core reference convolution
struct NaiveArgs {
  const __nv_bfloat16* in;
  const __nv_bfloat16* w;
  const float* bias;
  const __nv_bfloat16* residual;
  const __nv_bfloat16* up_src[kMaxUp];
  int up_shift[kMaxUp];
  int n_up, relu, out_nchw;
  void* out;
  int N, Hi, Wi, Cin, Ho, Wo, cout, cout_pad, k, stride;
};

__global__ void conv_naive_kernel(const NaiveArgs a) {
  const long long total = (long long)a.N * (a.Ho + 1) * (a.Wo + 1) * a.cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % a.cout);
    const long long q = i / a.cout;
    const int wo = (int)(q % (a.Wo + 1));
    const long long t = q / (a.Wo + 1);
    const int ho = (int)(t % (a.Ho + 1));
    const int n = (int)(t / (a.Ho + 1));
    const bool pad = ho == a.Ho || wo == a.Wo;
    float acc = 0.f;
    if (!pad) {
      const int r = a.k / 2;
      for (int kh = 0; kh < a.k; ++kh)
        for (int kw = 0; kw < a.k; ++kw) {
          const int h = ho * a.stride + kh - r, w = wo * a.stride + kw - r;
          if (h < 0 || h >= a.Hi || w < 0 || w >= a.Wi) continue;
          const __nv_bfloat16* ip = a.in + (((size_t)n * (a.Hi + 1) + h) * (a.Wi + 1) + w) * a.Cin;
          const __nv_bfloat16* wp = a.w + ((size_t)(kh * a.k + kw) * a.cout_pad + co) * a.Cin;
          for (int ci = 0; ci < a.Cin; ++ci) acc = fmaf(__bfloat162float(ip[ci]), __bfloat162float(wp[ci]), acc);
        }
      acc += a.bias[co];
      if (a.residual) acc += __bfloat162float(a.residual[(size_t)q * a.cout + co]);
      for (int u = 0; u < a.n_up; ++u) {
        const int s = a.up_shift[u];
        const size_t qs = ((size_t)n * ((a.Ho >> s) + 1) + (ho >> s)) * ((a.Wo >> s) + 1) + (wo >> s);
        acc += __bfloat162float(a.up_src[u][qs * a.cout + co]);
      }
      if (a.relu) acc = fmaxf(acc, 0.f);
    }
    if (a.out_nchw) {
      if (!pad) reinterpret_cast<float*>(a.out)[(((size_t)n * a.cout + co) * a.Ho + ho) * a.Wo + wo] = acc;
    } else {
      reinterpret_cast<__nv_bfloat16*>(a.out)[(size_t)q * a.cout + co] = __float2bfloat16(pad ? 0.f : acc);
    }
  }
}

int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148ll * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace

int nchw_to_padded(const float* x, __nv_bfloat16* y, int N, int C, int H, int W, int Cp, cudaStream_t st) {
  const long long total = (long long)N * (H + 1) * (W + 1) * Cp;
  nchw_to_padded_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, y, N, C, H, W, Cp);
  return check("nchw_to_padded");
}

int padded_to_nchw(const __nv_bfloat16* y, float* x, int N, int C, int H, int W, int Cp, cudaStream_t st) {
  const long long total = (long long)N * C * H * W;
  padded_to_nchw_kernel<<<grid_for(total, 256), 256, 0, st>>>(y, x, N, C, H, W, Cp);
  return check("padded_to_nchw");
}

int pack_weights(const float* w, const float* gamma, const float* beta, const float* mean, const float* var,
                 const float* cbias, float eps, int Cout, int Cin, int k, int Cout_pad, int Cin_pad,
                 __nv_bfloat16* wp, float* bias_out, cudaStream_t st) {
  const long long total = (long long)k * k * Cout_pad * Cin_pad;
  pack_weights_kernel<<<grid_for(total, 256), 256, 0, st>>>(w, gamma, beta, mean, var, cbias, eps, Cout, Cin, k,
                                                            Cout_pad, Cin_pad, wp, bias_out);
  return check("pack_weights");
}

int pack_weights_dgrad(const float* w, int Cout, int Cin, int k, int Rows_pad, int K_pad, __nv_bfloat16* wp,
                       float* bias_out, cudaStream_t st) {
  if (Rows_pad < Cin || K_pad < Cout) { set_error("pack_weights_dgrad: padded sizes smaller than the real ones"); return 1; }
  const long long total = (long long)k * k * Rows_pad * K_pad;
  pack_weights_dgrad_kernel<<<grid_for(total, 256), 256, 0, st>>>(w, Cout, Cin, k, Rows_pad, K_pad, wp, bias_out);
  return check("pack_weights_dgrad");
}

int pack_weights_batched(const PackItem* items, const int* block_offsets, int n_items, int total_blocks, cudaStream_t st) {
  if (n_items <= 0 || total_blocks <= 0) return 0;
  pack_weights_batched_kernel<<<total_blocks, 256, 0, st>>>(items, block_offsets, n_items);
  return check("pack_weights_batched");
}

int sgd_step_batched(const SgdItem* items, const int* block_offsets, int n_items, int total_blocks, float lr,
                     float momentum, float weight_decay, int nesterov, cudaStream_t st) {
  if (n_items <= 0 || total_blocks <= 0) return 0;
  sgd_batched_kernel<<<total_blocks, 256, 0, st>>>(items, block_offsets, n_items, lr, momentum, weight_decay, nesterov);
  return check("sgd_step_batched");
}

int stem_pack_input(const float* x, __nv_bfloat16* y, int n_total, int n_plain, int H, int W, cudaStream_t st) {
  const long long total = (long long)n_total * H * W;
  stem_pack_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, y, n_total, n_plain, H, W);
  return check("stem_pack_input");
}

int stem_im2col(const float* x, __nv_bfloat16* y, int n_total, int n_plain, int H, int W, cudaStream_t st) {
  const long long total = (long long)n_total * (H / 2) * (W / 2);
  stem_im2col_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, y, n_total, n_plain, H, W);
  return check("stem_im2col");
}

namespace {
__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}
}  // namespace

int add_f32(const float* a, const float* b, float* out, int n, cudaStream_t st) {
  if (n <= 0) return 0;
  add_f32_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, b, out, n);
  return check("add_f32");
}

int fuse_sum(const __nv_bfloat16* x, const __nv_bfloat16* const* z, const int* shift, int n_up, __nv_bfloat16* y,
             int N, int H, int W, int C, cudaStream_t st) {
  FuseArgs a{};
  a.x = x; a.y = y; a.n_up = n_up; a.N = N; a.H = H; a.W = W; a.C = C;
  for (int i = 0; i < n_up; ++i) { a.z[i] = z[i]; a.shift[i] = shift[i]; }
  const long long total = (long long)N * (H + 1) * (W + 1) * (C / 8);
  if (total >= (1ll << 31) - 4ll * 148 * 16 * 256) { set_error("fuse_sum: tensor too large for 32-bit item indexing"); return 1; }
  a.fd_c8.init((uint32_t)(C / 8));
  a.fd_wp.init((uint32_t)(W + 1));
  a.fd_hp.init((uint32_t)(H + 1));
  long long blocks = (total + 4 * 256 - 1) / (4 * 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  switch (n_up) {
    case 1: fuse_sum_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(a); break;
    case 2: fuse_sum_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(a); break;
    case 3: fuse_sum_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(a); break;
    default: set_error("fuse_sum: %d upsampled addends (1..3 supported)", n_up); return 1;
  }
  return check("fuse_sum");
}

int fuse_head(const __nv_bfloat16* x, const __nv_bfloat16* const* z, const int* shift, int n_up, const __nv_bfloat16* w,
              const float* bias, float* heat, int N, int H, int W, int C, int J, cudaStream_t st) {
  if (C != 32 || J < 1 || J > 32 || n_up < 1 || n_up > kMaxUp) { set_error("fuse_head: C=%d J=%d n_up=%d unsupported", C, J, n_up); return 1; }
  FuseHeadArgs a{};
  a.x = x; a.w = w; a.bias = bias; a.heat = heat; a.N = N; a.H = H; a.W = W; a.J = J;
  for (int i = 0; i < n_up; ++i) { a.z[i] = z[i]; a.shift[i] = shift[i]; }
  const long long total = (long long)N * (H + 1) * (W + 1);
  if (total >= (1ll << 31) - 148ll * 16 * 256) { set_error("fuse_head: tensor too large for 32-bit pixel indexing"); return 1; }
  a.fd_wp.init((uint32_t)(W + 1));
  a.fd_hp.init((uint32_t)(H + 1));
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  switch (n_up) {
    case 1: fuse_head_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(a); break;
    case 2: fuse_head_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(a); break;
    case 3: fuse_head_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(a); break;
  }
  return check("fuse_head");
}

int conv_launch_naive(const ConvSpec& s, cudaStream_t st) {
  NaiveArgs a{};
  a.in = s.in; a.w = s.weights; a.bias = s.bias; a.residual = s.residual;
  for (int i = 0; i < kMaxUp; ++i) { a.up_src[i] = s.up_src[i]; a.up_shift[i] = s.up_shift[i]; }
  a.n_up = s.n_up; a.relu = s.relu; a.out_nchw = s.out_nchw; a.out = s.out;
  a.N = s.in_geom.N; a.Hi = s.in_geom.H; a.Wi = s.in_geom.W; a.Cin = s.in_geom.C;
  a.Ho = a.Hi / s.stride; a.Wo = a.Wi / s.stride;
  a.cout = s.cout; a.cout_pad = s.cout_pad; a.k = s.ksize; a.stride = s.stride;
  const long long total = (long long)a.N * (a.Ho + 1) * (a.Wo + 1) * a.cout;
  conv_naive_kernel<<<grid_for(total, 128), 128, 0, st>>>(a);
  return check("conv_naive");
}

}  // namespace stl

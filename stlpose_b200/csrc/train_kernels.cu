// Training-path kernels (fine-tuning step of /root/reference/src/02_train.py:203-218): BatchNorm with batch
// statistics (forward and backward), the fuse-layer sum with its backward, and CUDA-core convolution gradients for
// the cases the tensor-core kernel does not cover yet (stride-2 dgrad, every wgrad).
//
// All activations and activation gradients use the engine's padded-linear NHWC bf16 layout (conv.h); zero cells are
// kept zero by every kernel here.  Statistics and parameter gradients are fp32.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "conv.h"
#include "ptx.cuh"
#include "train_kernels.h"

namespace stl {

namespace {

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

int check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return 1; }
  return 0;
}

// Flat 16-byte item index -> (channel group, w, h, n) of the padded layout with multiply-shift division: the 64-bit
// div/mod of the first version cost more than the memory traffic (1.2 TB/s on the 256-channel tensors).
struct PixIdx {
  FastDiv c8, wp, hp;
  int c8n, Wp, Hp;
  __host__ void init(int C, int H, int W) {
    c8n = C / 8; Wp = W + 1; Hp = H + 1;
    c8.init((uint32_t)c8n); wp.init((uint32_t)Wp); hp.init((uint32_t)Hp);
  }
  __device__ __forceinline__ void decode(uint32_t i, int& cg, int& w, int& h, int& n) const {
    const uint32_t q = c8.div(i);
    cg = (int)(i - q * (uint32_t)c8n);
    const uint32_t t = wp.div(q);
    w = (int)(q - t * (uint32_t)Wp);
    const uint32_t nn = hp.div(t);
    h = (int)(t - nn * (uint32_t)Hp);
    n = (int)nn;
  }
};

// gamma * (z - mean) * rstd + beta with a FIXED operation order: the forward (bn_apply_kernel) and the backward kernels
// that recompute the ReLU mask from z instead of reading y must produce the same float, bit for bit.
__device__ __forceinline__ float bn_affine(float z, float g, float m, float r, float b) {
  return __fmaf_rn(__fmul_rn(g, __fsub_rn(z, m)), r, b);
}

// ------------------------------------------------------------------ per-channel reductions over all pixels
// MODE 0: sums[c] = sum a, sums[C+c] = sum a^2                      (BatchNorm forward statistics; a = z)
// MODE 1: sums[c] = sum g, sums[C+c] = sum g * xhat                 (BatchNorm backward; g = dy masked by y > 0)
// Zero cells of the padded layout contribute zero to every sum, so the kernel walks the flat tensor.
// Block = 256 threads = (C/8 channel groups) x (256 / (C/8) pixel lanes); fp32 register partials, shared-memory
// reduction over the pixel lanes, one row of per-block partials in global memory.  The last block to finish (ticket
// counter) adds the rows in a fixed order -- no floating-point atomics, so the statistics are deterministic -- and, in
// MODE 0, also derives mean / rstd and updates the running statistics (momentum, unbiased variance: nn.BatchNorm2d).
constexpr int kReduceBlocks = 296;   // 2 per SM
constexpr int kCoopMaxC = 512;       // cooperative BatchNorm kernels stage 2 x C floats in shared memory

struct BnFinalize {
  float count, eps, momentum;
  float *mean, *rstd, *run_mean, *run_var;
};

template <int MODE>
__device__ __forceinline__ bool channel_reduce_body(const __nv_bfloat16* __restrict__ a,
                                                             const __nv_bfloat16* __restrict__ y,
                                                             const __nv_bfloat16* __restrict__ z,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ rstd,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, long long pixels, int C,
                                                             int relu_mask, float* __restrict__ sums,
                                                             float* __restrict__ partial, unsigned* __restrict__ counter,
                                                             const BnFinalize fin) {
  const int c8n = C / 8;
  const int lanes = 256 / c8n;          // pixel lanes per block (C <= 2048 / 8 ... C/8 <= 256)
  const int cg = threadIdx.x % c8n, pl = threadIdx.x / c8n;
  float s0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float mu[8], rs[8], ga[8], be[8];
  if (MODE == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mu[i] = mean[cg * 8 + i]; rs[i] = rstd[cg * 8 + i];
      ga[i] = relu_mask == 2 ? gamma[cg * 8 + i] : 0.f; be[i] = relu_mask == 2 ? beta[cg * 8 + i] : 0.f;
    }
  }
  if (pl < lanes) {
    // explicit batches: all loads of kBatch pixels are issued before the first one is consumed (the rolled loop kept
    // only two 16-byte loads in flight per thread: 0.7 TB/s on the large tensors, latency-bound on the small ones)
    constexpr int kBatch = MODE == 0 ? 8 : 4;
    const long long pstep = (long long)gridDim.x * lanes;
    for (long long p0 = (long long)blockIdx.x * lanes + pl; p0 < pixels; p0 += kBatch * pstep) {
      uint4 ua[kBatch], uy[MODE == 1 ? kBatch : 1], uz[MODE == 1 ? kBatch : 1];
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const long long p = p0 + k * pstep;
        const bool in = p < pixels;
        const size_t off = (size_t)(in ? p : 0) * C + cg * 8;
        ua[k] = in ? __ldg(reinterpret_cast<const uint4*>(a + off)) : make_uint4(0, 0, 0, 0);
        if (MODE == 1) {
          uz[k] = in ? __ldg(reinterpret_cast<const uint4*>(z + off)) : make_uint4(0, 0, 0, 0);
          uy[k] = (in && relu_mask == 1) ? __ldg(reinterpret_cast<const uint4*>(y + off)) : make_uint4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        if (p0 + k * pstep >= pixels) break;       // (an all-zero item would also be neutral in MODE 0, not in MODE 1)
        float f[8];
        unpack8(ua[k], f);
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { s0[i] += f[i]; s1[i] += f[i] * f[i]; }
        } else {
          float yy[8], zz[8];
          unpack8(uz[k], zz);
          unpack8(uy[k], yy);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // relu_mask 1: the stored output tells; 2 (no residual): y > 0 <=> the affine value is > 0, recomputed from z
            const bool off_ = relu_mask == 1 ? !(yy[i] > 0.f)
                                             : (relu_mask == 2 && !(bn_affine(zz[i], ga[i], mu[i], rs[i], be[i]) > 0.f));
            const float g = off_ ? 0.f : f[i];
            s0[i] += g;
            s1[i] += g * (zz[i] - mu[i]) * rs[i];
          }
        }
      }
    }
  }
  __shared__ __align__(16) float red[2][256][8 + 1];
  __shared__ bool is_last;
  // pixel lanes that share a warp (32 / c8n of them when c8n < 32) are combined by shuffles first, in a fixed order
  int lane_step = c8n;      // after the loop: lanes pl with (pl * c8n) % 32 == 0 hold a warp's sum
  if (c8n < 32 && (32 % c8n) == 0) {
    for (int off = c8n; off < 32; off <<= 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s0[i] += __shfl_down_sync(0xFFFFFFFFu, s0[i], off);
        s1[i] += __shfl_down_sync(0xFFFFFFFFu, s1[i], off);
      }
    }
    lane_step = 32;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[0][threadIdx.x][i] = s0[i]; red[1][threadIdx.x][i] = s1[i]; }
  __syncthreads();
  if (pl == 0) {
    // remaining partial sums: one per warp (lane_step == 32: threads cg, 32 + cg, ...) or one per pixel lane
    const int nthreads = lanes * c8n;
    for (int t = lane_step + cg; t < nthreads; t += lane_step)
#pragma unroll
      for (int i = 0; i < 8; ++i) { s0[i] += red[0][t][i]; s1[i] += red[1][t][i]; }
    float* row = partial + (size_t)blockIdx.x * 2 * C;
#pragma unroll
    for (int i = 0; i < 8; ++i) { row[cg * 8 + i] = s0[i]; row[C + cg * 8 + i] = s1[i]; }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  // fixed-order sum of the rows with 16-byte loads: thread = (row group `part`, 4 consecutive columns); every group
  // walks its interleaved share of the rows with four independent accumulators per column (4 x 16 B in flight), then
  // the groups are combined in order through shared memory.  2C is a multiple of 16.
  {
    const int C2 = 2 * C, q4 = C2 / 4;                          // float4 columns per row
    float* comb = &red[0][0][0];                                // 2 x 256 x 9 floats >= 256 x 4
    for (int c0 = 0; c0 < q4; c0 += 256) {
      const int width = q4 - c0 < 256 ? q4 - c0 : 256;          // float4 columns handled in this pass
      const int parts = 256 / width;                            // row groups (1 when the pass is 256 columns wide)
      const int part = threadIdx.x / width, cc = threadIdx.x % width;
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
      if (part < parts) {
        const float4* src = reinterpret_cast<const float4*>(partial) + c0 + cc;
        unsigned b2 = part;
        for (; b2 + 3 * parts < gridDim.x; b2 += 4 * parts) {
          const float4 v0 = __ldcg(src + (size_t)b2 * q4), v1 = __ldcg(src + (size_t)(b2 + parts) * q4);
          const float4 v2 = __ldcg(src + (size_t)(b2 + 2 * parts) * q4), v3 = __ldcg(src + (size_t)(b2 + 3 * parts) * q4);
          a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
          a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
          a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
          a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
        }
        for (; b2 < gridDim.x; b2 += parts) {
          const float4 v0 = __ldcg(src + (size_t)b2 * q4);
          a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
        }
      }
      float4 acc;
      acc.x = (a0.x + a1.x) + (a2.x + a3.x); acc.y = (a0.y + a1.y) + (a2.y + a3.y);
      acc.z = (a0.z + a1.z) + (a2.z + a3.z); acc.w = (a0.w + a1.w) + (a2.w + a3.w);
      __syncthreads();                                          // (the block-level reduction above is done with `red`)
      if (part < parts) reinterpret_cast<float4*>(comb)[part * width + cc] = acc;
      __syncthreads();
      if (part == 0) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < parts; ++q) {
          const float4 v = reinterpret_cast<const float4*>(comb)[q * width + cc];
          t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        reinterpret_cast<float4*>(sums)[c0 + cc] = t;
      }
    }
  }
  if (threadIdx.x == 0) *counter = 0u;
  if (MODE == 0) {
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      const float m = sums[c] / fin.count;
      float var = sums[C + c] / fin.count - m * m;
      var = var > 0.f ? var : 0.f;
      fin.mean[c] = m;
      fin.rstd[c] = rsqrtf(var + fin.eps);
      if (fin.run_mean) {
        const float unbiased = fin.count > 1.f ? var * fin.count / (fin.count - 1.f) : var;
        fin.run_mean[c] = (1.f - fin.momentum) * fin.run_mean[c] + fin.momentum * m;
        fin.run_var[c] = (1.f - fin.momentum) * fin.run_var[c] + fin.momentum * unbiased;
      }
    }
  }
  return true;
}

template <int MODE>
__global__ void __launch_bounds__(256) channel_reduce_kernel(const __nv_bfloat16* __restrict__ a,
                                                             const __nv_bfloat16* __restrict__ y,
                                                             const __nv_bfloat16* __restrict__ z,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ rstd,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, long long pixels, int C,
                                                             int relu_mask, float* __restrict__ sums,
                                                             float* __restrict__ partial, unsigned* __restrict__ counter,
                                                             const BnFinalize fin) {
  // programmatic dependent launch (no-ops without the launch attribute): the next kernel of the stream may be scheduled
  // as soon as all our blocks have started; we touch memory only once the previous kernel has completed
  pdl_launch_dependents();
  pdl_wait();
  channel_reduce_body<MODE>(a, y, z, mean, rstd, gamma, beta, pixels, C, relu_mask, sums, partial, counter, fin);
}

// Grid-wide hand-over inside ONE cooperative launch (all blocks co-resident): the block that finished the reduction
// (`is_last`) publishes its results and raises `flag`; every other block waits for it.  sync[0] = flag, sync[1] = number
// of blocks that have left; the last one to leave clears both, so the words are zero again between launches.
__device__ __forceinline__ void grid_handover(unsigned* sync, bool is_last) {
  __syncthreads();
  if (threadIdx.x == 0) {
    if (is_last) {
      __threadfence();
      atomicExch(sync, 1u);
    } else {
      while (atomicAdd(sync, 0u) == 0u) __nanosleep(32);
    }
    __threadfence();
  }
  __syncthreads();
}
__device__ __forceinline__ void grid_leave(unsigned* sync) {
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(sync + 1, 1u) == gridDim.x - 1) {
    sync[0] = 0u;
    sync[1] = 0u;
    __threadfence();
  }
}

// BatchNorm statistics from the per-CTA rows the convolution epilogue wrote (conv.h ConvSpec::stats: rows x [2][c_pad]
// floats = sum | sum of squares of the stored conv outputs over the valid pixels): rows added in a fixed order (four
// independent partial sums per channel), then mean / rstd and the running-statistics update exactly like the tail of
// channel_reduce_kernel<0>.
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ rows_, int rows, int C, int c_pad,
                                                          const BnFinalize fin) {
  // block = 32 channels x 8 row groups: group g adds rows g, g + 8, ... with four independent partial sums per statistic
  // (8 loads in flight per thread; a single thread per channel walking all ~148 rows was latency-bound at ~12 us), the
  // groups are then combined in order through shared memory
  __shared__ float part[2][8][32];
  const int cl = threadIdx.x & 31, g = threadIdx.x >> 5, c = blockIdx.x * 32 + cl;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    int r = g;
    for (; r + 24 < rows; r += 32) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s[k] += __ldcg(rows_ + (size_t)(r + 8 * k) * 2 * c_pad + c);
        q[k] += __ldcg(rows_ + (size_t)(r + 8 * k) * 2 * c_pad + c_pad + c);
      }
    }
    for (; r < rows; r += 8) { s[0] += __ldcg(rows_ + (size_t)r * 2 * c_pad + c); q[0] += __ldcg(rows_ + (size_t)r * 2 * c_pad + c_pad + c); }
  }
  part[0][g][cl] = (s[0] + s[1]) + (s[2] + s[3]);
  part[1][g][cl] = (q[0] + q[1]) + (q[2] + q[3]);
  __syncthreads();
  if (g == 0 && c < C) {
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sum += part[0][k][cl]; sq += part[1][k][cl]; }
    const float m = sum / fin.count;
    float var = sq / fin.count - m * m;
    var = var > 0.f ? var : 0.f;
    fin.mean[c] = m;
    fin.rstd[c] = rsqrtf(var + fin.eps);
    if (fin.run_mean) {
      const float unbiased = fin.count > 1.f ? var * fin.count / (fin.count - 1.f) : var;
      fin.run_mean[c] = (1.f - fin.momentum) * fin.run_mean[c] + fin.momentum * m;
      fin.run_var[c] = (1.f - fin.momentum) * fin.run_var[c] + fin.momentum * unbiased;
    }
  }
}

// y = [relu]( gamma * (z - mean) * rstd + beta [+ residual] ) on the valid pixels, zero on the zero cells
template <int NB>
__device__ __forceinline__ void bn_apply_body(const __nv_bfloat16* __restrict__ z,
                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                              const __nv_bfloat16* __restrict__ residual, int relu,
                                              __nv_bfloat16* __restrict__ y, int N, int H, int W, int C, const PixIdx& px) {
  const uint32_t total = (uint32_t)N * (uint32_t)(px.Hp * px.Wp * px.c8n);
  // the grid stride is a multiple of the channel-group count (launcher), so a thread keeps its 8 channels: the
  // per-channel parameters live in registers instead of 32 scalar loads per 16-byte item
  float g8[8], m8[8], r8[8], b8[8];
  {
    const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int cg0 = (int)(i0 - px.c8.div(i0) * (uint32_t)px.c8n);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      g8[k] = gamma[cg0 * 8 + k]; m8[k] = mean[cg0 * 8 + k]; r8[k] = rstd[cg0 * 8 + k]; b8[k] = beta[cg0 * 8 + k];
    }
  }
  // NB items per iteration, all loads issued before the first one is consumed (zero cells are loaded too: they are
  // inside the tensor); the stride keeps a thread on its channel group
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += NB * stride) {
    uint32_t idx[NB];
    uint4 uz[NB], ur[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      idx[j] = i0 + (uint32_t)j * stride;
      const bool in = idx[j] < total;
      uz[j] = in ? *reinterpret_cast<const uint4*>(z + (size_t)idx[j] * 8) : make_uint4(0, 0, 0, 0);
      ur[j] = (in && residual) ? *reinterpret_cast<const uint4*>(residual + (size_t)idx[j] * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const uint32_t i = idx[j];
      if (i >= total) break;
      int cg, w, h, n;
      px.decode(i, cg, w, h, n);
      uint4 out = make_uint4(0, 0, 0, 0);
      if (h < H && w < W) {
        float f[8], r[8];
        unpack8(uz[j], f);
        unpack8(ur[j], r);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float v = bn_affine(f[k], g8[k], m8[k], r8[k], b8[k]) + r[k];
          f[k] = relu ? fmaxf(v, 0.f) : v;
        }
        out = pack8(f);
      }
      *reinterpret_cast<uint4*>(y + (size_t)i * 8) = out;
    }
  }
}

template <int NB>
__global__ void __launch_bounds__(256) bn_apply_kernel(const __nv_bfloat16* __restrict__ z,
                                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const __nv_bfloat16* __restrict__ residual, int relu,
                                                       __nv_bfloat16* __restrict__ y, int N, int H, int W, int C, const PixIdx px) {
  pdl_launch_dependents();
  pdl_wait();
  bn_apply_body<NB>(z, mean, rstd, gamma, beta, residual, relu, y, N, H, W, C, px);
}

// Batch statistics + normalisation of one BatchNorm layer in ONE cooperative launch: phase 1 = channel_reduce_kernel<0>
// (the last block finalises mean / rstd / running statistics), grid hand-over, phase 2 = bn_apply_kernel.  At fine-tuning
// batch sizes the two separate launches are latency, not bandwidth (8 + 5 us on tensors of a few hundred KB).
__global__ void __launch_bounds__(256) bn_forward_coop_kernel(const __nv_bfloat16* __restrict__ z,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              const __nv_bfloat16* __restrict__ residual, int relu,
                                                              __nv_bfloat16* __restrict__ y, long long pixels, int N, int H,
                                                              int W, int C, float* __restrict__ sums,
                                                              float* __restrict__ partial, unsigned* __restrict__ counter,
                                                              unsigned* __restrict__ sync, const BnFinalize fin,
                                                              const PixIdx px) {
  pdl_launch_dependents();   // (a following convolution launched with the attribute may start its prologue)
  const bool last = channel_reduce_body<0>(z, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, pixels, C, 0, sums,
                                           partial, counter, fin);
  grid_handover(sync, last);
  // mean / rstd were written by another block of THIS launch: read them once, past L1, into shared memory
  __shared__ float s_stat[2 * kCoopMaxC];
  for (int c = threadIdx.x; c < C; c += blockDim.x) { s_stat[c] = __ldcg(fin.mean + c); s_stat[kCoopMaxC + c] = __ldcg(fin.rstd + c); }
  __syncthreads();
  bn_apply_body<2>(z, s_stat, s_stat + kCoopMaxC, gamma, beta, residual, relu, y, N, H, W, C, px);
  grid_leave(sync);
}

// dz = gamma * rstd * (g - sum_g/cnt - xhat * sum_gx/cnt),  g = dy masked by (y > 0) when relu; d_res = g
template <int NB>
__device__ __forceinline__ void bn_backward_body(const __nv_bfloat16* __restrict__ dy,
                                                          const __nv_bfloat16* __restrict__ y,
                                                          const __nv_bfloat16* __restrict__ z,
                                                          const float* __restrict__ mean,
                                                          const float* __restrict__ rstd,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          const float* __restrict__ sums, float count, int relu_mask,
                                                          __nv_bfloat16* __restrict__ dz,
                                                          __nv_bfloat16* __restrict__ dres, int N, int H, int W,
                                                          int C, const PixIdx& px) {
  const uint32_t total = (uint32_t)N * (uint32_t)(px.Hp * px.Wp * px.c8n);
  float gr8[8], m8[8], r8[8], sg8[8], sx8[8];      // gamma*rstd, mean, rstd, sum_g/cnt, sum_gx/cnt of this thread's channels
  float ga8[8], be8[8];                            // gamma, beta (relu_mask == 2: the mask is recomputed from z)
  {
    const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int cg0 = (int)(i0 - px.c8.div(i0) * (uint32_t)px.c8n);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = cg0 * 8 + k;
      m8[k] = mean[c]; r8[k] = rstd[c]; gr8[k] = gamma[c] * rstd[c];
      sg8[k] = sums[c] / count; sx8[k] = sums[C + c] / count;
      ga8[k] = gamma[c]; be8[k] = relu_mask == 2 ? beta[c] : 0.f;
    }
  }
  // NB items per iteration, all loads issued before the first one is consumed (see bn_apply_body)
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += NB * stride) {
    uint32_t idx[NB];
    uint4 ug[NB], uz[NB], uy[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      idx[j] = i0 + (uint32_t)j * stride;
      const bool in = idx[j] < total;
      ug[j] = in ? *reinterpret_cast<const uint4*>(dy + (size_t)idx[j] * 8) : make_uint4(0, 0, 0, 0);
      uz[j] = in ? *reinterpret_cast<const uint4*>(z + (size_t)idx[j] * 8) : make_uint4(0, 0, 0, 0);
      uy[j] = (in && relu_mask == 1) ? *reinterpret_cast<const uint4*>(y + (size_t)idx[j] * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const uint32_t i = idx[j];
      if (i >= total) break;
      int cg, w, h, n;
      px.decode(i, cg, w, h, n);
      uint4 o_dz = make_uint4(0, 0, 0, 0), o_dr = o_dz;
      if (h < H && w < W) {
        float g[8], yy[8], zz[8], d[8];
        unpack8(ug[j], g);
        unpack8(uz[j], zz);
        unpack8(uy[j], yy);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (relu_mask == 1 ? !(yy[k] > 0.f)
                             : (relu_mask == 2 && !(bn_affine(zz[k], ga8[k], m8[k], r8[k], be8[k]) > 0.f))) g[k] = 0.f;
          const float xhat = (zz[k] - m8[k]) * r8[k];
          d[k] = gr8[k] * (g[k] - sg8[k] - xhat * sx8[k]);
        }
        o_dz = pack8(d);
        o_dr = pack8(g);
      }
      *reinterpret_cast<uint4*>(dz + (size_t)i * 8) = o_dz;
      if (dres) *reinterpret_cast<uint4*>(dres + (size_t)i * 8) = o_dr;
    }
  }
}

template <int NB>
__global__ void __launch_bounds__(256) bn_backward_kernel(const __nv_bfloat16* __restrict__ dy,
                                                          const __nv_bfloat16* __restrict__ y,
                                                          const __nv_bfloat16* __restrict__ z,
                                                          const float* __restrict__ mean,
                                                          const float* __restrict__ rstd,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          const float* __restrict__ sums, float count, int relu_mask,
                                                          __nv_bfloat16* __restrict__ dz,
                                                          __nv_bfloat16* __restrict__ dres, int N, int H, int W,
                                                          int C, const PixIdx px) {
  pdl_launch_dependents();
  pdl_wait();
  bn_backward_body<NB>(dy, y, z, mean, rstd, gamma, beta, sums, count, relu_mask, dz, dres, N, H, W, C, px);
}

// Both passes of the BatchNorm backward in ONE cooperative launch (phase 1 = channel_reduce_kernel<1>: dbeta | dgamma,
// grid hand-over, phase 2 = bn_backward_kernel).
__global__ void __launch_bounds__(256) bn_backward_coop_kernel(const __nv_bfloat16* __restrict__ dy,
                                                               const __nv_bfloat16* __restrict__ y,
                                                               const __nv_bfloat16* __restrict__ z,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ rstd,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float count, int relu_mask,
                                                               __nv_bfloat16* __restrict__ dz,
                                                               __nv_bfloat16* __restrict__ dres, long long pixels, int N,
                                                               int H, int W, int C, float* __restrict__ sums,
                                                               float* __restrict__ partial,
                                                               unsigned* __restrict__ counter,
                                                               unsigned* __restrict__ sync, const PixIdx px) {
  pdl_launch_dependents();
  const bool last = channel_reduce_body<1>(dy, y, z, mean, rstd, gamma, beta, pixels, C, relu_mask, sums, partial, counter,
                                           BnFinalize{});
  grid_handover(sync, last);
  // dbeta | dgamma were written by another block of THIS launch: read them once, past L1, into shared memory
  __shared__ float s_sums[2 * kCoopMaxC];
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) s_sums[c] = __ldcg(sums + c);
  __syncthreads();
  bn_backward_body<2>(dy, y, z, mean, rstd, gamma, beta, s_sums, count, relu_mask, dz, dres, N, H, W, C, px);
  grid_leave(sync);
}

// ------------------------------------------------------------------ fuse-layer sum (HRnet.py:255-264) and backward
struct SumArgs {
  const __nv_bfloat16* same[4];
  const __nv_bfloat16* up[kMaxUp];
  int shift[kMaxUp];
  int n_same, n_up;
};

__global__ void __launch_bounds__(256) sum_relu_kernel(const SumArgs a, __nv_bfloat16* __restrict__ y, int N, int H,
                                                       int W, int C, const PixIdx px) {
  const uint32_t total = (uint32_t)N * (uint32_t)(px.Hp * px.Wp * px.c8n);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int cg, w, h, n;
    px.decode(i, cg, w, h, n);
    uint4 out = make_uint4(0, 0, 0, 0);
    if (h < H && w < W) {
      float f[8] = {0, 0, 0, 0, 0, 0, 0, 0}, r[8];
      for (int k = 0; k < a.n_same; ++k) {
        unpack8(*reinterpret_cast<const uint4*>(a.same[k] + (size_t)i * 8), r);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += r[j];
      }
      for (int k = 0; k < a.n_up; ++k) {
        const int s = a.shift[k];
        const size_t qs = ((size_t)n * ((H >> s) + 1) + (h >> s)) * ((W >> s) + 1) + (w >> s);
        unpack8(*reinterpret_cast<const uint4*>(a.up[k] + qs * C + cg * 8), r);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += r[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      out = pack8(f);
    }
    *reinterpret_cast<uint4*>(y + (size_t)i * 8) = out;
  }
}

// g = dy masked by (y > 0), at full resolution (gradient of every same-resolution addend)
__global__ void __launch_bounds__(256) relu_mask_kernel(const __nv_bfloat16* __restrict__ dy,
                                                        const __nv_bfloat16* __restrict__ y,
                                                        __nv_bfloat16* __restrict__ g, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    float d[8], yy[8];
    unpack8(*reinterpret_cast<const uint4*>(dy + (size_t)i * 8), d);
    unpack8(*reinterpret_cast<const uint4*>(y + (size_t)i * 8), yy);
#pragma unroll
    for (int k = 0; k < 8; ++k) if (!(yy[k] > 0.f)) d[k] = 0.f;
    *reinterpret_cast<uint4*>(g + (size_t)i * 8) = pack8(d);
  }
}

// gradient of a nearest-upsampled addend: sum of g over each 2^s x 2^s window -> low-resolution tensor
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const __nv_bfloat16* __restrict__ g,
                                                           __nv_bfloat16* __restrict__ dlow, int N, int H, int W,
                                                           int C, int s) {
  const int c8n = C / 8, Hs = H >> s, Ws = W >> s;
  const long long total = (long long)N * (Hs + 1) * (Ws + 1) * c8n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c8n);
    const long long q = i / c8n;
    const int w = (int)(q % (Ws + 1));
    const long long t = q / (Ws + 1);
    const int h = (int)(t % (Hs + 1)), n = (int)(t / (Hs + 1));
    uint4 out = make_uint4(0, 0, 0, 0);
    if (h < Hs && w < Ws) {
      float f[8] = {0, 0, 0, 0, 0, 0, 0, 0}, r[8];
      for (int dh = 0; dh < (1 << s); ++dh)
        for (int dw = 0; dw < (1 << s); ++dw) {
          const size_t qf = ((size_t)n * (H + 1) + ((h << s) + dh)) * (W + 1) + ((w << s) + dw);
          unpack8(*reinterpret_cast<const uint4*>(g + qf * C + cg * 8), r);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += r[j];
        }
      out = pack8(f);
    }
    *reinterpret_cast<uint4*>(dlow + (size_t)i * 8) = out;
  }
}

// Zero-stuffing: u[n,h,w] = dz[n,h/2,w/2] where h and w are even, 0 elsewhere.  A stride-2 convolution is the stride-1
// convolution sampled at even positions, so its input and weight gradients are the stride-1 ones of the stuffed dz.
__global__ void __launch_bounds__(256) zero_stuff_kernel(const __nv_bfloat16* __restrict__ dz,
                                                         __nv_bfloat16* __restrict__ u, int N, int H, int W, int C, const PixIdx px) {
  const int Wo = W / 2, Ho = H / 2;
  const uint32_t total = (uint32_t)N * (uint32_t)(px.Hp * px.Wp * px.c8n);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int cg, w, h, n;
    px.decode(i, cg, w, h, n);
    uint4 out = make_uint4(0, 0, 0, 0);
    if (h < H && w < W && !(h & 1) && !(w & 1)) {
      const size_t qs = ((size_t)n * (Ho + 1) + (h >> 1)) * (Wo + 1) + (w >> 1);
      out = *reinterpret_cast<const uint4*>(dz + qs * C + cg * 8);
    }
    *reinterpret_cast<uint4*>(u + (size_t)i * 8) = out;
  }
}

// ------------------------------------------------------------------ CUDA-core convolution gradients
// wgrad of the stem's first convolution (HRnet.py:286: 3 -> 64 channels, 3x3, stride 2): 1 728 outputs reduced over
// N*Ho*Wo pixels.  Thread = (output channel, pixel lane); the 27 input values of a pixel are staged in shared memory
// and broadcast, each thread keeps 27 fp32 accumulators; every block writes its partial sums to its own slab of the
// workspace and slab_sum_kernel adds the slabs in block order (bit-reproducible: no floating-point atomics).
constexpr int kStemPix = 32;   // output pixels per block iteration
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const __nv_bfloat16* __restrict__ x,
                                                         const __nv_bfloat16* __restrict__ dz,
                                                         float* __restrict__ slabs, int N, int Hi, int Wi, int Cin,
                                                         int Cout, int cin_real, long long pix_total) {
  __shared__ float patch[kStemPix][28];
  const int Ho = Hi / 2, Wo = Wi / 2, Wip = Wi + 1, Hip = Hi + 1, Wop = Wo + 1, Hop = Ho + 1;
  const int co = threadIdx.x % 64, lane4 = threadIdx.x / 64;          // 4 pixel lanes x 64 channels
  float acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.f;
  for (long long p0 = (long long)blockIdx.x * kStemPix; p0 < pix_total; p0 += (long long)gridDim.x * kStemPix) {
    __syncthreads();
    for (int i = threadIdx.x; i < kStemPix * 27; i += 256) {
      const int pp = i / 27, r = i % 27, tap = r / 3, ci = r % 3;
      const long long pix = p0 + pp;
      float v = 0.f;
      if (pix < pix_total && ci < cin_real) {
        const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
        const int h = 2 * ho + tap / 3 - 1, w = 2 * wo + tap % 3 - 1;
        if (h >= 0 && h < Hi && w >= 0 && w < Wi) v = __bfloat162float(x[(((size_t)n * Hip + h) * Wip + w) * Cin + ci]);
      }
      patch[pp][r] = v;
    }
    __syncthreads();
    for (int pp = lane4; pp < kStemPix; pp += 4) {
      const long long pix = p0 + pp;
      if (pix >= pix_total) break;
      const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
      const float g = __bfloat162float(dz[(((size_t)n * Hop + ho) * Wop + wo) * Cout + co]);
#pragma unroll
      for (int i = 0; i < 27; ++i) acc[i] += g * patch[pp][i];
    }
  }
  __shared__ float red[4][64][28];
#pragma unroll
  for (int i = 0; i < 27; ++i) red[lane4][co][i] = acc[i];
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 27; i += 256) {
    const int c = i / 27, r = i % 27, tap = r / 3, ci = r % 3;
    if (ci >= cin_real) continue;
    const float v = (red[0][c][r] + red[1][c][r]) + (red[2][c][r] + red[3][c][r]);
    slabs[(size_t)blockIdx.x * (64 * cin_real * 9) + ((size_t)c * cin_real + ci) * 9 + tap] = v;
  }
}

// out[i] = sum over slabs s (ascending) of slabs[s][i]: the fixed-order second pass of the CUDA-core weight gradients.
// Four independent partial sums keep four loads in flight; the grouping is a function of n_slabs only.
__global__ void __launch_bounds__(256) slab_sum_kernel(const float* __restrict__ slabs, float* __restrict__ out, int n,
                                                       int n_slabs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* src = slabs + i;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int sp = 0;
  for (; sp + 3 < n_slabs; sp += 4) {
    a0 += __ldcg(src + (size_t)sp * n);
    a1 += __ldcg(src + (size_t)(sp + 1) * n);
    a2 += __ldcg(src + (size_t)(sp + 2) * n);
    a3 += __ldcg(src + (size_t)(sp + 3) * n);
  }
  for (; sp < n_slabs; ++sp) a0 += __ldcg(src + (size_t)sp * n);
  out[i] = (a0 + a1) + (a2 + a3);
}


// dgrad for any (k, stride): dx[n,h,w,ci] = sum_{kh,kw,co} dz[n,ho,wo,co] * W[co,ci,kh,kw] with ho*stride+kh-pad = h.
// One thread per (pixel, 8 input channels); weights [taps][cout][cin] bf16 (the forward packing, unscaled).
__global__ void __launch_bounds__(128) conv_dgrad_kernel(const __nv_bfloat16* __restrict__ dz,
                                                         const __nv_bfloat16* __restrict__ wp,
                                                         __nv_bfloat16* __restrict__ dx, int N, int Hi, int Wi,
                                                         int Cin, int Cout, int k, int stride) {
  const int c8n = Cin / 8, Ho = Hi / stride, Wo = Wi / stride, pad = k / 2;
  const long long total = (long long)N * (Hi + 1) * (Wi + 1) * c8n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c8n);
    const long long q = i / c8n;
    const int w = (int)(q % (Wi + 1));
    const long long t = q / (Wi + 1);
    const int h = (int)(t % (Hi + 1)), n = (int)(t / (Hi + 1));
    uint4 out = make_uint4(0, 0, 0, 0);
    if (h < Hi && w < Wi) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int kh = 0; kh < k; ++kh) {
        const int hn = h + pad - kh;
        if (hn < 0 || hn % stride) continue;
        const int ho = hn / stride;
        if (ho >= Ho) continue;
        for (int kw = 0; kw < k; ++kw) {
          const int wn = w + pad - kw;
          if (wn < 0 || wn % stride) continue;
          const int wo = wn / stride;
          if (wo >= Wo) continue;
          const __nv_bfloat16* g = dz + (((size_t)n * (Ho + 1) + ho) * (Wo + 1) + wo) * Cout;
          const __nv_bfloat16* ww = wp + (size_t)(kh * k + kw) * Cout * Cin + cg * 8;
          for (int co = 0; co < Cout; ++co) {
            const float gv = __bfloat162float(g[co]);
            float r[8];
            unpack8(*reinterpret_cast<const uint4*>(ww + (size_t)co * Cin), r);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(gv, r[j], acc[j]);
          }
        }
      }
      out = pack8(acc);
    }
    *reinterpret_cast<uint4*>(dx + (size_t)i * 8) = out;
  }
}

// wgrad: dW[co][ci][kh][kw] = sum_{n,ho,wo} dz[n,ho,wo,co] * x[n,ho*s+kh-pad,wo*s+kw-pad,ci]   (fp32, OIHW)
// Block = one (tap, 16 x 16 (co, ci) tile) x one slice of the pixels; 256 threads = the 16 x 16 tile, each thread
// walks its pixel slice with both operands staged in shared memory 64 pixels at a time; the slices of one tile live in
// different blocks, each writes its partial sums into slab `slice` of the workspace (every element of a slab is written
// by exactly one thread) and slab_sum_kernel adds the slabs in slice order: bit-reproducible, no atomics.
constexpr int kWgPix = 64;
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const __nv_bfloat16* __restrict__ x,
                                                         const __nv_bfloat16* __restrict__ dz,
                                                         float* __restrict__ slabs, int N, int Hi, int Wi, int Cin,
                                                         int Cout, int k, int stride, int cin_real, int slices) {
  const int Ho = Hi / stride, Wo = Wi / stride, pad = k / 2, taps = k * k;
  const int ci_tiles = Cin / 16, co_tiles = Cout / 16;
  int b = blockIdx.x;
  const int slice = b % slices; b /= slices;
  const int cit = b % ci_tiles; b /= ci_tiles;
  const int cot = b % co_tiles; b /= co_tiles;
  const int tap = b;
  const int kh = tap / k, kw = tap % k;
  const int tci = threadIdx.x & 15, tco = threadIdx.x >> 4;
  __shared__ float sx[kWgPix][16 + 1], sg[kWgPix][16 + 1];
  const long long pix = (long long)N * Ho * Wo;
  const long long per = (pix + slices - 1) / slices;
  const long long p0 = (long long)slice * per, p1 = p0 + per < pix ? p0 + per : pix;
  float acc = 0.f;
  for (long long base = p0; base < p1; base += kWgPix) {
    // stage 64 pixels x 16 channels of both operands (256 threads: 4 pixels x 16 channels per pass, 16 passes ... 4 passes of 64x16/256)
    for (int e = threadIdx.x; e < kWgPix * 16; e += 256) {
      const int pp = e >> 4, c = e & 15;
      const long long p = base + pp;
      float xv = 0.f, gv = 0.f;
      if (p < p1) {
        const int wo = (int)(p % Wo);
        const long long t = p / Wo;
        const int ho = (int)(t % Ho), n = (int)(t / Ho);
        const int h = ho * stride + kh - pad, w = wo * stride + kw - pad;
        gv = __bfloat162float(dz[(((size_t)n * (Ho + 1) + ho) * (Wo + 1) + wo) * Cout + cot * 16 + c]);
        if (h >= 0 && h < Hi && w >= 0 && w < Wi)
          xv = __bfloat162float(x[(((size_t)n * (Hi + 1) + h) * (Wi + 1) + w) * Cin + cit * 16 + c]);
      }
      sx[pp][c] = xv;
      sg[pp][c] = gv;
    }
    __syncthreads();
#pragma unroll 16
    for (int pp = 0; pp < kWgPix; ++pp) acc = fmaf(sg[pp][tco], sx[pp][tci], acc);
    __syncthreads();
  }
  const int co = cot * 16 + tco, ci = cit * 16 + tci;
  if (ci < cin_real)
    slabs[(size_t)slice * ((size_t)Cout * cin_real * taps) + ((size_t)co * cin_real + ci) * taps + tap] = acc;
}

// Grid for a 256-thread grid-stride kernel over `total` items whose stride (grid * 256) must be a multiple of `c8n`
// (threads then keep their channel group): c8n = 2^a or 3 * 2^a -> round the block count up to a multiple of 3 if needed.
int grid_mult(long long total, int c8n, int nb = 1) {
  long long g = (total + 256 * nb - 1) / (256 * nb);   // (the kernels take nb items per thread and iteration)
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  if (256 % c8n) g = (g + 2) / 3 * 3;
  return (int)g;
}

int grid_for(long long total, int block, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// Block count of the channel reductions (<= kReduceBlocks rows of partial sums): a multiple of 3 when the channel-group
// count has a factor 3, so that the cooperative kernels - whose second phase needs that - reduce in the same order and
// give the same bits as the separate launches.
int reduce_grid(long long pixels, int C) {
  const int lanes = 256 / (C / 8);
  int g = grid_for(pixels, lanes * 8, kReduceBlocks);
  if (256 % (C / 8)) g = (g + 2) / 3 * 3 > kReduceBlocks ? kReduceBlocks / 3 * 3 : (g + 2) / 3 * 3;
  return g;
}

}  // namespace

namespace {
// STLPOSE_TRAIN_PDL=1: the BatchNorm kernels of the separate-launch path carry the programmatic stream serialization
// attribute (they wait with griddepcontrol.wait before touching memory), so that their launch latency hides behind the
// tail of the preceding kernel, inside captured graphs too.  Measured on B200 together with the same attribute on the
// training convolutions: no gain (17.6-17.8 vs 17.3-17.6 ms per step at batch 32, 43.4 vs 43.1 at 128) - off by default.
bool train_pdl() {
  static const bool on = getenv("STLPOSE_TRAIN_PDL") && atoi(getenv("STLPOSE_TRAIN_PDL")) == 1;
  return on;
}
template <typename... KArgs, typename... Args>
void launch_bn(void (*kern)(KArgs...), int grid, cudaStream_t st, Args&&... args) {
  if (!train_pdl()) {
    kern<<<grid, 256, 0, st>>>(args...);
    return;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);   // (errors surface in check())
}
}  // namespace

size_t bn_workspace_floats(int C) { return (size_t)2 * C * (1 + kReduceBlocks) + 32; }

int bn_train_forward(const __nv_bfloat16* z, const float* gamma, const float* beta, const __nv_bfloat16* residual,
                     int relu, float eps, float momentum, int N, int H, int W, int C, __nv_bfloat16* y, float* sums,
                     float* mean, float* rstd, float* run_mean, float* run_var, unsigned* ticket, cudaStream_t st) {
  if (C % 8 || C > 2048) { set_error("bn_train_forward: C=%d unsupported", C); return 1; }
  const long long pixels = (long long)N * (H + 1) * (W + 1);
  // the last-block ticket: a caller-owned word that is zero between launches (the kernel resets it), or the tail of the
  // workspace cleared here
  unsigned* counter = ticket ? ticket : reinterpret_cast<unsigned*>(sums + (size_t)2 * C * (1 + kReduceBlocks));
  if (!ticket) cudaMemsetAsync(counter, 0, sizeof(unsigned), st);
  BnFinalize fin{(float)((long long)N * H * W), eps, momentum, mean, rstd, run_mean, run_var};
  launch_bn(channel_reduce_kernel<0>, reduce_grid(pixels, C), st, z, (const __nv_bfloat16*)nullptr,
            (const __nv_bfloat16*)nullptr, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr,
            (const float*)nullptr, pixels, C, 0, sums, sums + 2 * C, counter, fin);
  if (check("bn stats")) return 1;
  PixIdx px;
  px.init(C, H, W);
  if (pixels * (C / 8) >= (1ll << 31)) { set_error("bn_train_forward: tensor too large for 32-bit item indexing"); return 1; }
  if (256 % (C / 8) && (C / 8) % 3) { set_error("bn_train_forward: C=%d unsupported", C); return 1; }
  // items per thread and iteration (STL_BN_NB_APPLY / _BWD = 1, 2, 4: measurement knobs).  Measured on B200: the backward
  // kernel gains most from two (13.7 -> 8.7 us at 6.5 MB, 81 -> 58 us at 52 MB, 270 -> 167 us at 209 MB; four: another
  // 10-15 % on the large tensors only), the normalisation little (5.0 -> 4.6 us at 6.5 MB, unchanged from 26 MB up)
  static const int nb = getenv("STL_BN_NB_APPLY") ? atoi(getenv("STL_BN_NB_APPLY")) : 2;
  if (nb == 4)
    launch_bn(bn_apply_kernel<4>, grid_mult(pixels * (C / 8), C / 8, 4), st, z, (const float*)mean, (const float*)rstd, gamma,
              beta, residual, relu, y, N, H, W, C, px);
  else if (nb == 2)
    launch_bn(bn_apply_kernel<2>, grid_mult(pixels * (C / 8), C / 8, 2), st, z, (const float*)mean, (const float*)rstd, gamma,
              beta, residual, relu, y, N, H, W, C, px);
  else
    launch_bn(bn_apply_kernel<1>, grid_mult(pixels * (C / 8), C / 8, 1), st, z, (const float*)mean, (const float*)rstd, gamma,
              beta, residual, relu, y, N, H, W, C, px);
  return check("bn apply");
}

namespace {
// largest co-resident grid of a cooperative BatchNorm kernel on the current device (blocks per SM x SMs), cached per device
template <typename K>
int coop_capacity(K kernel, int slot) {
  static std::mutex mu;
  static int cap[2][DeviceOnce::kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= DeviceOnce::kMaxDevices) return 0;
  std::lock_guard<std::mutex> lock(mu);
  if (!cap[slot][dev]) {
    int per_sm = 0, coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) == cudaSuccess && per_sm > 0)
      cap[slot][dev] = per_sm * device_sm_count();
    else
      cap[slot][dev] = -1;
  }
  return cap[slot][dev] > 0 ? cap[slot][dev] : 0;
}
// grid of the fused kernels: the reduction's block count, a multiple of 3 when the channel-group count has a factor 3
// (the normalisation phase keeps a thread on its channel group), within the co-resident capacity
int coop_grid(long long pixels, int C, int capacity) {
  int g = reduce_grid(pixels, C);
  if (g > capacity) g = capacity / 3 * 3;
  return g;
}
}  // namespace

int bn_train_forward_coop(const __nv_bfloat16* z, const float* gamma, const float* beta, const __nv_bfloat16* residual,
                          int relu, float eps, float momentum, int N, int H, int W, int C, __nv_bfloat16* y, float* sums,
                          float* mean, float* rstd, float* run_mean, float* run_var, unsigned* ticket, unsigned* sync,
                          cudaStream_t st) {
  if (C % 8 || C > 2048 || (256 % (C / 8) && (C / 8) % 3)) { set_error("bn_train_forward_coop: C=%d unsupported", C); return 1; }
  if (C > kCoopMaxC)
    return bn_train_forward(z, gamma, beta, residual, relu, eps, momentum, N, H, W, C, y, sums, mean, rstd, run_mean, run_var,
                            ticket, st);
  const long long pixels = (long long)N * (H + 1) * (W + 1);
  if (pixels * (C / 8) >= (1ll << 31)) { set_error("bn_train_forward_coop: tensor too large for 32-bit item indexing"); return 1; }
  const int cap = coop_capacity(bn_forward_coop_kernel, 0);
  const int grid = cap ? coop_grid(pixels, C, cap) : 0;
  if (grid < 1)    // no cooperative launch on this device: the two separate launches
    return bn_train_forward(z, gamma, beta, residual, relu, eps, momentum, N, H, W, C, y, sums, mean, rstd, run_mean, run_var,
                            ticket, st);
  BnFinalize fin{(float)((long long)N * H * W), eps, momentum, mean, rstd, run_mean, run_var};
  PixIdx px;
  px.init(C, H, W);
  long long pixels_ = pixels;
  float* partial = sums + 2 * C;
  void* args[] = {(void*)&z, (void*)&gamma, (void*)&beta, (void*)&residual, (void*)&relu, (void*)&y, (void*)&pixels_,
                  (void*)&N, (void*)&H, (void*)&W, (void*)&C, (void*)&sums, (void*)&partial, (void*)&ticket, (void*)&sync,
                  (void*)&fin, (void*)&px};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)bn_forward_coop_kernel, dim3(grid), dim3(256), args, 0, st);
  if (e != cudaSuccess) { set_error("bn_forward_coop launch: %s", cudaGetErrorString(e)); return 1; }
  return check("bn forward (cooperative)");
}

int bn_train_backward_coop(const __nv_bfloat16* dy, const __nv_bfloat16* y, const __nv_bfloat16* z, const float* mean,
                           const float* rstd, const float* gamma, const float* beta, int relu, int N, int H, int W, int C,
                           __nv_bfloat16* dz, __nv_bfloat16* dres, float* sums, float* partial, unsigned* ticket,
                           unsigned* sync, cudaStream_t st) {
  if (C % 8 || C > 2048 || (256 % (C / 8) && (C / 8) % 3)) { set_error("bn_train_backward_coop: C=%d unsupported", C); return 1; }
  if (relu == 2 && (!beta || dres)) { set_error("bn_train_backward_coop: the z-recomputed mask needs beta and no residual"); return 1; }
  const long long pixels = (long long)N * (H + 1) * (W + 1);
  if (pixels * (C / 8) >= (1ll << 31)) { set_error("bn_train_backward_coop: tensor too large for 32-bit item indexing"); return 1; }
  const int cap = C <= kCoopMaxC ? coop_capacity(bn_backward_coop_kernel, 1) : 0;
  const int grid = cap ? coop_grid(pixels, C, cap) : 0;
  if (grid < 1)
    return bn_train_backward(dy, y, z, mean, rstd, gamma, beta, relu, N, H, W, C, dz, dres, sums, partial, ticket, st);
  PixIdx px;
  px.init(C, H, W);
  long long pixels_ = pixels;
  float count = (float)((long long)N * H * W);
  void* args[] = {(void*)&dy, (void*)&y, (void*)&z, (void*)&mean, (void*)&rstd, (void*)&gamma, (void*)&beta, (void*)&count,
                  (void*)&relu, (void*)&dz, (void*)&dres, (void*)&pixels_, (void*)&N, (void*)&H, (void*)&W, (void*)&C,
                  (void*)&sums, (void*)&partial, (void*)&ticket, (void*)&sync, (void*)&px};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)bn_backward_coop_kernel, dim3(grid), dim3(256), args, 0, st);
  if (e != cudaSuccess) { set_error("bn_backward_coop launch: %s", cudaGetErrorString(e)); return 1; }
  return check("bn backward (cooperative)");
}

int bn_train_forward_fused(const __nv_bfloat16* z, const float* stat_rows, int rows, int c_pad, const float* gamma,
                           const float* beta, const __nv_bfloat16* residual, int relu, float eps, float momentum, int N,
                           int H, int W, int C, __nv_bfloat16* y, float* mean, float* rstd, float* run_mean, float* run_var,
                           cudaStream_t st) {
  if (C % 8 || C > 2048 || rows < 1 || c_pad < C) { set_error("bn_train_forward_fused: bad arguments"); return 1; }
  const long long pixels = (long long)N * (H + 1) * (W + 1);
  BnFinalize fin{(float)((long long)N * H * W), eps, momentum, mean, rstd, run_mean, run_var};
  bn_finalize_kernel<<<(C + 31) / 32, 256, 0, st>>>(stat_rows, rows, C, c_pad, fin);
  if (check("bn finalize")) return 1;
  PixIdx px;
  px.init(C, H, W);
  if (pixels * (C / 8) >= (1ll << 31)) { set_error("bn_train_forward_fused: tensor too large for 32-bit item indexing"); return 1; }
  if (256 % (C / 8) && (C / 8) % 3) { set_error("bn_train_forward_fused: C=%d unsupported", C); return 1; }
  bn_apply_kernel<1><<<grid_mult(pixels * (C / 8), C / 8), 256, 0, st>>>(z, mean, rstd, gamma, beta, residual, relu, y, N, H,
                                                                  W, C, px);
  return check("bn apply");
}

int bn_apply(const __nv_bfloat16* z, const float* mean, const float* rstd, const float* gamma, const float* beta,
             const __nv_bfloat16* residual, int relu, int N, int H, int W, int C, __nv_bfloat16* y, cudaStream_t st) {
  if (C % 8 || C > 2048) { set_error("bn_apply: C=%d unsupported", C); return 1; }
  const long long pixels = (long long)N * (H + 1) * (W + 1);
  PixIdx px;
  px.init(C, H, W);
  if (pixels * (C / 8) >= (1ll << 31)) { set_error("bn_apply: tensor too large for 32-bit item indexing"); return 1; }
  if (256 % (C / 8) && (C / 8) % 3) { set_error("bn_apply: C=%d unsupported", C); return 1; }
  bn_apply_kernel<1><<<grid_mult(pixels * (C / 8), C / 8), 256, 0, st>>>(z, mean, rstd, gamma, beta, residual, relu, y, N, H,
                                                                  W, C, px);
  return check("bn apply");
}

int bn_train_backward(const __nv_bfloat16* dy, const __nv_bfloat16* y, const __nv_bfloat16* z, const float* mean,
                      const float* rstd, const float* gamma, const float* beta, int relu, int N, int H, int W, int C,
                      __nv_bfloat16* dz, __nv_bfloat16* dres, float* sums, float* partial, unsigned* ticket,
                      cudaStream_t st) {
  if (C % 8 || C > 2048) { set_error("bn_train_backward: C=%d unsupported", C); return 1; }
  const long long pixels = (long long)N * (H + 1) * (W + 1);
  // partial == null: one workspace [sums 2C | partial rows | ticket]; else `sums` ([2C], dbeta | dgamma) is a tensor of its own
  if (!partial) partial = sums + 2 * C;
  unsigned* counter = ticket ? ticket : reinterpret_cast<unsigned*>(partial + (size_t)2 * C * kReduceBlocks);
  if (!ticket) cudaMemsetAsync(counter, 0, sizeof(unsigned), st);
  if (relu == 2 && (!beta || dres)) { set_error("bn_train_backward: the z-recomputed mask needs beta and no residual"); return 1; }
  launch_bn(channel_reduce_kernel<1>, reduce_grid(pixels, C), st, dy, y, z, mean, rstd, gamma, beta, pixels, C, relu, sums,
            partial, counter, BnFinalize{});
  if (check("bn backward reduce")) return 1;
  PixIdx px;
  px.init(C, H, W);
  if (pixels * (C / 8) >= (1ll << 31)) { set_error("bn_train_backward: tensor too large for 32-bit item indexing"); return 1; }
  if (256 % (C / 8) && (C / 8) % 3) { set_error("bn_train_backward: C=%d unsupported", C); return 1; }
  static const int nb = getenv("STL_BN_NB_BWD") ? atoi(getenv("STL_BN_NB_BWD")) : 2;
  if (nb == 4)
    launch_bn(bn_backward_kernel<4>, grid_mult(pixels * (C / 8), C / 8, 4), st, dy, y, z, mean, rstd, gamma, beta,
              (const float*)sums, (float)((long long)N * H * W), relu, dz, dres, N, H, W, C, px);
  else if (nb == 1)
    launch_bn(bn_backward_kernel<1>, grid_mult(pixels * (C / 8), C / 8, 1), st, dy, y, z, mean, rstd, gamma, beta,
              (const float*)sums, (float)((long long)N * H * W), relu, dz, dres, N, H, W, C, px);
  else
    launch_bn(bn_backward_kernel<2>, grid_mult(pixels * (C / 8), C / 8, 2), st, dy, y, z, mean, rstd, gamma, beta,
              (const float*)sums, (float)((long long)N * H * W), relu, dz, dres, N, H, W, C, px);
  return check("bn backward");
}

int sum_relu_forward(const __nv_bfloat16* const* same, int n_same, const __nv_bfloat16* const* up, const int* shift,
                     int n_up, __nv_bfloat16* y, int N, int H, int W, int C, cudaStream_t st) {
  if (n_same < 0 || n_same > 4 || n_up < 0 || n_up > kMaxUp || C % 8) { set_error("sum_relu: bad arguments"); return 1; }
  SumArgs a{};
  a.n_same = n_same; a.n_up = n_up;
  for (int i = 0; i < n_same; ++i) a.same[i] = same[i];
  for (int i = 0; i < n_up; ++i) { a.up[i] = up[i]; a.shift[i] = shift[i]; }
  const long long total = (long long)N * (H + 1) * (W + 1) * (C / 8);
  PixIdx px;
  px.init(C, H, W);
  if (total >= (1ll << 31)) { set_error("sum_relu: tensor too large for 32-bit item indexing"); return 1; }
  sum_relu_kernel<<<grid_for(total, 256), 256, 0, st>>>(a, y, N, H, W, C, px);
  return check("sum_relu");
}

int relu_mask(const __nv_bfloat16* dy, const __nv_bfloat16* y, __nv_bfloat16* g, long long elems, cudaStream_t st) {
  relu_mask_kernel<<<grid_for(elems / 8, 256), 256, 0, st>>>(dy, y, g, elems / 8);
  return check("relu_mask");
}

int upsample_backward(const __nv_bfloat16* g, __nv_bfloat16* dlow, int N, int H, int W, int C, int shift,
                      cudaStream_t st) {
  const long long total = (long long)N * ((H >> shift) + 1) * ((W >> shift) + 1) * (C / 8);
  upsample_bwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(g, dlow, N, H, W, C, shift);
  return check("upsample_backward");
}

int conv_dgrad_naive(const __nv_bfloat16* dz, const __nv_bfloat16* w_packed, __nv_bfloat16* dx, int N, int Hi, int Wi,
                     int Cin, int Cout, int k, int stride, cudaStream_t st) {
  if (Cin % 8) { set_error("conv_dgrad: Cin must be a multiple of 8"); return 1; }
  const long long total = (long long)N * (Hi + 1) * (Wi + 1) * (Cin / 8);
  conv_dgrad_kernel<<<grid_for(total, 128, 148 * 32), 128, 0, st>>>(dz, w_packed, dx, N, Hi, Wi, Cin, Cout, k, stride);
  return check("conv_dgrad");
}

int zero_stuff(const __nv_bfloat16* dz, __nv_bfloat16* u, int N, int H, int W, int C, cudaStream_t st) {
  if (C % 8 || (H & 1) || (W & 1)) { set_error("zero_stuff: C %% 8 and even H, W required"); return 1; }
  const long long total = (long long)N * (H + 1) * (W + 1) * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  PixIdx px;
  px.init(C, H, W);
  if (total >= (1ll << 31)) { set_error("zero_stuff: tensor too large for 32-bit item indexing"); return 1; }
  zero_stuff_kernel<<<(int)blocks, 256, 0, st>>>(dz, u, N, H, W, C, px);
  return check("zero_stuff");
}

namespace {
bool is_stem_wgrad(int Cout, int k, int stride, int cin_real) { return k == 3 && stride == 2 && Cout == 64 && cin_real <= 3; }
int stem_wgrad_blocks(long long pix) { return grid_for(pix, kStemPix, 148 * 4); }
int wgrad_slices(long long pix, int Cin, int Cout, int k) {
  const int tiles = k * k * (Cin / 16) * (Cout / 16);
  int slices = (148 * 8 + tiles - 1) / tiles;
  const long long max_slices = (pix + 4 * kWgPix - 1) / (4 * kWgPix);
  if (slices > max_slices) slices = (int)max_slices;
  return slices < 1 ? 1 : slices;
}
}  // namespace

size_t conv_wgrad_naive_workspace_bytes(int N, int Hi, int Wi, int Cin, int Cout, int k, int stride, int cin_real) {
  if (stride < 1 || k < 1) return 0;
  const long long pix = (long long)N * (Hi / stride) * (Wi / stride);
  const size_t n = (size_t)Cout * cin_real * k * k;
  const int slabs = is_stem_wgrad(Cout, k, stride, cin_real) ? stem_wgrad_blocks(pix) : wgrad_slices(pix, Cin, Cout, k);
  return sizeof(float) * n * slabs;
}

int conv_wgrad_naive(const __nv_bfloat16* x, const __nv_bfloat16* dz, float* dw, int N, int Hi, int Wi, int Cin,
                     int Cout, int k, int stride, int cin_real, void* workspace, size_t workspace_bytes,
                     cudaStream_t st) {
  if (Cin % 16 || Cout % 16) { set_error("conv_wgrad: channels must be multiples of 16"); return 1; }
  const size_t need = conv_wgrad_naive_workspace_bytes(N, Hi, Wi, Cin, Cout, k, stride, cin_real);
  if (!workspace || workspace_bytes < need) {
    set_error("conv_wgrad: workspace of %zu bytes required (stl_conv_wgrad_workspace_bytes), got %zu", need,
              workspace_bytes);
    return 1;
  }
  float* slabs = static_cast<float*>(workspace);
  const int n = Cout * cin_real * k * k;
  const long long pix = (long long)N * (Hi / stride) * (Wi / stride);
  int n_slabs;
  if (is_stem_wgrad(Cout, k, stride, cin_real)) {      // the stem's first convolution
    n_slabs = stem_wgrad_blocks(pix);
    stem_wgrad_kernel<<<n_slabs, 256, 0, st>>>(x, dz, slabs, N, Hi, Wi, Cin, Cout, cin_real, pix);
    if (check("stem_wgrad")) return 1;
  } else {
    const int tiles = k * k * (Cin / 16) * (Cout / 16);
    n_slabs = wgrad_slices(pix, Cin, Cout, k);
    conv_wgrad_kernel<<<tiles * n_slabs, 256, 0, st>>>(x, dz, slabs, N, Hi, Wi, Cin, Cout, k, stride, cin_real, n_slabs);
    if (check("conv_wgrad")) return 1;
  }
  slab_sum_kernel<<<(n + 255) / 256, 256, 0, st>>>(slabs, dw, n, n_slabs);
  return check("wgrad slab sum");
}

}  // namespace stl

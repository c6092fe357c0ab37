// One BasicBlock of a C = 32 branch in ONE kernel (HRnet.py:45-61, eval mode, BatchNorm folded):
//
//     y = relu( conv2( relu(conv1(x) + b1) ) + b2 + x )        both convs 3x3 / stride 1 / 32 -> 32
//
// Run as two launches of the conv kernel, these layers are bound by HBM (read x, write mid, read mid, read x, write y:
// five tensor passes) and by the A-operand fetch of N = 32 MMAs.  Fused, the intermediate activation never leaves the
// SM: conv1's accumulators are converted to bf16 straight into shared memory, in exactly the K-major swizzled layout
// the tensor core reads, and become conv2's A operand (two passes over HBM: read x, write y; x is read a second time
// for the residual, from L2).
//
// Tile = 384 output pixels of the padded-linear layout (conv.h).  conv2 needs conv1 on a halo of Wp+1 pixels either
// side, so conv1 is evaluated on 512 pixels [q0-h, q0-h+512) (h = Wp+1 <= 64) from an x tile of 512 + 2h pixels; the
// taps of both convs are row-shifted UMMA descriptors on those tiles.  Zero cells of the layout and pixels outside the
// tensor are written as zeros into the intermediate tile (they are conv2's zero padding).
//
// Warps: 0 TMA producer (weights of both convs once, then x tiles, 3 stages; a stage is overwritten in place by the
// tile's intermediate activation once conv1 has consumed it and is released by conv2) | 1 MMA issuer of conv1 | 19 MMA issuer of
// conv2 (conv2 of tile i-1 overlaps conv1 of tile i) | 2-9 epilogue 1 (TMEM -> +b1, ReLU -> bf16 -> intermediate tile) |
// 10-17 epilogue 2 (TMEM -> +b2 + residual, ReLU -> staging panels) | 18 DMA (TMA: residual panels in, finished panels
// out).  All hand-offs are mbarriers.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "conv.h"
#include "ptx.cuh"

namespace stl {
namespace {

constexpr int kC = 32;                 // channels
constexpr int kPitch = kC * 2;         // bytes per pixel row in shared memory = swizzle span (64)
constexpr int kOutRows = 384;          // output pixels per tile (3 accumulator blocks)
constexpr int kMidRows = 512;          // conv1 pixels per tile (4 accumulator blocks)
constexpr int kMaxHalo = 64;
constexpr int kXPieces = 3;
constexpr int kBlkThreads = 20 * 32;
constexpr int kTapBytes = kC * kPitch; // one filter tap: 32 output channels x 64 B = 2 KB

struct BlockParams {
  CUtensorMap tmX, tmW1, tmW2, tmR, tmO;
  const float* bias1;
  const float* bias2;
  int Wp, Hp, H, W, halo;
  long long P;                         // padded pixel count
  int x_piece_rows;                    // rows per TMA piece of the x tile (3 pieces)
  uint32_t x_stage_bytes;
  long long total_tiles;
  FastDiv fd_Wp, fd_Hp;
};

struct BlkCtl {
  uint64_t w_full;
  uint64_t x_full[3], x_empty[3];      // x tile stages; a stage then holds the tile's intermediate activation
  uint64_t acc1_full[2], acc1_empty[2];
  uint64_t acc2_full[2], acc2_empty[2];
  uint64_t mid_full[3];
  uint64_t res_full[2], out_done[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float blo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bhi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_relu(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// is padded pixel q a real pixel of the tensor (not a zero cell, inside [0, P))?
__device__ __forceinline__ bool real_pixel(const BlockParams& p, long long q) {
  if (q < 0 || q >= p.P) return false;
  const uint32_t t = p.fd_Wp.div((uint32_t)q);
  const int w = (int)q - (int)t * p.Wp;
  const int n = (int)p.fd_Hp.div(t);
  const int h = (int)t - n * p.Hp;
  return w != p.W && h != p.H;
}

__global__ void __launch_bounds__(kBlkThreads, 1) basic_block_kernel(const __grid_constant__ BlockParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  BlkCtl* ctl = reinterpret_cast<BlkCtl*>(smem);
  float* sbias = reinterpret_cast<float*>(smem + 512);            // [2][32]
  const uint32_t base = smem_u32(smem) + 1024;
  const uint32_t w_base = base;                                   // 2 convs x 9 taps x 2 KB = 36 KB
  const uint32_t x_base = w_base + 2 * 9 * kTapBytes;             // 3 stages: x tile, then (in place) the intermediate tile
  const uint32_t out_base = x_base + 3 * p.x_stage_bytes;         // 2 buffers x 3 panels x 8 KB

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tile0 = blockIdx.x, tstride = gridDim.x;
  const int halo = p.halo;

  // programmatic dependent launch (conv_tc.cu): the next kernel may start its prologue as our CTAs retire; everything here
  // that touches activations waits for the previous kernel below (producer: x tiles; DMA warp: residual loads AND stores -
  // the output buffer may be one the previous kernel is still reading)
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    mbar_init(&ctl->w_full, 1);
    for (int i = 0; i < 3; ++i) { mbar_init(&ctl->x_full[i], 1); mbar_init(&ctl->x_empty[i], 1); mbar_init(&ctl->mid_full[i], 8); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->acc1_full[i], 1); mbar_init(&ctl->acc1_empty[i], 8);
      mbar_init(&ctl->acc2_full[i], 1); mbar_init(&ctl->acc2_empty[i], 8);
      mbar_init(&ctl->res_full[i], 1); mbar_init(&ctl->out_done[i], 8);
    }
    fence_mbar_init();
    tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmW1); tma_prefetch_desc(&p.tmW2);
    tma_prefetch_desc(&p.tmR); tma_prefetch_desc(&p.tmO);
  }
  if (threadIdx.x < 64) sbias[threadIdx.x] = threadIdx.x < 32 ? p.bias1[threadIdx.x] : p.bias2[threadIdx.x - 32];
  if (warp == 1) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  // TMEM columns: conv1 accumulators 2 x (4 blocks x 32) at [0, 256), conv2 accumulators 2 x (3 x 32) at [256, 448)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(&ctl->w_full, 2u * 9u * kTapBytes);
      for (int t = 0; t < 9; ++t) {
        tma_load_3d_s(w_base + t * kTapBytes, &p.tmW1, &ctl->w_full, 0, 0, t);
        tma_load_3d_s(w_base + (9 + t) * kTapBytes, &p.tmW2, &ctl->w_full, 0, 0, t);
      }
      pdl_wait();
      uint32_t i = 0;
      for (long long tile = tile0; tile < p.total_tiles; tile += tstride, ++i) {
        const uint32_t s = i % 3, ph = (i / 3) & 1;
        mbar_wait(&ctl->x_empty[s], ph ^ 1u);                    // conv2 of the tile that last used this stage is done
        const long long q1 = tile * kOutRows - halo;             // first conv1 pixel; x tile starts another halo earlier
        mbar_expect_tx(&ctl->x_full[s], (uint32_t)(kXPieces * p.x_piece_rows * kPitch));
        for (int pc = 0; pc < kXPieces; ++pc)
          tma_load_2d_s(x_base + s * p.x_stage_bytes + (uint32_t)(pc * p.x_piece_rows * kPitch), &p.tmX, &ctl->x_full[s], 0,
                        (int)(q1 - halo) + pc * p.x_piece_rows);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer of conv1 (x tile -> accumulators 1)
    const uint32_t idesc = make_idesc_bf16(128, kC);
    mbar_wait(&ctl->w_full, 0);
    tc_fence_after();
    int offs[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) offs[t] = (t / 3 - 1) * p.Wp + (t % 3 - 1) + halo;     // tap row offset + halo >= 0
    uint32_t i = 0;
    for (long long tile = tile0; tile < p.total_tiles; tile += tstride, ++i) {
      const uint32_t s = i & 1, ph = (i >> 1) & 1, s3 = i % 3, ph3 = (i / 3) & 1;
      mbar_wait(&ctl->x_full[s3], ph3);
      mbar_wait(&ctl->acc1_empty[s], ph ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t xs = x_base + s3 * p.x_stage_bytes;
        const uint32_t d0 = tmem_base + s * 128u;
#pragma unroll 1
        for (int m = 0; m < 4; ++m) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t ad = make_kmajor_desc(xs + (uint32_t)((m * 128 + offs[t]) * kPitch + ks * 32), kPitch);
              const uint64_t bd = make_kmajor_desc(w_base + (uint32_t)(t * kTapBytes + ks * 32), kPitch);
              umma_bf16(d0 + (uint32_t)(m * kC), ad, bd, idesc, (t | ks) ? 1u : 0u);
            }
          }
        }
        umma_commit(&ctl->acc1_full[s]);
      }
      __syncwarp();
    }
  } else if (warp == 19) {
    // ------------------------------------------------------------------ MMA issuer of conv2 (intermediate tile -> accumulators 2)
    // A second issuing thread: one thread sustains a tcgen05.mma every ~50 cycles whatever its size, and these
    // 128x32x16 MMAs are shorter than that; conv2 of tile i-1 overlaps conv1 of tile i on the tensor pipe.
    const uint32_t idesc = make_idesc_bf16(128, kC);
    mbar_wait(&ctl->w_full, 0);
    tc_fence_after();
    int offs[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) offs[t] = (t / 3 - 1) * p.Wp + (t % 3 - 1) + halo;
    uint32_t j = 0;
    for (long long tile = tile0; tile < p.total_tiles; tile += tstride, ++j) {
      const uint32_t s = j & 1, ph = (j >> 1) & 1, s3 = j % 3, ph3 = (j / 3) & 1;
      mbar_wait(&ctl->mid_full[s3], ph3);
      mbar_wait(&ctl->acc2_empty[s], ph ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t mid_base = x_base + s3 * p.x_stage_bytes;
        const uint32_t d0 = tmem_base + 256u + s * 96u;
#pragma unroll 1
        for (int m = 0; m < 3; ++m) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t ad = make_kmajor_desc(mid_base + (uint32_t)((m * 128 + offs[t]) * kPitch + ks * 32), kPitch);
              const uint64_t bd = make_kmajor_desc(w_base + (uint32_t)((9 + t) * kTapBytes + ks * 32), kPitch);
              umma_bf16(d0 + (uint32_t)(m * kC), ad, bd, idesc, (t | ks) ? 1u : 0u);
            }
          }
        }
        umma_commit(&ctl->x_empty[s3]);                           // stage free for the x tile three tiles ahead
        umma_commit(&ctl->acc2_full[s]);
      }
      __syncwarp();
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ epilogue 1: conv1 accumulators -> intermediate tile
    const int e = warp - 2, quarter = warp & 3, sub = e >> 2;    // two warps per TMEM lane quarter, one 16-channel slice each
    const int row0 = quarter * 32 + lane;
    const float4* b4 = reinterpret_cast<const float4*>(sbias + sub * 16);
    const float4 bias[4] = {b4[0], b4[1], b4[2], b4[3]};
    const uint32_t xr = (uint32_t)((row0 >> 1) & 3);
    uint32_t i = 0;
    for (long long tile = tile0; tile < p.total_tiles; tile += tstride, ++i) {
      const uint32_t s = i & 1, ph = (i >> 1) & 1, s3 = i % 3;
      const long long q1 = tile * kOutRows - halo;
      const uint32_t mid_base = x_base + s3 * p.x_stage_bytes;   // conv1 has consumed the x tile: overwrite it in place
      mbar_wait(&ctl->acc1_full[s], ph);
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < 4; ++m) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + s * 128u + (uint32_t)(m * kC + sub * 16), v);
        const int r = m * 128 + row0;
        const bool real = real_pixel(p, q1 + r);
        tmem_ld_wait();
        float* f = reinterpret_cast<float*>(v);
        uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
        if (real) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            fadd2(v[g * 4 + 0], v[g * 4 + 1], __float_as_uint(bias[g].x), __float_as_uint(bias[g].y));
            fadd2(v[g * 4 + 2], v[g * 4 + 3], __float_as_uint(bias[g].z), __float_as_uint(bias[g].w));
          }
          o0 = make_uint4(pack_relu(f[0], f[1]), pack_relu(f[2], f[3]), pack_relu(f[4], f[5]), pack_relu(f[6], f[7]));
          o1 = make_uint4(pack_relu(f[8], f[9]), pack_relu(f[10], f[11]), pack_relu(f[12], f[13]), pack_relu(f[14], f[15]));
        }
        const uint32_t rowaddr = mid_base + (uint32_t)r * kPitch;
        const uint32_t c0 = (uint32_t)sub * 2u;
        sts128(rowaddr + ((c0 ^ xr) << 4), o0);
        sts128(rowaddr + (((c0 + 1) ^ xr) << 4), o1);
      }
      tc_fence_before();
      fence_async_smem();                                        // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) { mbar_arrive(&ctl->acc1_empty[s]); mbar_arrive(&ctl->mid_full[s3]); }
    }
  } else if (warp < 18) {
    // ------------------------------------------------------------------ epilogue 2: conv2 accumulators + residual -> panels
    const int e = warp - 10, quarter = warp & 3, sub = e >> 2;
    const int row0 = quarter * 32 + lane;
    const float4* b4 = reinterpret_cast<const float4*>(sbias + 32 + sub * 16);
    const float4 bias[4] = {b4[0], b4[1], b4[2], b4[3]};
    const uint32_t xr = (uint32_t)((row0 >> 1) & 3);
    uint32_t i = 0;
    for (long long tile = tile0; tile < p.total_tiles; tile += tstride, ++i) {
      const uint32_t s = i & 1, ph = (i >> 1) & 1;
      const long long q0 = tile * kOutRows;
      mbar_wait(&ctl->acc2_full[s], ph);
      tc_fence_after();
      mbar_wait(&ctl->res_full[s], ph);                          // residual panels landed in staging buffer s
#pragma unroll 1
      for (int m = 0; m < 3; ++m) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + 256u + s * 96u + (uint32_t)(m * kC + sub * 16), v);
        const bool real = real_pixel(p, q0 + m * 128 + row0);
        const uint32_t rowaddr = out_base + (uint32_t)((s * 3 + m) * 128 * kPitch) + (uint32_t)row0 * kPitch;
        const uint32_t c0 = (uint32_t)sub * 2u;
        const uint32_t a0 = rowaddr + ((c0 ^ xr) << 4), a1 = rowaddr + (((c0 + 1) ^ xr) << 4);
        const uint4 r0 = lds128(a0), r1 = lds128(a1);
        tmem_ld_wait();
        float* f = reinterpret_cast<float*>(v);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          fadd2(v[g * 4 + 0], v[g * 4 + 1], __float_as_uint(bias[g].x), __float_as_uint(bias[g].y));
          fadd2(v[g * 4 + 2], v[g * 4 + 3], __float_as_uint(bias[g].z), __float_as_uint(bias[g].w));
        }
        fadd2(v[0], v[1], r0.x << 16, r0.x & 0xFFFF0000u);   fadd2(v[2], v[3], r0.y << 16, r0.y & 0xFFFF0000u);
        fadd2(v[4], v[5], r0.z << 16, r0.z & 0xFFFF0000u);   fadd2(v[6], v[7], r0.w << 16, r0.w & 0xFFFF0000u);
        fadd2(v[8], v[9], r1.x << 16, r1.x & 0xFFFF0000u);   fadd2(v[10], v[11], r1.y << 16, r1.y & 0xFFFF0000u);
        fadd2(v[12], v[13], r1.z << 16, r1.z & 0xFFFF0000u); fadd2(v[14], v[15], r1.w << 16, r1.w & 0xFFFF0000u);
        uint4 o0 = make_uint4(pack_relu(f[0], f[1]), pack_relu(f[2], f[3]), pack_relu(f[4], f[5]), pack_relu(f[6], f[7]));
        uint4 o1 = make_uint4(pack_relu(f[8], f[9]), pack_relu(f[10], f[11]), pack_relu(f[12], f[13]), pack_relu(f[14], f[15]));
        if (!real) { o0 = make_uint4(0, 0, 0, 0); o1 = o0; }     // zero cells of the padded layout stay zero
        sts128(a0, o0);
        sts128(a1, o1);
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&ctl->acc2_empty[s]); mbar_arrive(&ctl->out_done[s]); }
    }
  } else if (warp == 18) {
    // ------------------------------------------------------------------ DMA warp: residual panels in, finished panels out
    if (lane == 0) {
      pdl_wait();
      auto load_res = [&](long long tile, uint32_t s) {
        mbar_expect_tx(&ctl->res_full[s], 3u * 128u * kPitch);
        for (int m = 0; m < 3; ++m)
          tma_load_2d_s(out_base + (uint32_t)((s * 3 + m) * 128 * kPitch), &p.tmR, &ctl->res_full[s], 0,
                        (int)(tile * kOutRows) + m * 128);
      };
      if (tile0 < p.total_tiles) load_res(tile0, 0);
      if (tile0 + tstride < p.total_tiles) load_res(tile0 + tstride, 1);
      uint32_t i = 0;
      for (long long tile = tile0; tile < p.total_tiles; tile += tstride, ++i) {
        const uint32_t s = i & 1, ph = (i >> 1) & 1;
        mbar_wait(&ctl->out_done[s], ph);
        for (int m = 0; m < 3; ++m)
          tma_store_2d_s(&p.tmO, out_base + (uint32_t)((s * 3 + m) * 128 * kPitch), 0, (int)(tile * kOutRows) + m * 128);
        bulk_commit();
        if (tile + 2 * tstride < p.total_tiles) {
          bulk_wait_read<0>();                                   // the stores have read buffer s: it can take the next residual
          load_res(tile + 2 * tstride, s);
        }
      }
      bulk_wait<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncFn enc_fn() {
  static EncFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncFn>(ptr);
  }
  return fn;
}
int enc(CUtensorMap* tm, const void* basep, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
        const cuuint32_t* box) {
  EncFn fn = enc_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return 1; }
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(basep), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("basic_block: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return 1; }
  return 0;
}

}  // namespace

bool basic_block_supported(int H, int W, int C) { return C == kC && W + 2 <= kMaxHalo && H > 0; }

// x, y: padded-linear bf16 [N][H+1][W+1][32]; w1, w2: packed [9][32][32] bf16; b1, b2: 32 fp32 (folded BatchNorm).
int basic_block_launch(const __nv_bfloat16* x, __nv_bfloat16* y, const __nv_bfloat16* w1, const float* b1,
                       const __nv_bfloat16* w2, const float* b2, int N, int H, int W, int max_ctas, cudaStream_t stream,
                       int pdl) {
  if (!basic_block_supported(H, W, kC)) { set_error("basic_block: unsupported geometry %dx%d", H, W); return 1; }
  BlockParams p{};
  p.Wp = W + 1; p.Hp = H + 1; p.H = H; p.W = W; p.halo = W + 2;
  p.P = (long long)N * p.Hp * p.Wp;
  if (p.P + kMidRows >= (1ll << 31)) { set_error("basic_block: tensor too large"); return 1; }
  p.bias1 = b1; p.bias2 = b2;
  p.fd_Wp.init((uint32_t)p.Wp); p.fd_Hp.init((uint32_t)p.Hp);
  const int x_rows = kMidRows + 2 * p.halo;
  p.x_piece_rows = (((x_rows + kXPieces - 1) / kXPieces) + 7) & ~7;
  p.x_stage_bytes = ((uint32_t)(kXPieces * p.x_piece_rows * kPitch) + 1023u) & ~1023u;
  p.total_tiles = (p.P + kOutRows - 1) / kOutRows;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kC, (cuuint64_t)p.P};
    const cuuint64_t strides[1] = {(cuuint64_t)kPitch};
    const cuuint32_t bx[2] = {(cuuint32_t)kC, (cuuint32_t)p.x_piece_rows};
    const cuuint32_t bp[2] = {(cuuint32_t)kC, 128};
    if (enc(&p.tmX, x, 2, dims, strides, bx)) return 1;
    if (enc(&p.tmR, x, 2, dims, strides, bp)) return 1;
    if (enc(&p.tmO, y, 2, dims, strides, bp)) return 1;
    const cuuint64_t wd[3] = {(cuuint64_t)kC, (cuuint64_t)kC, 9};
    const cuuint64_t ws[2] = {(cuuint64_t)kPitch, (cuuint64_t)kPitch * kC};
    const cuuint32_t wb[3] = {(cuuint32_t)kC, (cuuint32_t)kC, 1};
    if (enc(&p.tmW1, w1, 3, wd, ws, wb)) return 1;
    if (enc(&p.tmW2, w2, 3, wd, ws, wb)) return 1;
  }
  const size_t smem = 1024 + 1024 + 2 * 9 * kTapBytes + 3 * (size_t)p.x_stage_bytes + 2 * 3 * 128 * kPitch;
  if (smem > 227 * 1024) { set_error("basic_block: shared memory budget exceeded (%zu)", smem); return 1; }
  static DeviceOnce attr_once;
  cudaError_t e;
  if (attr_once.run([]() {
        cudaError_t e2 = cudaFuncSetAttribute(basic_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e2 != cudaSuccess) { set_error("basic_block attribute: %s", cudaGetErrorString(e2)); return 1; }
        return 0;
      }))
    return 1;
  const int sms = device_sm_count();
  long long grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if (pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kBlkThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, basic_block_kernel, p);
    if (e != cudaSuccess) { set_error("basic_block (attributed) launch: %s", cudaGetErrorString(e)); return 1; }
  } else {
    basic_block_kernel<<<(unsigned)grid, kBlkThreads, smem, stream>>>(p);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("basic_block launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace stl

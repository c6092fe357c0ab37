// The junction of two layer1 Bottlenecks in ONE kernel (HRnet.py:64-102, eval mode, BatchNorm folded):
//
//     out = relu( conv3(t) + b3 + x )          1x1, 64 -> 256, x = the block's input (residual)
//     a   = relu( conv1'(out) + b1' )          1x1, 256 -> 64, conv1 of the NEXT Bottleneck
//
// Both convolutions are bound by HBM (the 256-channel tensor at 64x48: 1.67 GB per 1 024 images): as two launches `out` is
// written by the first and read back by the second.  Here the finished bf16 tile of `out` - staged in shared memory for
// its TMA store anyway, in exactly the K-major swizzled layout the tensor core reads - is also the A operand of the
// second GEMM, so `out` crosses HBM once.
//
// Tile = 128 consecutive pixels of the padded-linear layout (conv.h).  A tile is processed as two UNITS of 128 output
// channels each: GEMM1 (K = 64, N = 128) -> epilogue 1 (+ b3 + residual, ReLU, bf16, in place over the residual panels in
// a staging buffer) -> TMA store of the two 64-channel panels AND GEMM2 partial sum (K = those 128 channels, N = 64);
// after the second unit: epilogue 2 (+ b1', ReLU) -> `a` tile -> TMA store.  Three 32 KB staging buffers: the residual of
// unit u+2 is in flight while unit u is converted and unit u-1 drains.
//
// Variant KC = 2 (layer1.0 -> layer1.1): GEMM1 runs over TWO input tensors, K = [t | x64] with the concatenated weights
// [W3 | Wd] - conv3 + downsample of the first Bottleneck (conv.h, ConvSpec::in2) - and there is no residual; two staging
// buffers and one t stage keep it inside 208 KB of shared memory.
//
// Warps: 0 TMA producer (weights once, then t tiles, 2 stages) | 1 MMA issuer of GEMM1 | 19 MMA issuer of GEMM2 |
// 2-9 epilogue 1 | 10-17 epilogue 2 | 18 DMA (residual panels in, `out` panels and `a` tiles out).  All hand-offs are
// mbarriers.  TMEM: GEMM1 accumulators 2 x 128 columns, GEMM2 accumulators 2 x 64 columns.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "conv.h"
#include "ptx.cuh"

namespace stl {
namespace {

constexpr int kCt = 64;                 // channels of t (conv3's input)
constexpr int kCo = 256;                // channels of out
constexpr int kCa = 64;                 // channels of a (conv1' output)
constexpr int kRows = 128;              // pixels per tile
constexpr int kRowBytes = 128;          // one shared-memory row: 64 bf16 = the 128-byte swizzle span
constexpr int kPanel = kRows * kRowBytes;           // 16 KB: 128 pixels x 64 channels
constexpr int kLinkThreads = 20 * 32;
constexpr uint32_t kW3Bytes = kCo * kRowBytes;      // 32 KB per K chunk: [256 output channels][64 input channels]
constexpr uint32_t kW1Bytes = 4 * kCa * kRowBytes;  // 32 KB: 4 K chunks of [64 output channels][64 input channels]

struct LinkParams {
  CUtensorMap tmT, tmT2, tmW3, tmW1, tmR, tmO, tmA;
  const float* bias3;
  const float* bias1;
  int Wp, Hp, H, W;
  long long P;                          // padded pixel count
  long long total_tiles;
  FastDiv fd_Wp, fd_Hp;
};

struct LinkCtl {
  uint64_t w_full;
  uint64_t t_full[2], t_empty[2];       // t tile stages                      (producer <-> GEMM1 issuer)
  uint64_t acc1_full[2], acc1_empty[2]; // GEMM1 accumulators, per unit       (GEMM1 issuer <-> epilogue 1)
  uint64_t res_full[3];                 // residual panels landed in a staging buffer        (DMA -> epilogue 1)
  uint64_t out_ready[3];                // bf16 `out` panels written to a staging buffer     (epilogue 1 -> DMA, GEMM2)
  uint64_t g2_done[3];                  // GEMM2 has read a staging buffer                   (GEMM2 issuer -> DMA)
  uint64_t acc2_full[2], acc2_empty[2]; // GEMM2 accumulators, per tile       (GEMM2 issuer <-> epilogue 2)
  uint64_t a_ready, a_free;             // `a` tile staging buffer            (epilogue 2 <-> DMA)
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_relu2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// is padded pixel q a real pixel of the tensor (not a zero cell, inside [0, P))?
__device__ __forceinline__ bool real_pixel(const LinkParams& p, long long q) {
  if (q >= p.P) return false;
  const uint32_t t = p.fd_Wp.div((uint32_t)q);
  const int w = (int)q - (int)t * p.Wp;
  const int n = (int)p.fd_Hp.div(t);
  const int h = (int)t - n * p.Hp;
  return w != p.W && h != p.H;
}

// KC: input tensors of GEMM1 (K chunks of 64 channels); HAS_RES: residual added in epilogue 1; NSTG: staging buffers;
// NTST: t tile stages
template <int KC, bool HAS_RES, int NSTG, int NTST>
__global__ void __launch_bounds__(kLinkThreads, 1) bottleneck_link_kernel(const __grid_constant__ LinkParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  LinkCtl* ctl = reinterpret_cast<LinkCtl*>(smem);
  float* sbias = reinterpret_cast<float*>(smem + 512);            // [256] b3 | [64] b1'
  const uint32_t base = smem_u32(smem) + 2048;
  const uint32_t w3_base = base;
  const uint32_t w1_base = w3_base + KC * kW3Bytes;
  const uint32_t t_base = w1_base + kW1Bytes;                     // NTST stages x KC x 16 KB
  const uint32_t stg_base = t_base + NTST * KC * kPanel;          // NSTG buffers x 2 panels x 16 KB
  const uint32_t a_base = stg_base + NSTG * 2 * kPanel;           // 16 KB

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tile0 = blockIdx.x, tstride = gridDim.x;
  const uint32_t n_my = tile0 < p.total_tiles ? (uint32_t)((p.total_tiles - tile0 + tstride - 1) / tstride) : 0u;
  const uint32_t n_units = 2u * n_my;

  // programmatic dependent launch (conv_tc.cu): the producer and the DMA warp - the only ones that touch activations -
  // wait for the previous kernel before their first load / store
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    mbar_init(&ctl->w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->t_full[i], 1); mbar_init(&ctl->t_empty[i], 1);
      mbar_init(&ctl->acc1_full[i], 1); mbar_init(&ctl->acc1_empty[i], 8);
      mbar_init(&ctl->acc2_full[i], 1); mbar_init(&ctl->acc2_empty[i], 8);
    }
    for (int i = 0; i < 3; ++i) { mbar_init(&ctl->res_full[i], 1); mbar_init(&ctl->out_ready[i], 8); mbar_init(&ctl->g2_done[i], 1); }
    mbar_init(&ctl->a_ready, 8);
    mbar_init(&ctl->a_free, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.tmT); tma_prefetch_desc(&p.tmW3); tma_prefetch_desc(&p.tmW1);
    if (KC == 2) tma_prefetch_desc(&p.tmT2);
    if (HAS_RES) tma_prefetch_desc(&p.tmR);
    tma_prefetch_desc(&p.tmO); tma_prefetch_desc(&p.tmA);
  }
  for (int i = threadIdx.x; i < kCo + kCa; i += kLinkThreads) sbias[i] = i < kCo ? p.bias3[i] : p.bias1[i - kCo];
  if (warp == 1) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  // TMEM columns: GEMM1 accumulators 2 x 128 at [0, 256), GEMM2 accumulators 2 x 64 at [256, 384)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(&ctl->w_full, KC * kW3Bytes + kW1Bytes);
      for (int kc = 0; kc < KC; ++kc) tma_load_2d_s(w3_base + kc * kW3Bytes, &p.tmW3, &ctl->w_full, kc * 64, 0);
      for (int c = 0; c < 4; ++c) tma_load_2d_s(w1_base + (uint32_t)(c * kCa * kRowBytes), &p.tmW1, &ctl->w_full, c * 64, 0);
      pdl_wait();
      for (uint32_t i = 0; i < n_my; ++i) {
        const uint32_t s = i % NTST, ph = (i / NTST) & 1;
        const int row = (int)((tile0 + (long long)i * tstride) * kRows);
        mbar_wait(&ctl->t_empty[s], ph ^ 1u);
        mbar_expect_tx(&ctl->t_full[s], (uint32_t)(KC * kPanel));
        tma_load_2d_s(t_base + s * (KC * kPanel), &p.tmT, &ctl->t_full[s], 0, row);
        if (KC == 2) tma_load_2d_s(t_base + s * (KC * kPanel) + kPanel, &p.tmT2, &ctl->t_full[s], 0, row);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ GEMM1 issuer: t tile x W3 half -> accumulators 1
    const uint32_t idesc = make_idesc_bf16(128, 128);
    mbar_wait(&ctl->w_full, 0);
    tc_fence_after();
    for (uint32_t u = 0; u < n_units; ++u) {
      const uint32_t i = u >> 1, h = u & 1, s = i % NTST, b = u & 1;
      if (h == 0) mbar_wait(&ctl->t_full[s], (i / NTST) & 1);
      mbar_wait(&ctl->acc1_empty[b], ((u >> 1) & 1) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = make_kmajor_desc(t_base + s * (KC * kPanel) + kc * kPanel + (uint32_t)(ks * 32), kRowBytes);
            const uint64_t bd = make_kmajor_desc(w3_base + kc * kW3Bytes + h * (uint32_t)(128 * kRowBytes) + (uint32_t)(ks * 32),
                                                 kRowBytes);
            umma_bf16(tmem_base + b * 128u, ad, bd, idesc, (kc | ks) ? 1u : 0u);
          }
        }
        umma_commit(&ctl->acc1_full[b]);
        if (h == 1) umma_commit(&ctl->t_empty[s]);
      }
      __syncwarp();
    }
  } else if (warp == 19) {
    // ------------------------------------------------------------------ GEMM2 issuer: `out` panels x W1' -> accumulators 2
    const uint32_t idesc = make_idesc_bf16(128, kCa);
    mbar_wait(&ctl->w_full, 0);
    tc_fence_after();
    for (uint32_t u = 0; u < n_units; ++u) {
      const uint32_t i = u >> 1, h = u & 1, sb = u % NSTG, b2 = i & 1;
      mbar_wait(&ctl->out_ready[sb], (u / NSTG) & 1);
      if (h == 0) mbar_wait(&ctl->acc2_empty[b2], ((i >> 1) & 1) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int pn = 0; pn < 2; ++pn) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = make_kmajor_desc(stg_base + (sb * 2 + pn) * kPanel + (uint32_t)(ks * 32), kRowBytes);
            const uint64_t bd = make_kmajor_desc(w1_base + (2 * h + pn) * (uint32_t)(kCa * kRowBytes) + (uint32_t)(ks * 32), kRowBytes);
            umma_bf16(tmem_base + 256u + b2 * 64u, ad, bd, idesc, (h | pn | ks) ? 1u : 0u);
          }
        }
        umma_commit(&ctl->g2_done[sb]);
        if (h == 1) umma_commit(&ctl->acc2_full[b2]);
      }
      __syncwarp();
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ epilogue 1: accumulators 1 + b3 + residual -> `out` panels
    const int quarter = warp & 3, sub = (warp - 2) >> 2;         // two warps per TMEM lane quarter, alternate 16-channel slices
    const int row0 = quarter * 32 + lane;
    const uint32_t xr = (uint32_t)(row0 & 7);
    for (uint32_t u = 0; u < n_units; ++u) {
      const uint32_t i = u >> 1, h = u & 1, sb = u % NSTG, b = u & 1;
      const long long q = (tile0 + (long long)i * tstride) * kRows + row0;
      const bool real = real_pixel(p, q);
      mbar_wait(&ctl->acc1_full[b], (u >> 1) & 1);
      tc_fence_after();
      mbar_wait(&ctl->res_full[sb], (u / NSTG) & 1);             // residual landed, or (no residual) the buffer is free
#pragma unroll 1
      for (int sl = sub; sl < 8; sl += 2) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + b * 128u + (uint32_t)(sl * 16), v);
        const uint32_t rowaddr = stg_base + (sb * 2 + (uint32_t)(sl >> 2)) * kPanel + (uint32_t)row0 * kRowBytes;
        const uint32_t c0 = (uint32_t)(sl & 3) * 2u;
        const uint32_t a0 = rowaddr + ((c0 ^ xr) << 4), a1 = rowaddr + (((c0 + 1) ^ xr) << 4);
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
        if (HAS_RES) { r0 = lds128(a0); r1 = lds128(a1); }
        const float4* b4 = reinterpret_cast<const float4*>(sbias + h * 128 + sl * 16);
        const float4 bias[4] = {b4[0], b4[1], b4[2], b4[3]};
        tmem_ld_wait();
        float* f = reinterpret_cast<float*>(v);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          fadd2(v[g * 4 + 0], v[g * 4 + 1], __float_as_uint(bias[g].x), __float_as_uint(bias[g].y));
          fadd2(v[g * 4 + 2], v[g * 4 + 3], __float_as_uint(bias[g].z), __float_as_uint(bias[g].w));
        }
        if (HAS_RES) {
          fadd2(v[0], v[1], r0.x << 16, r0.x & 0xFFFF0000u);   fadd2(v[2], v[3], r0.y << 16, r0.y & 0xFFFF0000u);
          fadd2(v[4], v[5], r0.z << 16, r0.z & 0xFFFF0000u);   fadd2(v[6], v[7], r0.w << 16, r0.w & 0xFFFF0000u);
          fadd2(v[8], v[9], r1.x << 16, r1.x & 0xFFFF0000u);   fadd2(v[10], v[11], r1.y << 16, r1.y & 0xFFFF0000u);
          fadd2(v[12], v[13], r1.z << 16, r1.z & 0xFFFF0000u); fadd2(v[14], v[15], r1.w << 16, r1.w & 0xFFFF0000u);
        }
        uint4 o0 = make_uint4(pack_relu2(f[0], f[1]), pack_relu2(f[2], f[3]), pack_relu2(f[4], f[5]), pack_relu2(f[6], f[7]));
        uint4 o1 = make_uint4(pack_relu2(f[8], f[9]), pack_relu2(f[10], f[11]), pack_relu2(f[12], f[13]), pack_relu2(f[14], f[15]));
        if (!real) { o0 = make_uint4(0, 0, 0, 0); o1 = o0; }     // zero cells of the padded layout stay zero
        sts128(a0, o0);
        sts128(a1, o1);
      }
      tc_fence_before();
      fence_async_smem();                                        // generic-proxy stores -> visible to TMA and the tensor core
      __syncwarp();
      if (lane == 0) { mbar_arrive(&ctl->acc1_empty[b]); mbar_arrive(&ctl->out_ready[sb]); }
    }
  } else if (warp < 18) {
    // ------------------------------------------------------------------ epilogue 2: accumulators 2 + b1' -> `a` tile
    const int quarter = warp & 3, sub = (warp - 10) >> 2;
    const int row0 = quarter * 32 + lane;
    const uint32_t xr = (uint32_t)(row0 & 7);
    const uint32_t rowaddr = a_base + (uint32_t)row0 * kRowBytes;
    for (uint32_t i = 0; i < n_my; ++i) {
      const uint32_t b2 = i & 1;
      const long long q = (tile0 + (long long)i * tstride) * kRows + row0;
      const bool real = real_pixel(p, q);
      mbar_wait(&ctl->acc2_full[b2], (i >> 1) & 1);
      tc_fence_after();
      mbar_wait(&ctl->a_free, (i & 1) ^ 1u);                     // the previous tile's `a` store has read the buffer
#pragma unroll 1
      for (int sl = sub; sl < 4; sl += 2) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + 256u + b2 * 64u + (uint32_t)(sl * 16), v);
        const float4* b4 = reinterpret_cast<const float4*>(sbias + kCo + sl * 16);
        const float4 bias[4] = {b4[0], b4[1], b4[2], b4[3]};
        tmem_ld_wait();
        float* f = reinterpret_cast<float*>(v);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          fadd2(v[g * 4 + 0], v[g * 4 + 1], __float_as_uint(bias[g].x), __float_as_uint(bias[g].y));
          fadd2(v[g * 4 + 2], v[g * 4 + 3], __float_as_uint(bias[g].z), __float_as_uint(bias[g].w));
        }
        uint4 o0 = make_uint4(pack_relu2(f[0], f[1]), pack_relu2(f[2], f[3]), pack_relu2(f[4], f[5]), pack_relu2(f[6], f[7]));
        uint4 o1 = make_uint4(pack_relu2(f[8], f[9]), pack_relu2(f[10], f[11]), pack_relu2(f[12], f[13]), pack_relu2(f[14], f[15]));
        if (!real) { o0 = make_uint4(0, 0, 0, 0); o1 = o0; }
        const uint32_t c0 = (uint32_t)sl * 2u;
        sts128(rowaddr + ((c0 ^ xr) << 4), o0);
        sts128(rowaddr + (((c0 + 1) ^ xr) << 4), o1);
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&ctl->acc2_empty[b2]); mbar_arrive(&ctl->a_ready); }
    }
  } else if (warp == 18) {
    // ------------------------------------------------------------------ DMA warp: residual panels in, `out` panels and `a` tiles out
    if (lane == 0 && n_units > 0) {
      pdl_wait();
      auto row_of = [&](uint32_t u) { return (int)((tile0 + (long long)(u >> 1) * tstride) * kRows); };
      // staging buffer of unit u: its residual panels (or, without a residual, just "free") -> res_full
      auto load_res = [&](uint32_t u) {
        const uint32_t sb = u % NSTG;
        if (HAS_RES) {
          mbar_expect_tx(&ctl->res_full[sb], 2u * kPanel);
          for (int pn = 0; pn < 2; ++pn)
            tma_load_2d_s(stg_base + (sb * 2 + pn) * kPanel, &p.tmR, &ctl->res_full[sb], (int)(u & 1) * 128 + pn * 64, row_of(u));
        } else {
          mbar_arrive(&ctl->res_full[sb]);
        }
      };
      for (uint32_t u = 0; u + 1 < (uint32_t)NSTG && u < n_units; ++u) load_res(u);
      bool a_inflight = false;
      for (uint32_t u = 0; u < n_units; ++u) {
        const uint32_t sb = u % NSTG;
        if (u >= 1) {
          bulk_wait_read<0>();                                   // every store issued so far has read its source
          if (a_inflight) { mbar_arrive(&ctl->a_free); a_inflight = false; }
        }
        if (u + NSTG - 1 < n_units) {                            // buffer of unit u-1: drained by its store (above) and by GEMM2
          if (u >= 1) mbar_wait(&ctl->g2_done[(u - 1) % NSTG], ((u - 1) / NSTG) & 1);
          load_res(u + NSTG - 1);
        }
        if (u >= 2 && !(u & 1)) {                                // the previous tile's `a`
          const uint32_t ip = (u >> 1) - 1;
          mbar_wait(&ctl->a_ready, ip & 1);
          tma_store_2d_s(&p.tmA, a_base, 0, row_of(u - 2));
          bulk_commit();
          a_inflight = true;
        }
        mbar_wait(&ctl->out_ready[sb], (u / NSTG) & 1);
        for (int pn = 0; pn < 2; ++pn)
          tma_store_2d_s(&p.tmO, stg_base + (sb * 2 + pn) * kPanel, (int)(u & 1) * 128 + pn * 64, row_of(u));
        bulk_commit();
      }
      mbar_wait(&ctl->a_ready, (n_my - 1) & 1);
      tma_store_2d_s(&p.tmA, a_base, 0, row_of(n_units - 1));
      bulk_commit();
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
EncFn link_enc_fn() {
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
// 2-D bf16 tensor [rows][cols] (cols contiguous), box = 64 columns x box_rows rows, 128-byte swizzle
int enc2d(CUtensorMap* tm, const void* basep, long long cols, long long rows, int box_rows) {
  EncFn fn = link_enc_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return 1; }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(basep), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("bottleneck_link: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return 1; }
  return 0;
}

}  // namespace

bool bottleneck_link_supported(int ct, int co, int ca) { return ct == kCt && co == kCo && ca == kCa; }

namespace {
template <int KC, bool HAS_RES, int NSTG, int NTST>
int launch_variant(const LinkParams& p, long long grid, cudaStream_t stream, int pdl) {
  auto kern = bottleneck_link_kernel<KC, HAS_RES, NSTG, NTST>;
  const size_t smem = 1024 + 2048 + KC * kW3Bytes + kW1Bytes + (size_t)NTST * KC * kPanel + (size_t)NSTG * 2 * kPanel + kPanel;
  static_assert(1024 + 2048 + KC * kW3Bytes + kW1Bytes + NTST * KC * kPanel + NSTG * 2 * kPanel + kPanel <= 227 * 1024,
                "bottleneck_link: shared memory budget");
  static DeviceOnce attr_once;   // (one per instantiation)
  if (attr_once.run([&]() {
        cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e2 != cudaSuccess) { set_error("bottleneck_link attribute: %s", cudaGetErrorString(e2)); return 1; }
        return 0;
      }))
    return 1;
  cudaError_t e;
  if (pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kLinkThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, p);
    if (e != cudaSuccess) { set_error("bottleneck_link (attributed) launch: %s", cudaGetErrorString(e)); return 1; }
  } else {
    kern<<<(unsigned)grid, kLinkThreads, smem, stream>>>(p);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("bottleneck_link launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}
}  // namespace

// t: padded-linear bf16 [N][H+1][W+1][64]; out: [..][256]; a: [..][64]; w1: packed [64][256] bf16; b3: 256 fp32,
// b1: 64 fp32 (folded BatchNorm).  Either  x (residual, [..][256], must not alias out) with w3 packed [256][64],
// or       t2 (second input of GEMM1, [..][64]) with w3 packed [256][128] = [W3 | Wd] and b3 the summed bias, no residual.
int bottleneck_link_launch(const __nv_bfloat16* t, const __nv_bfloat16* t2, const __nv_bfloat16* x, __nv_bfloat16* out,
                           __nv_bfloat16* a, const __nv_bfloat16* w3, const float* b3, const __nv_bfloat16* w1,
                           const float* b1, int N, int H, int W, int max_ctas, cudaStream_t stream, int pdl) {
  if (N <= 0 || H <= 0 || W <= 0) { set_error("bottleneck_link: bad geometry"); return 1; }
  if (!t || !out || !a || !w3 || !b3 || !w1 || !b1 || (!x && !t2) || (x && t2)) {
    set_error("bottleneck_link: null pointer (exactly one of residual / second input)");
    return 1;
  }
  if (out == x) { set_error("bottleneck_link: out must not alias the residual"); return 1; }
  LinkParams p{};
  p.Wp = W + 1; p.Hp = H + 1; p.H = H; p.W = W;
  p.P = (long long)N * p.Hp * p.Wp;
  if (p.P + kRows >= (1ll << 31)) { set_error("bottleneck_link: tensor too large"); return 1; }
  p.bias3 = b3; p.bias1 = b1;
  p.fd_Wp.init((uint32_t)p.Wp); p.fd_Hp.init((uint32_t)p.Hp);
  p.total_tiles = (p.P + kRows - 1) / kRows;
  const int kc = t2 ? 2 : 1;
  if (enc2d(&p.tmT, t, kCt, p.P, kRows)) return 1;
  if (t2 && enc2d(&p.tmT2, t2, kCt, p.P, kRows)) return 1;
  if (x && enc2d(&p.tmR, x, kCo, p.P, kRows)) return 1;
  if (enc2d(&p.tmO, out, kCo, p.P, kRows)) return 1;
  if (enc2d(&p.tmA, a, kCa, p.P, kRows)) return 1;
  if (enc2d(&p.tmW3, w3, kCt * kc, kCo, kCo)) return 1;    // one 64-column box of all 256 rows per K chunk
  if (enc2d(&p.tmW1, w1, kCo, kCa, kCa)) return 1;         // four boxes of 64 columns x 64 rows
  const int sms = device_sm_count();
  long long grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  return t2 ? launch_variant<2, false, 2, 1>(p, grid, stream, pdl) : launch_variant<1, true, 3, 2>(p, grid, stream, pdl);
}

}  // namespace stl

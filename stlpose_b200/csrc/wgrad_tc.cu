// Weight gradient of a stride-1 convolution on the tcgen05 tensor cores.
//
//   dW[co][ci][kh][kw] = sum_q dz[q][co] * x[q + (kh-1)*Wp + (kw-1)][ci]        q over the padded-linear pixels
//
// (the reference gets this from autograd of nn.Conv2d, HRnet.py:48-59 under 02_train.py:216).  In the padded-linear
// layout (conv.h) the filter taps are flat row shifts, zero cells contribute zero, and rows outside the tensor are
// zero-filled by TMA, so the whole batch is ONE GEMM per tap with the pixel index as the reduction dimension:
// M = output channels, N = input channels, K = pixels.  Both operands are "MN-major" for the tensor core (channels
// contiguous, K = rows), which the UMMA shared-memory descriptor supports for bf16 directly: a TMA box of
// [rows][<=64 channels] with the hardware swizzle IS the canonical MN-major layout, and a filter tap is a row
// offset added to the descriptor's start address.
//
// Narrow layers fill the 128 accumulator rows with row-shifted replicas of dz: with co = 32 the A descriptor's
// leading-dimension stride is one pixel row, so MN blocks 0..3 of the same tile are dz shifted by 0..3 pixels and a
// single MMA yields three taps of a filter row ( sum_q dz[q+j] x[q+b] = dW(offset b-j) ); with co = 64 two replicas
// give two taps per MMA.  Pixel tiles therefore start at row -4 (TMA zero-fills) so that every replica covers [0,P).
//
// Work decomposition: CTA = (128-row block of co, 128-column block of ci, tap group, pixel range); partial sums are
// accumulated in TMEM over the CTA's pixel range and added to the fp32 OIHW result with red.global.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "conv.h"
#include "ptx.cuh"
#include "train_kernels.h"

namespace stl {
namespace {

constexpr int kKT = 128;          // pixel rows per pipeline stage
constexpr int kLead = 4;          // tiles start at row -kLead (room for 3 replica shifts)
constexpr int kMaxMma = 6;        // MMAs per 16-row K step
constexpr int kMaxStagesW = 8;
constexpr int kThreadsW = 192;    // warp 0: TMA, warp 1: MMA, warps 2-5: epilogue

struct WgradParams {
  CUtensorMap tmDz, tmX;
  int m_real, m_blks, n_blks, n_groups, nt, n_mma, n_acc;
  int a_panels, b_panels, pitch_a, pitch_b, dz_rows, x_rows;
  uint32_t a_panel_bytes, b_panel_bytes, stage_bytes, lbo_a, lbo_b;
  int stages, tmem_cols;
  int x_row0[3];                  // first x row of a tile relative to q0, per tap group
  int mma_b[3][kMaxMma];          // B operand row offset inside the x tile, per tap group and MMA
  int mma_acc[kMaxMma];
  int tap_of[3][kMaxMma][4];      // filter tap produced by (group, accumulator, replica) or -1
  int tiles_total, tiles_per_split, n_splits;
  float* dw;
  float* ws;                      // [grid][n_acc][128][nt] fp32 partial sums, one slab per CTA
  int dbg;                        // measurement only: 1 = no MMAs, 2 = no TMA loads (results are garbage)
  int co, ci, ci_real, taps;
};

// MN-major operand: rows (K) are `pitch` bytes apart (pitch = swizzle span), 8-row groups contiguous, 64-channel
// (32 for the 64-byte span) MN blocks `lbo` bytes apart.
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr, uint32_t pitch, uint32_t lbo) {
  const uint64_t layout = pitch == 128 ? 2ull : 4ull;
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((8u * pitch) >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= layout << 61;
  return d;
}

__global__ void __launch_bounds__(kThreadsW, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStagesW], empty_bar[kMaxStagesW], done_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;

  // work item
  const int group = blockIdx.x % p.n_groups, split = blockIdx.x / p.n_groups;
  const int tg = group % ((p.n_groups / (p.m_blks * p.n_blks)));
  const int mn = group / (p.n_groups / (p.m_blks * p.n_blks));
  const int n_blk = mn % p.n_blks, m_blk = mn / p.n_blks;
  const int t_lo = split * p.tiles_per_split;
  int t_hi = t_lo + p.tiles_per_split;
  if (t_hi > p.tiles_total) t_hi = p.tiles_total;
  const int n_tiles = t_hi > t_lo ? t_hi - t_lo : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (n_tiles > 0) {
    if (warp == 0) {
      // ------------------------------------------------------------ TMA producer
      if (elect_one()) {
        tma_prefetch_desc(&p.tmDz);
        tma_prefetch_desc(&p.tmX);
        for (int i = 0; i < n_tiles; ++i) {
          const int s = i % p.stages;
          const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          const int q0 = -kLead + (t_lo + i) * kKT;
          const uint32_t a_dst = smem0 + (uint32_t)s * p.stage_bytes;
          const uint32_t b_dst = a_dst + (uint32_t)p.a_panels * p.a_panel_bytes;
          if (p.dbg == 2) { mbar_arrive(&full_bar[s]); continue; }
          mbar_expect_tx(&full_bar[s], (uint32_t)p.a_panels * (uint32_t)(p.dz_rows * p.pitch_a) +
                                           (uint32_t)p.b_panels * (uint32_t)(p.x_rows * p.pitch_b));
          for (int a = 0; a < p.a_panels; ++a)
            tma_load_2d_s(a_dst + a * p.a_panel_bytes, &p.tmDz, &full_bar[s], m_blk * 128 + a * 64, q0);
          for (int b = 0; b < p.b_panels; ++b)
            tma_load_2d_s(b_dst + b * p.b_panel_bytes, &p.tmX, &full_bar[s], n_blk * p.nt + b * 64, q0 + p.x_row0[tg]);
        }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------ MMA issuer
      // dbg 3 / 4 (measurement only, garbage results): A / A and B addressed as K-major operands - what would the MMA
      // stream cost if the tiles were transposed first?
      const uint32_t idesc = make_idesc_bf16(128, (uint32_t)p.nt) | (p.dbg >= 3 ? 0u : (1u << 15)) |
                             (p.dbg >= 4 ? 0u : (1u << 16));                                  // A and B MN-major
      for (int i = 0; i < n_tiles; ++i) {
        const int s = i % p.stages;
        const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_src = smem0 + (uint32_t)s * p.stage_bytes;
          const uint32_t b_src = a_src + (uint32_t)p.a_panels * p.a_panel_bytes;
#pragma unroll 1
          for (int ks = 0; ks < (p.dbg == 1 ? 0 : kKT / 16); ++ks) {
            const uint64_t adesc = p.dbg >= 3 ? make_kmajor_desc(a_src + (uint32_t)((ks & 3) * 32), 128u)
                                              : make_mnmajor_desc(a_src + (uint32_t)(ks * 16 * p.pitch_a), (uint32_t)p.pitch_a, p.lbo_a);
#pragma unroll 1
            for (int m = 0; m < p.n_mma; ++m) {
              const uint64_t bdesc = p.dbg >= 4 ? make_kmajor_desc(b_src + (uint32_t)((ks & 3) * 32 + m * 1024), 128u)
                                                : make_mnmajor_desc(b_src + (uint32_t)((p.mma_b[tg][m] + ks * 16) * p.pitch_b),
                                                                    (uint32_t)p.pitch_b, p.lbo_b);
              umma_bf16(tmem_base + (uint32_t)(p.mma_acc[m] * p.nt), adesc, bdesc, idesc, (i | ks) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[s]);
          if (i == n_tiles - 1) umma_commit(&done_bar);
        }
        __syncwarp();
      }
    } else {
      // ------------------------------------------------------------ epilogue: TMEM -> this CTA's slab of partial sums
      mbar_wait(&done_bar, 0);
      tc_fence_after();
      const int quad = warp & 3;                       // TMEM lane quadrant this warp may read
      const int row = quad * 32 + lane;                // accumulator row
      const int rep = row / p.m_real;
      float* slab = p.ws + (size_t)blockIdx.x * p.n_acc * 128 * p.nt;
      for (int a = 0; a < p.n_acc; ++a) {
        const bool valid = p.tap_of[tg][a][rep] >= 0;
        for (int c0 = 0; c0 < p.nt; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * p.nt + c0), v);
          tmem_ld_wait();
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(slab + ((size_t)a * 128 + row) * p.nt + c0);
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[c] = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// Sum of the per-CTA slabs over the pixel splits, in a fixed order (deterministic), scattered to OIHW.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ WgradParams p) {
  const int per_group = p.n_acc * 128 * p.nt;
  const int total = p.n_groups * per_group;
  const int tap_groups = p.n_groups / (p.m_blks * p.n_blks);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int group = i / per_group, r = i % per_group;
    const int a = r / (128 * p.nt), row = (r / p.nt) % 128, col = r % p.nt;
    const int tg = group % tap_groups, mn = group / tap_groups;
    const int n_blk = mn % p.n_blks, m_blk = mn / p.n_blks;
    const int tap = p.tap_of[tg][a][row / p.m_real];
    const int co = m_blk * 128 + row % p.m_real, ci = n_blk * p.nt + col;
    if (tap < 0 || co >= p.co || ci >= p.ci_real) continue;
    const size_t stride = (size_t)p.n_groups * per_group;
    const float* src = p.ws + (size_t)group * per_group + r;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int sp = 0;
    for (; sp + 3 < p.n_splits; sp += 4) {                      // four loads in flight, fixed summation order
      a0 += __ldcg(src + (size_t)sp * stride);
      a1 += __ldcg(src + (size_t)(sp + 1) * stride);
      a2 += __ldcg(src + (size_t)(sp + 2) * stride);
      a3 += __ldcg(src + (size_t)(sp + 3) * stride);
    }
    for (; sp < p.n_splits; ++sp) a0 += __ldcg(src + (size_t)sp * stride);
    p.dw[((size_t)co * p.ci_real + ci) * p.taps + tap] = (a0 + a1) + (a2 + a3);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_rows(CUtensorMap* tm, const void* base, int channels, long long rows, int box_c, int box_rows) {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !ptr) {
      set_error("cuTensorMapEncodeTiled entry point not available");
      return 1;
    }
    fn = reinterpret_cast<EncodeFn>(ptr);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)channels, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)channels * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  const CUtensorMapSwizzle sw = box_c * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("wgrad: cuTensorMapEncodeTiled failed (CUresult %d) channels %d rows %lld box %d x %d", (int)r, channels,
              rows, box_c, box_rows);
    return 1;
  }
  return 0;
}

int sm_count() { return device_sm_count(); }

}  // namespace

namespace {
struct WgradShape { int m_real, R, nt, n_acc, per_kh; };
// Channel counts that are not a power of two (HRNet-W48: 48 / 96 / 192 / 384) run on the next larger MMA shape: the TMA
// boxes are 64 channels wide whatever the tensor's channel count, channels beyond it are out of bounds and arrive as
// zeros, so the surplus accumulator rows / columns are zero and simply not written back (wgrad_reduce_kernel).
bool wgrad_shape(int W, int cin, int cout, int cin_real, int k, int stride, WgradShape* o) {
  if (stride != 1 || (k != 1 && k != 3)) return false;
  if (cout % 16 || cin % 16 || cout < 32 || cin < 32 || cin_real > cin) return false;
  o->m_real = cout <= 32 ? 32 : (cout <= 64 ? 64 : 128);       // accumulator rows of one replica of dz
  o->R = 128 / o->m_real;
  // one CTA covers all nine taps when the halo'd x tile fits a TMA box (<= 256 rows); otherwise one filter row per CTA
  o->per_kh = k == 3 && (o->R == 1 || kKT + 2 * (W + 2) > 256);
  if (k == 1) o->n_acc = 1;
  else if (o->per_kh) o->n_acc = o->R >= 3 ? 1 : (o->R == 2 ? 2 : 3);
  else o->n_acc = o->R >= 3 ? 3 : 6;
  o->nt = 0;
  const int want = cin <= 32 ? 32 : (cin <= 64 ? 64 : 128);    // input-channel columns per CTA (zero-filled past cin)
  for (int nt = want; nt >= 32; nt /= 2)
    if (o->n_acc * nt <= 512) { o->nt = nt; break; }
  return o->nt != 0;
}
}  // namespace

bool wgrad_tc_supported(int W, int cin, int cout, int cin_real, int k, int stride) {
  WgradShape sh;
  return wgrad_shape(W, cin, cout, cin_real, k, stride, &sh);
}

namespace {
// Everything but the tensor maps and buffers.
int wgrad_setup(int N, int H, int W, int cin, int cout, int k, int cin_real, WgradParams* pp) {
  WgradShape sh;
  if (!wgrad_shape(W, cin, cout, cin_real, k, 1, &sh)) { set_error("wgrad_tc: unsupported shape"); return 1; }
  WgradParams& p = *pp;
  const int Wp = W + 1;
  const long long P = (long long)N * (H + 1) * Wp;
  p.co = cout; p.ci = cin; p.ci_real = cin_real; p.taps = k * k;
  p.m_real = sh.m_real;
  const int R = sh.R;
  p.m_blks = (cout + 127) / 128;
  p.nt = sh.nt;
  p.n_blks = (cin + p.nt - 1) / p.nt;
  p.n_acc = sh.n_acc;
  p.pitch_a = (cout <= 32 ? 32 : 64) * 2;      // TMA box width in bytes = swizzle span (64 B or 128 B)
  p.pitch_b = (cin <= 32 ? 32 : 64) * 2;
  p.a_panels = p.m_real > 64 ? 2 : 1;
  p.b_panels = p.nt > 64 ? 2 : 1;
  p.dz_rows = kKT + kLead;
  int tap_groups = 1;
  for (int g = 0; g < 3; ++g)
    for (int m = 0; m < kMaxMma; ++m)
      for (int j = 0; j < 4; ++j) p.tap_of[g][m][j] = -1;
  for (int m = 0; m < kMaxMma; ++m) p.mma_acc[m] = m;
  if (k == 1) {
    p.n_mma = 1;
    p.x_rows = kKT;
    p.x_row0[0] = 0;
    p.mma_b[0][0] = 0;
    p.tap_of[0][0][0] = 0;
  } else if (!sh.per_kh && R >= 3) {   // one MMA per filter row: replicas 0,1,2 -> kw = 2,1,0
    p.n_mma = 3;
    p.x_rows = kKT + 2 * (Wp + 1);
    p.x_row0[0] = -(Wp + 1);
    for (int kh = 0; kh < 3; ++kh) {
      p.mma_b[0][kh] = (kh - 1) * Wp + 1 - p.x_row0[0];
      for (int j = 0; j < 3; ++j) p.tap_of[0][kh][j] = kh * 3 + (2 - j);
    }
  } else if (!sh.per_kh) {             // R == 2: two MMAs per filter row: (kw = 2,1) and (kw = 0, unused)
    p.n_mma = 6;
    p.x_rows = kKT + 2 * (Wp + 1);
    p.x_row0[0] = -(Wp + 1);
    for (int kh = 0; kh < 3; ++kh) {
      p.mma_b[0][2 * kh] = (kh - 1) * Wp + 1 - p.x_row0[0];
      p.mma_b[0][2 * kh + 1] = (kh - 1) * Wp - 1 - p.x_row0[0];
      p.tap_of[0][2 * kh][0] = kh * 3 + 2;
      p.tap_of[0][2 * kh][1] = kh * 3 + 1;
      p.tap_of[0][2 * kh + 1][0] = kh * 3 + 0;
    }
  } else {                             // one tap group (CTA) per filter row; x tile = rows [q0 + (kh-1)Wp - 2, +KT+4)
    tap_groups = 3;
    p.x_rows = kKT + 4;
    p.n_mma = p.n_acc;
    for (int kh = 0; kh < 3; ++kh) {
      p.x_row0[kh] = (kh - 1) * Wp - 2;
      if (R >= 3) {
        p.mma_b[kh][0] = 3;
        for (int j = 0; j < 3; ++j) p.tap_of[kh][0][j] = kh * 3 + (2 - j);
      } else if (R == 2) {
        p.mma_b[kh][0] = 3;
        p.mma_b[kh][1] = 1;
        p.tap_of[kh][0][0] = kh * 3 + 2;
        p.tap_of[kh][0][1] = kh * 3 + 1;
        p.tap_of[kh][1][0] = kh * 3 + 0;
      } else {
        for (int kw = 0; kw < 3; ++kw) {
          p.mma_b[kh][kw] = kw + 1;
          p.tap_of[kh][kw][0] = kh * 3 + kw;
        }
      }
    }
  }
  if (p.x_rows > 256) { set_error("wgrad_tc: image too wide for one TMA box (Wp %d)", Wp); return 1; }
  p.n_groups = p.m_blks * p.n_blks * tap_groups;
  p.a_panel_bytes = ((uint32_t)(p.dz_rows * p.pitch_a) + 1023u) & ~1023u;
  p.b_panel_bytes = ((uint32_t)(p.x_rows * p.pitch_b) + 1023u) & ~1023u;
  p.stage_bytes = p.a_panels * p.a_panel_bytes + p.b_panels * p.b_panel_bytes;
  p.lbo_a = R > 1 ? (uint32_t)p.pitch_a : p.a_panel_bytes;
  p.lbo_b = p.b_panel_bytes;
  p.stages = (int)((200u * 1024u) / p.stage_bytes);
  if (p.stages > kMaxStagesW) p.stages = kMaxStagesW;
  if (p.stages < 2) { set_error("wgrad_tc: stage too large"); return 1; }
  int cols = p.n_acc * p.nt;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols *= 2;
  p.tiles_total = (int)((P + kLead + kKT - 1) / kKT);
  p.n_splits = sm_count() / p.n_groups;
  if (p.n_splits < 1) p.n_splits = 1;
  // every split writes a whole slab (n_acc x 128 x nt floats, up to 196 KB) that the reduction reads back: a split should
  // cover enough pixel tiles to be worth its slab (small batches / low-resolution branches would otherwise move far more
  // slab bytes than activations)
  {
    int min_tiles = 8;   // measured: 18.2 -> 17.7 ms per step at batch 32, 41.3 -> 40.7 at batch 128 (1 vs 8)
    if (const char* e = getenv("STL_WGRAD_MIN_TILES")) min_tiles = atoi(e) > 0 ? atoi(e) : 1;
    const int cap = (p.tiles_total + min_tiles - 1) / min_tiles;
    if (p.n_splits > cap) p.n_splits = cap;
  }
  if (p.n_splits > p.tiles_total) p.n_splits = p.tiles_total;
  p.tiles_per_split = (p.tiles_total + p.n_splits - 1) / p.n_splits;
  p.n_splits = (p.tiles_total + p.tiles_per_split - 1) / p.tiles_per_split;
  return 0;
}
size_t slab_bytes(const WgradParams& p) {
  return (size_t)p.n_groups * p.n_splits * p.n_acc * 128 * p.nt * sizeof(float);
}
}  // namespace

size_t wgrad_tc_workspace_bytes(int N, int H, int W, int cin, int cout, int k, int cin_real) {
  WgradParams p{};
  if (wgrad_setup(N, H, W, cin, cout, k, cin_real, &p)) return 0;
  return slab_bytes(p);
}

int wgrad_tc_launch(const __nv_bfloat16* x, const __nv_bfloat16* dz, float* dw, int N, int H, int W, int cin, int cout,
                    int k, int cin_real, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  WgradParams p{};
  if (wgrad_setup(N, H, W, cin, cout, k, cin_real, &p)) return 1;
  if (!workspace || workspace_bytes < slab_bytes(p)) {
    set_error("wgrad_tc: workspace of %zu bytes required (stl_conv_wgrad_workspace_bytes), got %zu", slab_bytes(p),
              workspace_bytes);
    return 1;
  }
  p.dw = dw;
  p.ws = static_cast<float*>(workspace);
  p.dbg = getenv("STL_WGRAD_DBG") ? atoi(getenv("STL_WGRAD_DBG")) : 0;
  const long long P = (long long)N * (H + 1) * (W + 1);
  if (encode_rows(&p.tmDz, dz, cout, P, p.pitch_a / 2, p.dz_rows)) return 1;
  if (encode_rows(&p.tmX, x, cin, P, p.pitch_b / 2, p.x_rows)) return 1;
  cudaError_t e;
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  static DeviceOnce attr_once;
  if (attr_once.run([]() {
        cudaError_t e2 = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
        if (e2 != cudaSuccess) { set_error("wgrad_tc attribute: %s", cudaGetErrorString(e2)); return 1; }
        return 0;
      }))
    return 1;
  wgrad_tc_kernel<<<p.n_groups * p.n_splits, kThreadsW, smem, stream>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("wgrad_tc launch: %s", cudaGetErrorString(e)); return 1; }
  const int total = p.n_groups * p.n_acc * 128 * p.nt;
  wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("wgrad_reduce launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace stl

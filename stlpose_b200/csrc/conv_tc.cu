// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   D[pixels, Cout] = sum over taps t, channel chunks c :  A_t,c[pixels, ck] * W_t,c[Cout, ck]^T
//
// Persistent, warp-specialised CTA (one per SM):
//   warp 0 / lane 0 : TMA producer   (activation ring + weight ring, mbarrier full/empty pairs)
//   warp 1 / lane 0 : tcgen05.mma issuer, fp32 accumulators in TMEM (double-buffered across tiles)
//   warps 2..5      : epilogue - tcgen05.ld, + bias (folded BN) [+ residual] [+ upsampled addends] [ReLU],
//                     bf16 padded-NHWC store or fp32 NCHW store (heatmap head)
//
// Stride-1 convs ("flat" mode) exploit the padded-linear layout (conv.h): a tile is 128*mb consecutive padded
// pixels; the 3x3 taps are the same smem tile read at 9 row offsets, so each activation byte is fetched from
// L2 once per tile (plus halo) instead of 9 times.  The K-major, swizzled smem rows are addressed by UMMA
// descriptors whose start address is advanced by whole rows; the swizzle is a function of absolute smem
// address bits (the same property the usual +32 B K-advance relies on), so any row offset is legal.
// `a_shift = 0` falls back to one aligned TMA load per tap.
// Stride-2 convs ("structured" mode) load each tap with a 4-D tensor map whose W/H traversal stride is 2.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "conv.h"
#include "ptx.cuh"

namespace stl {

namespace {

constexpr int kThreads = 672;          // warp 0: TMA, warp 1: MMA, warps 2..17: two epilogue groups of 8 warps,
                                       // warps 18..19: one epilogue DMA warp per group (panel loads / stores),
                                       // warp 20: second MMA issuer (burst mode: the 128-row blocks of a tile are split)
constexpr int kThreadsKW = 640;         // kw-merged kernels: no second MMA issuer, 96 registers per thread instead of 80
constexpr uint32_t kCtlBytes = 4096;   // [0,1024): barriers + TMEM base ; [1024, ctl_bytes): fp32 bias (ctl_bytes <= 4096)
constexpr uint32_t kBiasOffset = 1024;
constexpr int kMaxCoutPad = 768;
constexpr int kEpiStaged = 0, kEpiDirect = 1, kEpiNchw = 2;
constexpr int kEpiStagedS2 = 3;  // staged epilogue on structured (stride-2) tiles: panels are 4-D TMA boxes (c, w, h, n)
// kw-merged 3x3 (flat mode, resident weights, staged epilogue).  An N <= 64 tcgen05.mma is bound by its A-operand fetch
// from shared memory (128 rows x 32 B in ~45 cycles, whatever N is), so narrow layers run at a third to a half of the
// tensor rate.  Here the three taps of a filter ROW share one instruction: B = the three tap tiles (contiguous in the
// resident weight block) = N 3*Cout, A = the tile shifted by the row offset (kh-1)*Wp only:
//   D_kw[q'] = sum_{kh,ci} X[q' + (kh-1)*Wp][ci] * W[kh][kw][co][ci]      (column group kw of the accumulator)
//   out[q]   = D_0[q-1] + D_1[q] + D_2[q+1]
// The epilogue thread of accumulator row r adds group 0 of row r-2, group 1 of row r-1 and its own group 2 (warp shuffles;
// the two rows below a 32-lane quarter come from the neighbouring warp through shared memory) and so holds output pixel
// q_tile + r - 2: a 128-row block yields 126 outputs, tiles advance by 126 pixels, panels are 126-row TMA boxes.
constexpr int kEpiStagedKW = 4;
constexpr size_t kMaxSmem = 227 * 1024;

struct Ctl {
  uint64_t a_full[kMaxStages], a_empty[kMaxStages];
  uint64_t b_full[kMaxStages], b_empty[kMaxStages];
  uint64_t acc_full[4], acc_empty[4];
  uint64_t res_full[4], epi_free[4], epi_done[4];  // [epilogue group][staging buffer]
  uint32_t tmem_base;
  uint32_t last_cta;   // STATS: set in the CTA that arrives last at the statistics ticket
};
static_assert(sizeof(Ctl) <= kBiasOffset, "control block too large");

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int NC>
__device__ __forceinline__ void add_bf16_regs(float (&f)[NC], const uint4 (&r)[NC / 8]) {
#pragma unroll
  for (int g = 0; g < NC / 8; ++g) {
    f[g * 8 + 0] += bf16_lo(r[g].x); f[g * 8 + 1] += bf16_hi(r[g].x);
    f[g * 8 + 2] += bf16_lo(r[g].y); f[g * 8 + 3] += bf16_hi(r[g].y);
    f[g * 8 + 4] += bf16_lo(r[g].z); f[g * 8 + 5] += bf16_hi(r[g].z);
    f[g * 8 + 6] += bf16_lo(r[g].w); f[g * 8 + 7] += bf16_hi(r[g].w);
  }
}

template <int NC>
__device__ __forceinline__ void load_bf16_row(uint4 (&r)[NC / 8], const __nv_bfloat16* src, bool pred) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
#pragma unroll
  // (ld.global.cg, not the non-coherent path: with programmatic dependent launch this kernel is already resident while
  // the layer that writes these activations is still running)
  for (int g = 0; g < NC / 8; ++g) r[g] = pred ? __ldcg(s + g) : make_uint4(0, 0, 0, 0);
}

struct RowPos {
  bool valid;    // row maps to a pixel of the output tensor
  bool is_pad;   // ... which is one of the zero cells of the padded layout
  int q;         // padded-linear pixel index
  int n, h, w;
};

__device__ __forceinline__ RowPos row_position(const ConvParams& p, long long mt, int r) {
  RowPos pos;
  if (p.mode == 0) {
    const long long q = p.q_lo + mt * p.tile_rows + r;
    pos.valid = q < p.P;
    pos.q = (int)q;
    const uint32_t t = p.fd_Wp.div((uint32_t)pos.q);
    pos.w = pos.q - (int)t * p.Wp;
    pos.n = (int)p.fd_Hp.div(t);
    pos.h = (int)t - pos.n * p.Hp;
    pos.is_pad = (pos.w == p.W) || (pos.h == p.H);
  } else {
    const uint32_t t = p.fd_bw.div((uint32_t)r);
    const int wl = r - (int)t * p.bw;
    const int nl = (int)p.fd_bh.div(t);
    const int hl = (int)t - nl * p.bh;
    pos.w = (int)(mt % p.tiles_w) * p.bw + wl;
    pos.h = (int)((mt / p.tiles_w) % p.tiles_h) * p.bh + hl;
    pos.n = (int)(mt / ((long long)p.tiles_w * p.tiles_h)) * p.bn + nl;
    pos.valid = pos.w < p.W && pos.h < p.H && pos.n < p.N;
    pos.is_pad = false;
    pos.q = (pos.n * p.Hp + pos.h) * p.Wp + pos.w;
  }
  return pos;
}

// Epilogue parameters copied into registers once per kernel: the tcgen05/TMA asm statements carry memory
// clobbers, so anything read through `p.` inside the unit loop would be re-fetched from the constant bank
// every iteration.
struct EpiParams {
  void* out;
  const __nv_bfloat16* residual;
  int cout, relu, n_up, H, W, skip;
};

__device__ __forceinline__ uint32_t pack_bf16_relu(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));  // first source -> upper half
  return r;
}

// Epilogue for NC consecutive output channels of one pixel; `res` holds the (pre-fetched) residual values and
// `bias` the per-channel bias (fetched from shared memory while the TMEM load was in flight).
template <int NC, bool NCHW>
__device__ __forceinline__ void epilogue_store(const ConvParams& p, const EpiParams& e, uint32_t (&v)[NC], int ch0,
                                               const RowPos& r, const uint4 (&res)[NC / 8],
                                               const float4 (&bias)[NC / 4]) {
  if (!r.valid || e.skip == 1) return;
  float* f = reinterpret_cast<float*>(v);
  if constexpr (NCHW) {
    if (r.is_pad) return;
    float* out = reinterpret_cast<float*>(e.out) + (((size_t)r.n * e.cout + ch0) * e.H + r.h) * e.W + r.w;
    const size_t plane = (size_t)e.H * e.W;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      if (ch0 + i < e.cout) {
        const float b = i % 4 == 0 ? bias[i / 4].x : i % 4 == 1 ? bias[i / 4].y : i % 4 == 2 ? bias[i / 4].z : bias[i / 4].w;
        float t = f[i] + b;
        if (e.relu) t = fmaxf(t, 0.f);
        out[i * plane] = t;
      }
    }
  } else {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out) + (size_t)r.q * e.cout + ch0);
    if (r.is_pad) {
#pragma unroll
      for (int g = 0; g < NC / 8; ++g) o[g] = make_uint4(0, 0, 0, 0);
      return;
    }
#pragma unroll
    for (int g = 0; g < NC / 4; ++g) {
      f[g * 4 + 0] += bias[g].x; f[g * 4 + 1] += bias[g].y; f[g * 4 + 2] += bias[g].z; f[g * 4 + 3] += bias[g].w;
    }
    if (e.residual) {
#pragma unroll
      for (int g = 0; g < NC / 8; ++g) {
        f[g * 8 + 0] += bf16_lo(res[g].x); f[g * 8 + 1] += bf16_hi(res[g].x);
        f[g * 8 + 2] += bf16_lo(res[g].y); f[g * 8 + 3] += bf16_hi(res[g].y);
        f[g * 8 + 4] += bf16_lo(res[g].z); f[g * 8 + 5] += bf16_hi(res[g].z);
        f[g * 8 + 6] += bf16_lo(res[g].w); f[g * 8 + 7] += bf16_hi(res[g].w);
      }
    }
    if (e.n_up) {
#pragma unroll 1
      for (int u = 0; u < e.n_up; ++u) {
        const int s = p.up_shift[u];
        const int hs = e.H >> s, ws = e.W >> s;
        const size_t qs = ((size_t)r.n * (hs + 1) + (r.h >> s)) * (ws + 1) + (r.w >> s);
        const uint4* src = reinterpret_cast<const uint4*>(p.up_src[u] + qs * e.cout + ch0);
#pragma unroll
        for (int g = 0; g < NC / 8; ++g) {
          const uint4 t = __ldcg(src + g);
          f[g * 8 + 0] += bf16_lo(t.x); f[g * 8 + 1] += bf16_hi(t.x);
          f[g * 8 + 2] += bf16_lo(t.y); f[g * 8 + 3] += bf16_hi(t.y);
          f[g * 8 + 4] += bf16_lo(t.z); f[g * 8 + 5] += bf16_hi(t.z);
          f[g * 8 + 6] += bf16_lo(t.w); f[g * 8 + 7] += bf16_hi(t.w);
        }
      }
    }
    if (e.skip == 2) {  // measurement: keep the loads and the math alive, drop (almost) all stores
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < NC; ++i) acc += f[i];
      if (acc != 12345.678f) return;
    }
    if (e.relu) {
#pragma unroll
      for (int g = 0; g < NC / 8; ++g)
        o[g] = make_uint4(pack_bf16_relu(f[g * 8 + 0], f[g * 8 + 1]), pack_bf16_relu(f[g * 8 + 2], f[g * 8 + 3]),
                          pack_bf16_relu(f[g * 8 + 4], f[g * 8 + 5]), pack_bf16_relu(f[g * 8 + 6], f[g * 8 + 7]));
    } else {
#pragma unroll
      for (int g = 0; g < NC / 8; ++g)
        o[g] = make_uint4(pack_bf16(f[g * 8 + 0], f[g * 8 + 1]), pack_bf16(f[g * 8 + 2], f[g * 8 + 3]),
                          pack_bf16(f[g * 8 + 4], f[g * 8 + 5]), pack_bf16(f[g * 8 + 6], f[g * 8 + 7]));
    }
  }
}

// Column sums over a warp: every lane holds 16 values (one row, 16 channels); returns, in lane l, the sum over the 32
// lanes of channel (l >> 1) & 15 (both lanes of a pair hold the same total).  Recursive halving: 16 shuffles, fixed order.
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float keep = h4 ? v[i + 8] : v[i], send = h4 ? v[i] : v[i + 8];
    a[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = h3 ? a[i + 4] : a[i], send = h3 ? a[i] : a[i + 4];
    b[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = h2 ? b[i + 2] : b[i], send = h2 ? b[i] : b[i + 2];
    c[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 4);
  }
  const float keep = h1 ? c[1] : c[0], send = h1 ? c[0] : c[1];
  const float d = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 2);
  return d + __shfl_xor_sync(0xFFFFFFFFu, d, 1);
}
constexpr int kStatSlots = 8;   // (n-tile, panel, slice pair) combinations one epilogue warp can meet

// MB    : 128-row accumulator blocks per tile (compile time so that the MMA issue loop fully unrolls)
// KSTEPS: UMMA K-steps (16 channels each) per K chunk; the smem row / swizzle span is 32*KSTEPS bytes.  The LAST chunk
//         may hold fewer valid channels (HRNet-W48: 48 = 64 - 16, 96 = 64 + 32): its TMA box reaches past the tensor's
//         channel count, the surplus arrives as zeros and only p.ksteps_last K steps are issued for it
// TAPS  : 1 (1x1) or 9 (3x3)
// EPI   : which epilogue the kernel contains (each kernel carries exactly one, for instruction-cache footprint):
//         kEpiStaged - flat mode, bf16 out: panels staged in shared memory, moved by TMA (the common case)
//         kEpiDirect - per-thread global loads/stores: stride-2 convs, upsampled addends, odd channel counts
//         kEpiNchw   - fp32 NCHW store of the heatmap head
// Code size matters here: ten warps run four different roles out of one instruction cache, so everything that
// does not have to be unrolled is a rolled loop and every epilogue path exists exactly once per kernel.
// STATS : (staged epilogues) also accumulate per-channel sum / sum of squares of the stored outputs (training, conv.h)
template <int MB, int KSTEPS, int TAPS, int EPI, bool PAIR, bool STATS = false>
__global__ void __launch_bounds__(EPI == kEpiStagedKW ? kThreadsKW : kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: swizzle-128B atoms (8 rows x 128 B) must start on their natural boundary.
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  Ctl* ctl = reinterpret_cast<Ctl*>(smem);
  float* sbias = reinterpret_cast<float*>(smem + kBiasOffset);
  const uint32_t a_base = smem_u32(smem + p.ctl_bytes);
  const uint32_t b_base = a_base + (uint32_t)p.a_stages * p.a_stage_bytes;

  constexpr bool NCHW = EPI == kEpiNchw;
  constexpr bool KWM = EPI == kEpiStagedKW;                         // kw-merged 3x3 (MB = 1, TAPS = 9, never PAIR)
  constexpr bool kFlatOnly = EPI == kEpiStaged || EPI == kEpiNchw || KWM;  // these kernels are only ever launched in flat mode
  constexpr bool kStruct = EPI == kEpiStagedS2;                     // ... and this one only on structured tiles
  constexpr bool kStagedEpi = EPI == kEpiStaged || EPI == kEpiStagedS2 || KWM;
  constexpr int kPanelRows = KWM ? 126 : 128;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kSpan = 32u * KSTEPS;
  const uint32_t acc_cols = (uint32_t)((KWM ? 3 : MB) * p.nt);
  // weights resident + (shifted-descriptor taps or 1x1): the whole K loop of a chunk is one straight MMA burst
  const bool burst = p.b_resident && (p.a_shift || TAPS == 1);
  // PAIR: two CTAs (one cluster) share every tcgen05.mma: cta_group::2, M = 256 = 128 rows of each CTA, each CTA
  // holds half of the weight rows.  The leader issues for both, which halves the per-SM instruction issue load
  // (the bottleneck for N <= 64) and the weight bytes each CTA keeps resident.  Only used in burst mode.
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const long long tile0 = PAIR ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
  const long long tstride = PAIR ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;
  auto mtile = [&](long long t) { return PAIR ? 2 * (t / p.n_ntiles) + cta_rank : t / p.n_ntiles; };

  // Programmatic dependent launch: let the next layer's CTAs take over SMs as ours retire and run their prologue
  // (barriers, TMEM, resident weights); everything that touches activations waits for the previous layer below.
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(&ctl->a_full[i], 1); mbar_init(&ctl->a_empty[i], p.n_mma); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(&ctl->b_full[i], 1); mbar_init(&ctl->b_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&ctl->acc_full[i], p.n_mma); mbar_init(&ctl->acc_empty[i], PAIR ? 16 : 8); }
    for (int i = 0; i < 4; ++i) { mbar_init(&ctl->res_full[i], 1); mbar_init(&ctl->epi_free[i], 1); mbar_init(&ctl->epi_done[i], 8); }
    fence_mbar_init();
    tma_prefetch_desc(&p.tmA);
    if (p.n_chunks_a < p.n_chunks) tma_prefetch_desc(&p.tmA2);
    tma_prefetch_desc(&p.tmB);
    if (p.epi_tma) { tma_prefetch_desc(&p.tmO); if (p.residual) tma_prefetch_desc(&p.tmR); }
  }
  for (int i = threadIdx.x; i < p.cout_pad; i += (int)blockDim.x) sbias[i] = p.bias[i];
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(&ctl->tmem_base, p.tmem_cols); else tmem_alloc(&ctl->tmem_base, p.tmem_cols);
  }
  tc_fence_before();
  if (PAIR) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  if (PAIR && p.b_resident) {
    // resident weights, this CTA's half of the output-channel rows of every (n-tile, chunk, tap) tile
    if (threadIdx.x == 0) {
      mbar_expect_tx(&ctl->b_full[0], p.b_resident_bytes);
      uint32_t dst = b_base;
      for (int nti = 0; nti < p.n_ntiles; ++nti)
        for (int chunk = 0; chunk < p.n_chunks; ++chunk)
          for (int tap = 0; tap < TAPS; ++tap, dst += p.b_stage_bytes)
            tma_load_3d_s(dst, &p.tmB, &ctl->b_full[0], chunk * p.ck, nti * p.nt + (int)cta_rank * (p.nt / 2), tap);
    }
    mbar_wait(&ctl->b_full[0], 0);
    cluster_sync();  // both halves are in place before the leader's first MMA reads them
  }
  if (warp != 0) pdl_wait();   // (the producer warp first queues the resident weights, which no kernel produces)

#ifdef STL_CONV_COUNTERS
  long long dbg[4] = {0, 0, 0, 0};
  long long dbg_t = 0;
#endif
#ifdef STL_CONV_COUNTERS
#define DBG_TICK() do { if (p.dbg_counters) dbg_t = clock64(); } while (0)
#define DBG_TOCK(i) do { if (p.dbg_counters) { const long long t_ = clock64(); dbg[i] += t_ - dbg_t; dbg_t = t_; } } while (0)
#define DBG_DUMP(role) do { if (p.dbg_counters && lane == 0) for (int i_ = 0; i_ < 4; ++i_) p.dbg_counters[(blockIdx.x * 3 + (role)) * 4 + i_] = dbg[i_]; } while (0)
#else
#define DBG_TICK() do { } while (0)
#define DBG_TOCK(i) do { } while (0)
#define DBG_DUMP(role) do { } while (0)
#endif

  // STATS: per-lane accumulators of this warp's (n-tile, panel, slice pair) slots, see the staged epilogue
  float st_sum[STATS ? kStatSlots : 1], st_sq[STATS ? kStatSlots : 1];
#pragma unroll
  for (int i = 0; i < (STATS ? kStatSlots : 1); ++i) { st_sum[i] = 0.f; st_sq[i] = 0.f; }

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp loops, one lane issues)
    uint32_t a_it = 0, b_it = 0;
    DBG_TICK();
    if (!PAIR && p.b_resident && elect_one()) {  // (PAIR: done before the role split)
      mbar_expect_tx(&ctl->b_full[0], p.b_resident_bytes);
      uint32_t dst = b_base;
      for (int nti = 0; nti < p.n_ntiles; ++nti)
        for (int chunk = 0; chunk < p.n_chunks; ++chunk)
          for (int tap = 0; tap < TAPS; ++tap, dst += p.b_stage_bytes)
            tma_load_3d_s(dst, &p.tmB, &ctl->b_full[0], chunk * p.ck, nti * p.nt, tap);
    }
    pdl_wait();
    __syncwarp();
#pragma unroll 1
    for (long long tile = tile0; tile < p.total_tiles; tile += tstride) {
      const int nti = (int)(tile % p.n_ntiles);
      const long long mt = mtile(tile);
      int q0 = 0, wo0 = 0, ho0 = 0, n0 = 0;
      if (kFlatOnly || (!kStruct && p.mode == 0)) {
        q0 = p.q_lo + (int)(mt * p.tile_rows) - (KWM ? 1 : 0);  // KWM: accumulator row r <-> pixel q0 + r, output q0 + r - 1
      } else {
        wo0 = (int)(mt % p.tiles_w) * p.bw;
        ho0 = (int)((mt / p.tiles_w) % p.tiles_h) * p.bh;
        n0 = (int)(mt / ((long long)p.tiles_w * p.tiles_h)) * p.bn;
      }
#pragma unroll 1
      for (int chunk = 0; chunk < p.n_chunks; ++chunk) {
        int kh = TAPS == 9 ? 0 : 1, kw = TAPS == 9 ? 0 : 1;
#pragma unroll 1
        for (int tap = 0; tap < TAPS; ++tap) {
          if (!p.a_shift || tap == 0) {
            const uint32_t s = a_it % p.a_stages, ph = (a_it / p.a_stages) & 1;
            mbar_wait(&ctl->a_empty[s], ph ^ 1);
            DBG_TOCK(0);
            if (elect_one()) {
              // PAIR: both CTAs' tiles complete on the leader's barrier, which expects the bytes of both
              if (!PAIR) mbar_expect_tx(&ctl->a_full[s], p.a_tx_bytes);
              else if (is_leader) mbar_expect_tx(&ctl->a_full[s], 2u * p.a_tx_bytes);
              const uint32_t dst = a_base + s * p.a_stage_bytes;
              if (kFlatOnly || (!kStruct && p.mode == 0)) {
                const int row0 = p.a_shift ? q0 - p.halo : q0 + (kh - 1) * p.in_Wp + (kw - 1);
                // (two-input 1x1 convolutions: the K chunks past the first tensor's channels come from the second)
                const bool second = chunk >= p.n_chunks_a;
                const CUtensorMap* tm = second ? &p.tmA2 : &p.tmA;
                const int c0 = (second ? chunk - p.n_chunks_a : chunk) * p.ck;
                for (int i = 0; i < p.a_pieces; ++i) {
                  if (PAIR)
                    tma_load_2d_pair(dst + (uint32_t)(i * p.a_box_rows) * kSpan, tm, &ctl->a_full[s], c0,
                                     row0 + i * p.a_box_rows);
                  else
                    tma_load_2d_s(dst + (uint32_t)(i * p.a_box_rows) * kSpan, tm, &ctl->a_full[s], c0,
                                  row0 + i * p.a_box_rows);
                }
              } else {
                tma_load_4d_s(dst, &p.tmA, &ctl->a_full[s], chunk * p.ck, 2 * wo0 + kw - 1, 2 * ho0 + kh - 1, n0);
              }
            }
            __syncwarp();
            ++a_it;
            DBG_TOCK(2);
          }
          if (!p.b_resident && (p.b_taps == 1 || kw == 0)) {    // (b_taps == 3: one stage per filter row)
            const uint32_t s = b_it % p.b_stages, ph = (b_it / p.b_stages) & 1;
            mbar_wait(&ctl->b_empty[s], ph ^ 1);
            DBG_TOCK(1);
            if (elect_one()) {
              if (PAIR) {  // each CTA streams its half of the weight rows; both halves complete on the leader's barrier
                if (is_leader) mbar_expect_tx(&ctl->b_full[s], 2u * p.b_tx_bytes);
                tma_load_3d_pair(b_base + s * p.b_stage_bytes, &p.tmB, &ctl->b_full[s], chunk * p.ck,
                                 nti * p.nt + (int)cta_rank * (p.nt / 2), tap);
              } else {
                mbar_expect_tx(&ctl->b_full[s], p.b_tx_bytes);
                tma_load_3d_s(b_base + s * p.b_stage_bytes, &p.tmB, &ctl->b_full[s], chunk * p.ck, nti * p.nt, tap);
              }
            }
            __syncwarp();
            ++b_it;
            DBG_TOCK(2);
          }
          if (++kw == 3) { kw = 0; ++kh; }
        }
      }
    }
    DBG_DUMP(0);
  } else if (warp == 1 || warp == 20) {
    // ------------------------------------------------------------------ MMA issuer(s) (whole warp loops, one lane issues).
    // A single thread sustains one tcgen05.mma per ~45 cycles, more than the math of an N <= 64 instruction, so in
    // burst mode a second warp issues the upper 128-row blocks of every tile (disjoint accumulators).
    const int part = warp == 20 ? 1 : 0;
    const int m_split = p.n_mma == 2 ? (MB + 1) / 2 : MB;
    const int m_lo = part == 0 ? 0 : m_split, m_hi = part == 0 ? m_split : MB;
    if ((part == 1 && p.n_mma != 2) || (PAIR && !is_leader)) goto role_done;
    {
    const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, (uint32_t)(KWM ? 3 * p.nt : p.nt));
    const uint64_t desc_hi = make_kmajor_desc(0, kSpan) & 0xFFFFFFFF00000000ull;
    const uint32_t desc_lo_flags = (uint32_t)(make_kmajor_desc(0, kSpan) & 0xFFFFFFFFull);
    // descriptors differ only in the 14-bit start-address field (16-byte units) of the low word
    const uint32_t a_lo_base = desc_lo_flags | (a_base >> 4);
    const uint32_t b_lo_base = desc_lo_flags | (b_base >> 4);
    const uint32_t a_stage_u = p.a_stage_bytes >> 4, b_stage_u = p.b_stage_bytes >> 4;
    const uint32_t wp_u = (uint32_t)p.in_Wp * (kSpan >> 4);  // one padded image row, in 16-byte units
    const uint32_t nt = (uint32_t)p.nt;
    uint32_t a_it = 0, b_it = 0, acc_it = 0;
    DBG_TICK();
    if (!PAIR && p.b_resident) {
      mbar_wait(&ctl->b_full[0], 0);
      DBG_TOCK(2);
    }
#pragma unroll 1
    for (long long tile = tile0; tile < p.total_tiles; tile += tstride) {
      const uint32_t nti = (uint32_t)(tile % p.n_ntiles);
      const uint32_t buf = acc_it % p.n_accbuf, aph = (acc_it / p.n_accbuf) & 1;
      mbar_wait(&ctl->acc_empty[buf], aph ^ 1);
      DBG_TOCK(0);
      tc_fence_after();
      const uint32_t d_base = tmem_base + buf * acc_cols;
      if (burst) {
#pragma unroll 1
        for (int chunk = 0; chunk < p.n_chunks; ++chunk) {
          const uint32_t a_stage = a_it % p.a_stages;
          mbar_wait(&ctl->a_full[a_stage], (a_it / p.a_stages) & 1);
          DBG_TOCK(1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo0 = a_lo_base + a_stage * a_stage_u;
            const uint32_t b_lo0 = b_lo_base + ((nti * p.n_chunks + chunk) * TAPS) * b_stage_u;
            const int kmax = chunk == p.n_chunks - 1 ? p.ksteps_last : KSTEPS;   // K steps of this chunk that hold data
            if constexpr (KWM) {
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  const uint64_t da = desc_hi | (uint64_t)(a_lo0 + (uint32_t)kh * wp_u + (uint32_t)((k * 32) >> 4));
                  const uint64_t db = desc_hi | (uint64_t)(b_lo0 + (uint32_t)(3 * kh) * b_stage_u + (uint32_t)((k * 32) >> 4));
                  umma_bf16(d_base, da, db, idesc, (kh | k) != 0 ? 1u : (uint32_t)chunk);
                }
              }
            } else
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
              const uint32_t a_lo = a_lo0 + (TAPS == 9 ? (uint32_t)(tap / 3) * wp_u + (uint32_t)(tap % 3) * (kSpan >> 4) : 0u);
              const uint32_t b_lo = b_lo0 + (uint32_t)tap * b_stage_u;
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                if (k >= kmax) continue;
#pragma unroll
                for (int m = 0; m < MB; ++m) {
                  if (m < m_lo || m >= m_hi) continue;
                  const uint64_t da = desc_hi | (uint64_t)(a_lo + (uint32_t)((m * 128 * (int)kSpan + k * 32) >> 4));
                  const uint64_t db = desc_hi | (uint64_t)(b_lo + (uint32_t)((k * 32) >> 4));
                  if (PAIR) umma_bf16_pair(d_base + (uint32_t)m * nt, da, db, idesc, (tap | k) != 0 ? 1u : (uint32_t)chunk);
                  else umma_bf16(d_base + (uint32_t)m * nt, da, db, idesc, (tap | k) != 0 ? 1u : (uint32_t)chunk);
                }
              }
            }
            if (PAIR) umma_commit_pair(&ctl->a_empty[a_stage]); else umma_commit(&ctl->a_empty[a_stage]);
          }
          __syncwarp();
          ++a_it;
          DBG_TOCK(3);
        }
      } else {
        uint32_t a_stage = 0;
#pragma unroll 1
        for (int chunk = 0; chunk < p.n_chunks; ++chunk) {
          uint32_t tap_u = 0;  // row offset of the current tap inside the halo'd tile, 16-byte units
          int kw = 0;
#pragma unroll 1
          for (int tap = 0; tap < TAPS; ++tap) {
            if (!p.a_shift || tap == 0) {
              a_stage = a_it % p.a_stages;
              mbar_wait(&ctl->a_full[a_stage], (a_it / p.a_stages) & 1);
              DBG_TOCK(1);
            }
            uint32_t bs;
            const bool b_first = p.b_taps == 1 || kw == 0, b_last = p.b_taps == 1 || kw == 2;   // of this weight stage
            if (p.b_resident) {
              bs = (nti * p.n_chunks + chunk) * TAPS + tap;
            } else {
              bs = b_it % p.b_stages;
              if (b_first) {
                mbar_wait(&ctl->b_full[bs], (b_it / p.b_stages) & 1);
                DBG_TOCK(2);
              }
            }
            tc_fence_after();
            const bool release_a = !p.a_shift || tap == TAPS - 1;
            if (elect_one()) {
              const uint32_t a_lo = a_lo_base + a_stage * a_stage_u + (p.a_shift ? tap_u : 0u);
              const uint32_t b_lo = b_lo_base + bs * b_stage_u + (p.b_taps == 3 ? (uint32_t)kw * (b_stage_u / 3u) : 0u);
              const uint32_t first = (uint32_t)(chunk | tap);
              const int kmax = chunk == p.n_chunks - 1 ? p.ksteps_last : KSTEPS;
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                if (k >= kmax) continue;
#pragma unroll
                for (int m = 0; m < MB; ++m) {
                  const uint64_t da = desc_hi | (uint64_t)(a_lo + (uint32_t)((m * 128 * (int)kSpan + k * 32) >> 4));
                  const uint64_t db = desc_hi | (uint64_t)(b_lo + (uint32_t)((k * 32) >> 4));
                  if (PAIR) umma_bf16_pair(d_base + (uint32_t)m * nt, da, db, idesc, first | (uint32_t)k);
                  else umma_bf16(d_base + (uint32_t)m * nt, da, db, idesc, first | (uint32_t)k);
                }
              }
              if (PAIR) {
                if (!p.b_resident && b_last) umma_commit_pair(&ctl->b_empty[bs]);
                if (release_a) umma_commit_pair(&ctl->a_empty[a_stage]);
              } else {
                if (!p.b_resident && b_last) umma_commit(&ctl->b_empty[bs]);
                if (release_a) umma_commit(&ctl->a_empty[a_stage]);
              }
            }
            __syncwarp();
            if (!p.b_resident && b_last) ++b_it;
            if (release_a) ++a_it;
            tap_u += kSpan >> 4;
            if (++kw == 3) { kw = 0; tap_u += wp_u - 3 * (kSpan >> 4); }
            DBG_TOCK(3);
          }
        }
      }
      if (elect_one()) {
        if (PAIR) umma_commit_pair(&ctl->acc_full[buf]); else umma_commit(&ctl->acc_full[buf]);
      }
      __syncwarp();
      ++acc_it;
    }
    if (part == 0) DBG_DUMP(1);
    }
  } else if (warp >= 18) {
    // ------------------------------------------------------------------ epilogue DMA warp of one group: moves staged
    // panels between shared and global memory with TMA so that the 8 compute warps never wait on an issue slot
    const int group = warp - 18;
    if (kStagedEpi && (p.n_accbuf >= 2 || group == 0)) {
      const int nt = p.nt, n_ntiles = p.n_ntiles;
      const int panel_ch = p.panel_ch, npanels = nt / panel_ch;
      const int PT = MB * npanels, bp = p.epi_batch;
      const uint32_t panel_bytes = p.epi_panel_bytes, pitch = (uint32_t)panel_ch * 2u;
      const uint32_t stage0 = a_base + p.epi_base_off + (uint32_t)group * 2u * (uint32_t)bp * panel_bytes;
      const bool has_res = p.residual != nullptr;
      const long long total_tiles = p.total_tiles;
      const long long tile_step = tstride * (p.n_accbuf >= 2 ? 2 : 1);
      const long long first_tile = tile0 + (p.n_accbuf >= 2 ? (long long)group * tstride : 0);
      const int skip = p.dbg_skip_epilogue;
      if (lane == 0) {
        // residual of batch (tile, b0) -> staging buffer sb
        auto load_res = [&](long long t, int b0, uint32_t sb) {
          const int nti_ = (int)(t % n_ntiles);
          const int q0 = p.q_lo + (int)(mtile(t) * p.tile_rows);
          const int cnt = PT - b0 < bp ? PT - b0 : bp;
          uint64_t* bar = &ctl->res_full[group * 2 + sb];
          mbar_expect_tx(bar, (uint32_t)cnt * (uint32_t)kPanelRows * pitch);
          const long long mt_ = mtile(t);
          for (int i = 0; i < cnt; ++i) {
            const int idx = b0 + i, m = idx / npanels, pn = idx - m * npanels;
            const uint32_t dst = stage0 + (sb * (uint32_t)bp + (uint32_t)i) * panel_bytes;
            if (kStruct)   // block m of a structured tile = images [n0 + m*bn1, +bn1) of the (bw x bh) window
              tma_load_4d_s(dst, &p.tmR, bar, nti_ * nt + pn * panel_ch, (int)(mt_ % p.tiles_w) * p.bw,
                            (int)((mt_ / p.tiles_w) % p.tiles_h) * p.bh,
                            (int)(mt_ / ((long long)p.tiles_w * p.tiles_h)) * p.bn + m * (p.bn / MB));
            else
              tma_load_2d_s(dst, &p.tmR, bar, nti_ * nt + pn * panel_ch, q0 + m * 128);
          }
        };
        uint32_t kb = 0;
        if (has_res && first_tile < total_tiles) load_res(first_tile, 0, 0);
#pragma unroll 1
        for (long long tile = first_tile; tile < total_tiles; tile += tile_step) {
          const int nti = (int)(tile % n_ntiles);
          const int q0 = p.q_lo + (int)(mtile(tile) * p.tile_rows);
#pragma unroll 1
          for (int b0 = 0; b0 < PT; b0 += bp, ++kb) {
            const uint32_t sb = kb & 1;
            // the other buffer was last read by the stores of batch kb-1: once they have drained it, it can take the
            // next batch (its residual, or just become writable)
            bulk_wait_read<0>();
            {
              long long t2 = tile;
              int b2 = b0 + bp;
              if (b2 >= PT) { b2 = 0; t2 += tile_step; }
              if (t2 < total_tiles) {
                if (has_res) load_res(t2, b2, sb ^ 1);
                else if (kb >= 1) mbar_arrive(&ctl->epi_free[group * 2 + (sb ^ 1)]);
              }
            }
            // all 8 compute warps have written (and fenced) their cells of batch kb
            mbar_wait(&ctl->epi_done[group * 2 + sb], (kb >> 1) & 1);
            if (skip != 1) {
              const int cnt = PT - b0 < bp ? PT - b0 : bp;
              const long long mt_ = mtile(tile);
              for (int i = 0; i < cnt; ++i) {
                const int idx = b0 + i, m = idx / npanels, pn = idx - m * npanels;
                const uint32_t src = stage0 + (sb * (uint32_t)bp + (uint32_t)i) * panel_bytes;
                if (kStruct)
                  tma_store_4d_s(&p.tmO, src, nti * nt + pn * panel_ch, (int)(mt_ % p.tiles_w) * p.bw,
                                 (int)((mt_ / p.tiles_w) % p.tiles_h) * p.bh,
                                 (int)(mt_ / ((long long)p.tiles_w * p.tiles_h)) * p.bn + m * (p.bn / MB));
                else
                  tma_store_2d_s(&p.tmO, src, nti * nt + pn * panel_ch, q0 + m * 128);
              }
              bulk_commit();
            }
          }
        }
        bulk_wait<0>();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: two groups of 8 warps, one group
    // per accumulator buffer (the TMEM drain of tile t overlaps the MMAs of tile t+1 and the drain of t+1).
    // Inside a group every TMEM lane quarter has two warps; a work unit is one 128-row block x 16 channels and
    // the two warps take alternate 16-channel slices.  Many warps with short dependent chains: the drain is
    // latency-bound per warp, so throughput comes from warp-level parallelism.
    const int e_idx = warp - 2;
    const int group = e_idx >> 3;
    const int sub = (e_idx >> 2) & 1;
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are the ones this warp may read
    const int nt = p.nt;
    const int nslices = nt >> 4;
    const int n_ntiles = p.n_ntiles, n_accbuf = p.n_accbuf;
    EpiParams e;
    e.out = p.out; e.residual = NCHW ? nullptr : p.residual; e.cout = p.cout; e.relu = p.relu; e.n_up = p.n_up;
    e.H = p.H; e.W = p.W; e.skip = p.dbg_skip_epilogue;
    const bool has_res = e.residual != nullptr;
    const long long total_tiles = p.total_tiles;
    const int row0 = quarter * 32 + lane;
    uint32_t acc_it = 0;
    DBG_TICK();
    if constexpr (KWM) {
      // ---- kw-merged staged epilogue (one 126-row panel per tile, one tile per batch).  The tensor pipe needs only
      // ~1000 cycles per tile here, so this loop is written for instruction count: no generic tile / panel index math,
      // packed fp32x2 adds, the zero-cell test once per tile.
      const int spp = nt >> 4;                              // 16-channel slices of the panel (panel_ch == nt)
      const uint32_t pitch = (uint32_t)nt * 2u, swz = (uint32_t)p.panel_swz;
      const uint32_t panel_bytes = p.epi_panel_bytes;
      const uint32_t stage0 = a_base + p.epi_base_off + (uint32_t)group * 2u * panel_bytes;
      const int prow = row0 - 2;                            // panel row = tile-local output pixel (rows 0, 1: none)
      const uint32_t xr = swz == 128 ? (uint32_t)(prow & 7) : (swz == 64 ? (uint32_t)((prow >> 1) & 3) : 0u);
      const uint32_t row_off = (uint32_t)prow * pitch;
      // exchange area [group][sub][parity][quarter][3 rows x 16 fp32]: rows = group 0 of lanes 30, 31, group 1 of lane 31
      const uint32_t xch = a_base + p.xch_off + (uint32_t)((group * 2 + sub) * 2) * (4u * 192u) + (uint32_t)quarter * 192u;
      const uint32_t bar_id = 1u + (uint32_t)(group * 2 + sub);
      const bool x_rd = quarter > 0 && lane < 2;
      const int Wp = p.Wp, Hp = p.Hp;
      const FastDiv fdW = p.fd_Wp, fdH = p.fd_Hp;
      const int ntiles = (int)total_tiles, tstep = 2 * (int)tstride;
      int q = p.q_lo + ((int)tile0 + group * (int)tstride) * 126 + prow;   // this thread's output pixel
      const int qstep = tstep * 126;
      uint32_t kb = 0, xk = 0;
      acc_it = (uint32_t)group;
#pragma unroll 1
      for (int tile = (int)tile0 + group * (int)tstride; tile < ntiles; tile += tstep, acc_it += 2, ++kb, q += qstep) {
        const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
        const uint32_t t_tile = tmem_base + buf * acc_cols + ((uint32_t)(quarter * 32) << 16);
        const uint32_t sbuf = kb & 1u;
        const uint32_t base = stage0 + sbuf * panel_bytes + row_off;
        bool is_pad = false;
        if (prow >= 0) {
          const uint32_t t = fdW.div((uint32_t)q);
          const int w = q - (int)t * Wp;
          const int h = (int)t - (int)fdH.div(t) * Hp;
          is_pad = (w == e.W) || (h == e.H);
        }
        DBG_TOCK(1);
        mbar_wait(&ctl->acc_full[buf], aph);
        DBG_TOCK(0);
        tc_fence_after();
        bool ready = false;
#pragma unroll 1
        for (int sl = sub; sl < spp; sl += 2) {
          uint32_t v0[16], v1[16], v[16];
          tmem_ld16(t_tile + (uint32_t)(sl * 16), v0);
          tmem_ld16(t_tile + (uint32_t)(nt + sl * 16), v1);
          tmem_ld16(t_tile + (uint32_t)(2 * nt + sl * 16), v);
          const uint32_t c0 = (uint32_t)sl * 2u;
          const uint32_t ad0 = base + ((c0 ^ xr) << 4), ad1 = base + (((c0 + 1) ^ xr) << 4);
          float4 bias[4];
          const float4* b4 = reinterpret_cast<const float4*>(sbias + sl * 16);
#pragma unroll
          for (int g = 0; g < 4; ++g) bias[g] = b4[g];
          if (!ready) {
            if (has_res) mbar_wait(&ctl->res_full[group * 2 + sbuf], (kb >> 1) & 1);
            else mbar_wait(&ctl->epi_free[group * 2 + sbuf], ((kb >> 1) & 1) ^ 1);
            ready = true;
            DBG_TOCK(3);
          }
          uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
          if (has_res && prow >= 0) { r0 = lds128(ad0); r1 = lds128(ad1); }
          tmem_ld_wait();
          // out[row - 2] = D0[row - 2] + D1[row - 1] + D2[row]: the rows below this warp's 32 lanes come from the
          // neighbouring quarter's warp (same group, same slice sequence) through the exchange area
          const uint32_t xq = xch + (xk & 1u) * (4u * 192u);
          ++xk;
          if (lane >= 30) {
            const uint32_t a = xq + (uint32_t)(lane - 30) * 64u;
            sts128(a, make_uint4(v0[0], v0[1], v0[2], v0[3]));
            sts128(a + 16, make_uint4(v0[4], v0[5], v0[6], v0[7]));
            sts128(a + 32, make_uint4(v0[8], v0[9], v0[10], v0[11]));
            sts128(a + 48, make_uint4(v0[12], v0[13], v0[14], v0[15]));
            if (lane == 31) {
              sts128(a + 64, make_uint4(v1[0], v1[1], v1[2], v1[3]));
              sts128(a + 80, make_uint4(v1[4], v1[5], v1[6], v1[7]));
              sts128(a + 96, make_uint4(v1[8], v1[9], v1[10], v1[11]));
              sts128(a + 112, make_uint4(v1[12], v1[13], v1[14], v1[15]));
            }
          }
          named_bar_sync(bar_id, 128u);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v0[i] = __shfl_up_sync(0xFFFFFFFFu, v0[i], 2);
            v1[i] = __shfl_up_sync(0xFFFFFFFFu, v1[i], 1);
          }
          if (x_rd) {
            const uint32_t a = xq - 192u + (uint32_t)lane * 64u;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 t = lds128(a + 16u * g);
              v0[g * 4 + 0] = t.x; v0[g * 4 + 1] = t.y; v0[g * 4 + 2] = t.z; v0[g * 4 + 3] = t.w;
            }
            if (lane == 0) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint4 t = lds128(xq - 192u + 128u + 16u * g);
                v1[g * 4 + 0] = t.x; v1[g * 4 + 1] = t.y; v1[g * 4 + 2] = t.z; v1[g * 4 + 3] = t.w;
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            fadd2(v0[i], v0[i + 1], v1[i], v1[i + 1]);
            fadd2(v[i], v[i + 1], v0[i], v0[i + 1]);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            fadd2(v[g * 4 + 0], v[g * 4 + 1], __float_as_uint(bias[g].x), __float_as_uint(bias[g].y));
            fadd2(v[g * 4 + 2], v[g * 4 + 3], __float_as_uint(bias[g].z), __float_as_uint(bias[g].w));
          }
          fadd2(v[0], v[1], r0.x << 16, r0.x & 0xFFFF0000u);   fadd2(v[2], v[3], r0.y << 16, r0.y & 0xFFFF0000u);
          fadd2(v[4], v[5], r0.z << 16, r0.z & 0xFFFF0000u);   fadd2(v[6], v[7], r0.w << 16, r0.w & 0xFFFF0000u);
          fadd2(v[8], v[9], r1.x << 16, r1.x & 0xFFFF0000u);   fadd2(v[10], v[11], r1.y << 16, r1.y & 0xFFFF0000u);
          fadd2(v[12], v[13], r1.z << 16, r1.z & 0xFFFF0000u); fadd2(v[14], v[15], r1.w << 16, r1.w & 0xFFFF0000u);
          const float* f = reinterpret_cast<const float*>(v);
          uint4 o0, o1;
          if (e.relu) {
            o0 = make_uint4(pack_bf16_relu(f[0], f[1]), pack_bf16_relu(f[2], f[3]), pack_bf16_relu(f[4], f[5]),
                            pack_bf16_relu(f[6], f[7]));
            o1 = make_uint4(pack_bf16_relu(f[8], f[9]), pack_bf16_relu(f[10], f[11]), pack_bf16_relu(f[12], f[13]),
                            pack_bf16_relu(f[14], f[15]));
          } else {
            o0 = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            o1 = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
          }
          if (is_pad) { o0 = make_uint4(0, 0, 0, 0); o1 = o0; }  // zero cells of the padded layout stay zero
          if (prow >= 0) {
            sts128(ad0, o0);
            sts128(ad1, o1);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->acc_empty[buf]);
        DBG_TOCK(1);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->epi_done[group * 2 + sbuf]);
        DBG_TOCK(2);
      }
    } else if constexpr (kStagedEpi) {
      // ---- staged epilogue: batches of 128-row x panel_ch panels live in shared memory; the residual arrives by
      // TMA, each thread updates its own row cells in place, finished panels leave by TMA (issued by this group's
      // DMA warp).  Global memory only ever sees whole lines; no block-wide barrier is involved.
      const int panel_ch = p.panel_ch, npanels = nt / panel_ch, spp = panel_ch >> 4;  // 16-channel slices per panel
      const int PT = MB * npanels, bp = p.epi_batch;
      const uint32_t pitch = (uint32_t)panel_ch * 2u, swz = (uint32_t)p.panel_swz;
      const uint32_t panel_bytes = p.epi_panel_bytes;
      const uint32_t stage0 = a_base + p.epi_base_off + (uint32_t)group * 2u * (uint32_t)bp * panel_bytes;
      const long long tile_step = tstride * (n_accbuf >= 2 ? 2 : 1);
      const long long first_tile = tile0 + (n_accbuf >= 2 ? (long long)group * tstride : 0);
      const bool active = n_accbuf >= 2 || group == 0;
      const int upp = (spp - sub + 1) >> 1;  // 16-channel units of one panel handled by this warp
      const uint32_t xr = swz == 128 ? (uint32_t)(row0 & 7) : (swz == 64 ? (uint32_t)((row0 >> 1) & 3) : 0u);
      uint32_t kb = 0;  // batches processed by this group
      acc_it = n_accbuf >= 2 ? (uint32_t)group : 0u;
#pragma unroll 1
      for (long long tile = first_tile; active && tile < total_tiles; tile += tile_step, acc_it += (n_accbuf >= 2 ? 2 : 1)) {
        const int nti = (int)(tile % n_ntiles);
        const long long mt = mtile(tile);
        const uint32_t buf = acc_it % (uint32_t)n_accbuf;
        const uint32_t aph = (acc_it / n_accbuf) & 1;
        const int chbase = nti * nt;
        const uint32_t t_tile = tmem_base + buf * acc_cols + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
        for (int b0 = 0; b0 < PT; b0 += bp, ++kb) {
          const int cnt = PT - b0 < bp ? PT - b0 : bp;
          const int nunits = cnt * upp;  // <= 3 by construction (host)
          const uint32_t sbuf = kb & 1;
          const uint32_t stage = stage0 + sbuf * (uint32_t)bp * panel_bytes;
          if (b0 == 0) {
            DBG_TOCK(1);
            mbar_wait(&ctl->acc_full[buf], aph);
            DBG_TOCK(0);
            tc_fence_after();
          }
          bool ready = false;
          // units of this warp: panels pi = 0 .. cnt-1 of the batch, 16-channel slices sl = sub, sub+2, .. of each panel.
          // Written for instruction count (the drain competes with four other roles for issue slots): incremental
          // indices, the zero-cell test once per 128-row block, packed fp32x2 adds.
          int pi = 0, sl = sub, m_cached = -1;
          bool is_pad = false, stat_ok = false;
          RowPos upos;          // structured tiles with upsampled addends: the output pixel of this thread's row
          upos.valid = false;
#pragma unroll 1
          for (int u = 0; u < nunits; ++u) {
            const int idx = b0 + pi;
            const int m = npanels == 1 ? idx : idx >> 1, pn = npanels == 1 ? 0 : idx & 1;   // npanels is 1 or 2 (nt <= 128)
            const int ch = pn * panel_ch + sl * 16;
            uint32_t v[16];
            tmem_ld16(t_tile + (uint32_t)(m * nt + ch), v);
            if (m != m_cached) {
              if (!kStruct) {
                const RowPos rp = row_position(p, mt, m * 128 + row0);
                is_pad = rp.is_pad;                                                  // (structured tiles: valid pixels only)
                stat_ok = rp.valid && !rp.is_pad;
              } else if (e.n_up || STATS) {
                upos = row_position(p, mt, m * 128 + row0);
                stat_ok = upos.valid;
              }
              m_cached = m;
            }
            const uint32_t base = stage + (uint32_t)pi * panel_bytes + (uint32_t)row0 * pitch;
            const uint32_t c0 = (uint32_t)sl * 2u;  // 16-byte chunks of the slice in its panel row
            const uint32_t ad0 = base + ((c0 ^ xr) << 4), ad1 = base + (((c0 + 1) ^ xr) << 4);
            float4 bias[4];
            const float4* b4 = reinterpret_cast<const float4*>(sbias + chbase + ch);
#pragma unroll
            for (int g = 0; g < 4; ++g) bias[g] = b4[g];
            if (!ready) {
              // staging buffer ready: residual landed (which implies the buffer was free), or buffer free
              if (has_res) mbar_wait(&ctl->res_full[group * 2 + sbuf], (kb >> 1) & 1);
              else mbar_wait(&ctl->epi_free[group * 2 + sbuf], ((kb >> 1) & 1) ^ 1);
              ready = true;
              DBG_TOCK(3);
            }
            uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
            if (has_res) { r0 = lds128(ad0); r1 = lds128(ad1); }
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              fadd2(v[g * 4 + 0], v[g * 4 + 1], __float_as_uint(bias[g].x), __float_as_uint(bias[g].y));
              fadd2(v[g * 4 + 2], v[g * 4 + 3], __float_as_uint(bias[g].z), __float_as_uint(bias[g].w));
            }
            if (has_res) {
              fadd2(v[0], v[1], r0.x << 16, r0.x & 0xFFFF0000u);   fadd2(v[2], v[3], r0.y << 16, r0.y & 0xFFFF0000u);
              fadd2(v[4], v[5], r0.z << 16, r0.z & 0xFFFF0000u);   fadd2(v[6], v[7], r0.w << 16, r0.w & 0xFFFF0000u);
              fadd2(v[8], v[9], r1.x << 16, r1.x & 0xFFFF0000u);   fadd2(v[10], v[11], r1.y << 16, r1.y & 0xFFFF0000u);
              fadd2(v[12], v[13], r1.z << 16, r1.z & 0xFFFF0000u); fadd2(v[14], v[15], r1.w << 16, r1.w & 0xFFFF0000u);
            }
            if (kStruct && e.n_up && upos.valid) {
              // nearest-upsampled addends of the fuse row (HRnet.py:198-209, 258-264), gathered from the low-resolution maps
#pragma unroll 1
              for (int uu = 0; uu < e.n_up; ++uu) {
                const int sh = p.up_shift[uu];
                const int hs = e.H >> sh, ws = e.W >> sh;
                const size_t qs = ((size_t)upos.n * (hs + 1) + (upos.h >> sh)) * (ws + 1) + (upos.w >> sh);
                const uint4* src = reinterpret_cast<const uint4*>(p.up_src[uu] + qs * e.cout + chbase + ch);
                const uint4 t0 = __ldcg(src), t1 = __ldcg(src + 1);   // coherent loads: see load_bf16_row
                fadd2(v[0], v[1], t0.x << 16, t0.x & 0xFFFF0000u);   fadd2(v[2], v[3], t0.y << 16, t0.y & 0xFFFF0000u);
                fadd2(v[4], v[5], t0.z << 16, t0.z & 0xFFFF0000u);   fadd2(v[6], v[7], t0.w << 16, t0.w & 0xFFFF0000u);
                fadd2(v[8], v[9], t1.x << 16, t1.x & 0xFFFF0000u);   fadd2(v[10], v[11], t1.y << 16, t1.y & 0xFFFF0000u);
                fadd2(v[12], v[13], t1.z << 16, t1.z & 0xFFFF0000u); fadd2(v[14], v[15], t1.w << 16, t1.w & 0xFFFF0000u);
              }
            }
            const float* f = reinterpret_cast<const float*>(v);
            uint4 o0, o1;
            if (e.relu) {
              o0 = make_uint4(pack_bf16_relu(f[0], f[1]), pack_bf16_relu(f[2], f[3]), pack_bf16_relu(f[4], f[5]),
                              pack_bf16_relu(f[6], f[7]));
              o1 = make_uint4(pack_bf16_relu(f[8], f[9]), pack_bf16_relu(f[10], f[11]), pack_bf16_relu(f[12], f[13]),
                              pack_bf16_relu(f[14], f[15]));
            } else {
              o0 = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
              o1 = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]),
                              pack_bf16(f[14], f[15]));
            }
            if (is_pad) { o0 = make_uint4(0, 0, 0, 0); o1 = o0; }  // zero cells of the padded layout stay zero
            sts128(ad0, o0);
            sts128(ad1, o1);
            if constexpr (STATS) {
              // batch statistics of the STORED values (what BatchNorm will normalise), valid pixels only
              float r[16];
              r[0] = bf16_lo(o0.x); r[1] = bf16_hi(o0.x); r[2] = bf16_lo(o0.y); r[3] = bf16_hi(o0.y);
              r[4] = bf16_lo(o0.z); r[5] = bf16_hi(o0.z); r[6] = bf16_lo(o0.w); r[7] = bf16_hi(o0.w);
              r[8] = bf16_lo(o1.x); r[9] = bf16_hi(o1.x); r[10] = bf16_lo(o1.y); r[11] = bf16_hi(o1.y);
              r[12] = bf16_lo(o1.z); r[13] = bf16_hi(o1.z); r[14] = bf16_lo(o1.w); r[15] = bf16_hi(o1.w);
              if (!stat_ok) {
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = 0.f;
              }
              const float cs = warp_colsum16(r, lane);
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] *= r[i];
              const float cq = warp_colsum16(r, lane);
              const int slot = (nti * npanels + pn) * ((spp + 1) >> 1) + (sl >> 1);
#pragma unroll
              for (int k = 0; k < kStatSlots; ++k)
                if (slot == k) { st_sum[k] += cs; st_sq[k] += cq; }
            }
            sl += 2;
            if (sl >= spp) { sl = sub; ++pi; }
          }
          if (!ready) {  // a warp without units in this batch still has to observe the buffer hand-over in order
            if (has_res) mbar_wait(&ctl->res_full[group * 2 + sbuf], (kb >> 1) & 1);
            else mbar_wait(&ctl->epi_free[group * 2 + sbuf], ((kb >> 1) & 1) ^ 1);
          }
          if (b0 + bp >= PT) {  // accumulator fully drained: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR) mbar_arrive_cluster(map_to_cta(smem_u32(&ctl->acc_empty[buf]), 0));  // leader's barrier
              else mbar_arrive(&ctl->acc_empty[buf]);
            }
          }
          DBG_TOCK(1);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->epi_done[group * 2 + sbuf]);
          DBG_TOCK(2);
        }
      }
    } else
#pragma unroll 1
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acc_it) {
      if (n_accbuf >= 2 ? (int)(acc_it & 1) != group : group != 0) continue;
      const int nti = (int)(tile % n_ntiles);
      const long long mt = tile / n_ntiles;
      const uint32_t buf = acc_it % (uint32_t)n_accbuf;
      const uint32_t aph = (acc_it / n_accbuf) & 1;
      const int chbase = nti * nt;
      RowPos pos = row_position(p, mt, row0);
      // the residual of the first slice is fetched before waiting for the accumulator; afterwards the fetch of
      // slice s+1 is issued before slice s is processed
      uint4 rcur[2], rnext[2];
      if (has_res && sub < nslices)
        load_bf16_row<16>(rcur, e.residual + (size_t)pos.q * e.cout + chbase + sub * 16, pos.valid && !pos.is_pad);
      DBG_TOCK(1);
      mbar_wait(&ctl->acc_full[buf], aph);
      DBG_TOCK(0);
      tc_fence_after();
      const uint32_t t_tile = tmem_base + buf * acc_cols + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int m = 0; m < MB; ++m) {
        RowPos pos_next = pos;
        if (m + 1 < MB) pos_next = row_position(p, mt, (m + 1) * 128 + row0);
#pragma unroll 1
        for (int sl = sub; sl < nslices; sl += 2) {
          uint32_t v[16];
          tmem_ld16(t_tile + (uint32_t)(m * nt + sl * 16), v);
          float4 bias[4];
          const float4* b4 = reinterpret_cast<const float4*>(sbias + chbase + sl * 16);
#pragma unroll
          for (int g = 0; g < 4; ++g) bias[g] = b4[g];
          if (has_res) {
            const bool same_row = sl + 2 < nslices;
            const RowPos& pn = same_row ? pos : pos_next;
            const int cn = same_row ? (sl + 2) * 16 : sub * 16;
            if (same_row || m + 1 < MB)
              load_bf16_row<16>(rnext, e.residual + (size_t)pn.q * e.cout + chbase + cn, pn.valid && !pn.is_pad);
          }
          tmem_ld_wait();
          epilogue_store<16, NCHW>(p, e, v, chbase + sl * 16, pos, rcur, bias);
          rcur[0] = rnext[0];
          rcur[1] = rnext[1];
        }
        pos = pos_next;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->acc_empty[buf]);
      DBG_TOCK(1);
    }
    if (warp == 2) DBG_DUMP(2);
  }
role_done:
#undef DBG_TICK
#undef DBG_TOCK
#undef DBG_DUMP

  tc_fence_before();
  if (PAIR) cluster_sync(); else __syncthreads();  // PAIR: neither CTA may retire while the other can still signal it
  if constexpr (STATS) {
    // every MMA and every TMA transfer of this CTA has completed: the activation ring is free.  Each epilogue warp writes
    // its slot accumulators to row (group, quarter) of an [8][2][cout_pad] table there (the two warps of a quarter own
    // disjoint 16-channel slices), then the rows are added in a fixed order into this CTA's row of p.stats.
    float* tab = reinterpret_cast<float*>(smem + p.ctl_bytes);
    const int cp = p.cout_pad;
    if (warp >= 2 && warp < 18 && !(lane & 1)) {
      const int e_idx = warp - 2, group = e_idx >> 3, sub = (e_idx >> 2) & 1, quarter = warp & 3;
      const int panel_ch = p.panel_ch, npanels = p.nt / panel_ch, spp = panel_ch >> 4, pairs = (spp + 1) >> 1;
      float* row = tab + (size_t)(group * 4 + quarter) * 2 * cp;
#pragma unroll
      for (int k = 0; k < kStatSlots; ++k) {
        const int sp = k % pairs, pnl = (k / pairs) % npanels, nti = k / (pairs * npanels);
        const int sl = 2 * sp + sub;
        if (nti < p.n_ntiles && sl < spp) {
          const int ch = nti * p.nt + pnl * panel_ch + sl * 16 + ((lane >> 1) & 15);
          row[ch] = st_sum[k];
          row[cp + ch] = st_sq[k];
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * cp; i += (int)blockDim.x) {
      float acc = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) acc += tab[(size_t)r * 2 * cp + i];
      p.stats[(size_t)blockIdx.x * 2 * cp + i] = acc;
    }
    if (p.stats_ticket) {
      // the last CTA to arrive adds the rows of all CTAs in a fixed order and finalises the batch statistics: no launch
      // between this convolution and the normalisation.  Thread = (column of [2][cout_pad], row group): kGroups row
      // groups with four independent partial sums each, combined in order through shared memory.
      // (no static shared memory in this kernel: the dynamic allocation already asks for the 227 KB maximum)
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) ctl->last_cta = atomicAdd(p.stats_ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
      __syncthreads();
      if (ctl->last_cta) {
        __threadfence();
        const int cols = 2 * cp, rows = (int)gridDim.x;         // cols <= blockDim.x (host)
        int groups = (int)blockDim.x / cols;
        groups = groups > 8 ? 8 : groups;
        const int col = (int)threadIdx.x % cols, g = (int)threadIdx.x / cols;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (g < groups) {
          int r = g;
          for (; r + 3 * groups < rows; r += 4 * groups) {
            a0 += __ldcg(p.stats + (size_t)r * cols + col);
            a1 += __ldcg(p.stats + (size_t)(r + groups) * cols + col);
            a2 += __ldcg(p.stats + (size_t)(r + 2 * groups) * cols + col);
            a3 += __ldcg(p.stats + (size_t)(r + 3 * groups) * cols + col);
          }
          for (; r < rows; r += groups) a0 += __ldcg(p.stats + (size_t)r * cols + col);
        }
        __syncthreads();                                          // (the per-CTA reduction above is done with `tab`)
        if (g < groups) tab[(size_t)g * cols + col] = (a0 + a1) + (a2 + a3);
        __syncthreads();
        if (g == 0) {
          float t = 0.f;
          for (int k = 0; k < groups; ++k) t += tab[(size_t)k * cols + col];
          tab[(size_t)8 * cols + col] = t;                        // totals: row 8 of the table
        }
        __syncthreads();
        for (int c = threadIdx.x; c < p.cout; c += (int)blockDim.x) {
          const float* tot = tab + (size_t)8 * cols;
          const float m = tot[c] / p.stats_count;
          float var = tot[cp + c] / p.stats_count - m * m;
          var = var > 0.f ? var : 0.f;
          p.stats_mean[c] = m;
          p.stats_rstd[c] = rsqrtf(var + p.stats_eps);
          if (p.stats_run_mean) {
            const float unbiased = p.stats_count > 1.f ? var * p.stats_count / (p.stats_count - 1.f) : var;
            p.stats_run_mean[c] = (1.f - p.stats_momentum) * p.stats_run_mean[c] + p.stats_momentum * m;
            p.stats_run_var[c] = (1.f - p.stats_momentum) * p.stats_run_var[c] + p.stats_momentum * unbiased;
          }
        }
        if (threadIdx.x == 0) *p.stats_ticket = 0u;
      }
    }
  }
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
           const cuuint32_t* box, const cuuint32_t* estrides, uint32_t span) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return 1; }
  CUtensorMapSwizzle sw = span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : span == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : span == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                       : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                  strides_bytes, box, estrides, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims %llu,%llu box %u,%u span %u", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1], span);
    return 1;
  }
  return 0;
}

int pick_ck(int cin) {
  if (cin % 64 == 0) return 64;
  if (cin % 32 == 0) return 32;
  if (cin % 16 == 0) return 16;
  return 0;
}

int pick_nt(int cout_pad) {
  // largest multiple of 16 that divides cout_pad and is <= 128
  for (int nt = 128; nt >= 16; nt -= 16)
    if (cout_pad % nt == 0) return nt;
  return 0;
}

int num_sms() { return device_sm_count(); }

}  // namespace

int conv_prepare(const ConvSpec& s, ConvParams* pp, int* grid, size_t* smem_bytes) {
  ConvParams& p = *pp;
  memset(&p, 0, sizeof(p));
  const PaddedGeom& gi = s.in_geom;
  if (!(s.ksize == 1 || s.ksize == 3) || !(s.stride == 1 || s.stride == 2) || (s.stride == 2 && s.ksize != 3)) {
    set_error("conv: unsupported ksize %d / stride %d", s.ksize, s.stride);
    return 1;
  }
  if (s.cout_pad % 16 || s.cout_pad < s.cout || s.cout_pad > kMaxCoutPad || (!s.out_nchw && s.cout != s.cout_pad)) {
    set_error("conv: bad cout %d / cout_pad %d", s.cout, s.cout_pad);
    return 1;
  }
  // K = the channels of both inputs (ConvSpec::in2: 1x1 / stride 1 only, each a whole number of K chunks)
  const int cin_total = gi.C + (s.in2 ? s.in2_C : 0);
  if (s.in2 && (s.ksize != 1 || s.stride != 1 || s.in2_C <= 0 || s.in2_C % 16)) {
    set_error("conv: a second input needs a 1x1 stride-1 convolution and a multiple of 16 channels");
    return 1;
  }
  p.ck = s.in2 ? pick_ck(gi.C) < pick_ck(s.in2_C) ? pick_ck(gi.C) : pick_ck(s.in2_C) : pick_ck(gi.C);
  // 64 -> 64 3x3 stride 1 (layer1 conv2, the 64-channel branch): two 32-channel K chunks instead of one of 64.  The
  // activation ring then holds four half-size stages instead of two (same bytes), i.e. loads run 1.5 tiles ahead of the
  // MMAs instead of one: the issuing warp's wait for activations drops from 25 % to 17 % of its time at 64x48
  // (301 -> 258 us per launch, 84.7 -> 83.3 us at 32x24; gpurun_out/r02_convprobe2.txt).  STL_DBG_CK overrides.
  bool ck_split = gi.C == 64 && s.cout_pad == 64 && s.ksize == 3 && s.stride == 1 && !s.force_tap_reload;
  if (const char* e = getenv("STL_DBG_CK")) {
    const int v = atoi(e);
    ck_split = false;
    if ((v == 16 || v == 32) && gi.C % v == 0 && v < p.ck) p.ck = v;
  }
  if (ck_split) p.ck = 32;
  // Channel counts that are not a multiple of 32 or 64 (HRNet-W48: 48, 96): as 16- / 32-channel chunks every tap needs
  // three chunks (48 channels: one K step per chunk, 32-byte swizzle: 22.5 % tensor-pipe activity, profiles/r02_ncu.md).
  // 64-channel TMA boxes instead - the box of the last chunk reaches past the tensor's channels, which arrive as zeros -
  // and only the K steps that hold data are issued for that chunk.  STL_DBG_CK=16/32 restores the small chunks.
  p.ksteps_last = 0;
  if (!s.in2 && gi.C > 32 && gi.C % 64 != 0 && gi.C % 16 == 0 && !getenv("STL_DBG_CK")) {
    p.ck = 64;
    p.ksteps_last = (gi.C % 64) / 16;
  }
  p.nt = pick_nt(s.cout_pad);
  if (!p.ck || !p.nt) { set_error("conv: channels must be multiples of 16 (cin %d cout_pad %d)", gi.C, s.cout_pad); return 1; }
  if (s.stride == 2 && ((gi.H | gi.W) & 1)) { set_error("conv: stride 2 needs even H, W"); return 1; }
  const uint32_t span = 2u * p.ck;
  p.mode = s.stride == 2 ? 1 : 0;
  p.taps = s.ksize * s.ksize;
  p.n_chunks = (cin_total + p.ck - 1) / p.ck;
  p.n_chunks_a = s.in2 ? gi.C / p.ck : p.n_chunks;
  if (!p.ksteps_last) p.ksteps_last = p.ck / 16;
  p.n_ntiles = s.cout_pad / p.nt;
  p.in_Wp = gi.Wp();
  p.N = gi.N;
  p.H = gi.H / s.stride;
  p.W = gi.W / s.stride;
  p.Hp = p.H + 1;
  p.Wp = p.W + 1;
  p.P = (long long)p.N * p.Hp * p.Wp;
  p.q_lo = 0;
  if (s.img_hi > 0) {  // only images [img_lo, img_hi): the tiles start at the first image's row and the kernel's notion
                       // of "end of tensor" (P, and the bounds of the output / residual tensor maps) is the range end
    if (s.img_lo < 0 || s.img_hi > gi.N || s.img_lo >= s.img_hi || (p.mode != 0 && (s.img_lo != 0 || s.img_hi != gi.N))) {
      set_error("conv: bad image range [%d,%d) of %d (sub-ranges need a stride-1 conv)", s.img_lo, s.img_hi, gi.N);
      return 1;
    }
    p.q_lo = s.img_lo * p.Hp * p.Wp;
    p.P = (long long)s.img_hi * p.Hp * p.Wp;
  }
  if ((long long)gi.pixels() >= (1ll << 31) || p.P >= (1ll << 31)) { set_error("conv: tensor too large"); return 1; }

  // kw-merged MMA shape (see kEpiStagedKW): stride-1 3x3 with one N tile, 3*nt accumulator columns x 2 buffers, the three
  // tap tiles of a filter row contiguous in the resident weight block (no padding between stages), staged epilogue.
  // Falls back (recursive call with kw_merge = -1) when the weights turn out not to stay resident.
  // Measured on B200: the MMA stream gets 40 % shorter but the epilogue's row exchange makes the
  // drain the bottleneck, 89 vs 83 us on the 64-channel layers - so it is off unless asked for (spec.kw_merge = 1 /
  // stl_conv_desc.impl = 3 / STLPOSE_KW_MERGE=1).
#ifdef STL_KW_MERGE   // measured-negative variant: only in builds made with -DSTL_KW_MERGE (profiles/r01_kwm_experiment.md)
  const bool kwm_wanted = s.kw_merge == 1 || (s.kw_merge == 0 && getenv("STLPOSE_KW_MERGE") && atoi(getenv("STLPOSE_KW_MERGE")) == 1);
#else
  const bool kwm_wanted = false;
#endif
  const bool kwm = kwm_wanted && p.mode == 0 && p.taps == 9 && !s.force_tap_reload && !s.force_mb && !s.out_nchw &&
                   s.n_up == 0 && p.n_ntiles == 1 && 3 * p.nt <= 256 && (p.nt * span) % 1024 == 0 && s.img_hi == 0 &&
                   !getenv("STL_DBG_NO_EPI_TMA") && !getenv("STL_DBG_NO_RESIDENT") && !getenv("STL_DBG_PAIR_ALL");
  p.kwm = kwm ? 1 : 0;
  // accumulator blocks per tile: keep two accumulator buffers in 512 TMEM columns when possible
  int mb = s.force_mb ? s.force_mb : (kwm ? 1 : (p.nt <= 64 ? 3 : 2));
  if (mb > 3) mb = 3;
  while (mb > 1 && mb * p.nt > 256) --mb;
  if (p.mode == 0) {
    // do not make tiles larger than the problem
    while (mb > 1 && (long long)(mb - 1) * 128 >= p.P - p.q_lo) --mb;
  }

  if (p.mode == 0) {
    p.mb = mb;
    p.a_shift = (p.taps == 9 && !s.force_tap_reload) ? 1 : 0;
    p.halo = p.a_shift ? p.in_Wp + (kwm ? 0 : 1) : 0;   // kw-merged: the A operand only shifts by whole image rows
    p.tile_rows = kwm ? 126 : 128 * mb;
    const int rows_needed = 128 * mb + 2 * p.halo;
    p.a_pieces = (rows_needed + 255) / 256;
    p.a_box_rows = (((rows_needed + p.a_pieces - 1) / p.a_pieces) + 7) & ~7;
    p.total_tiles = ((p.P - p.q_lo + p.tile_rows - 1) / p.tile_rows) * p.n_ntiles;
    cuuint64_t dims[2] = {(cuuint64_t)gi.C, (cuuint64_t)gi.pixels()};
    cuuint64_t strides[1] = {(cuuint64_t)gi.C * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.ck, (cuuint32_t)p.a_box_rows};
    cuuint32_t es[2] = {1, 1};
    if (encode(&p.tmA, s.in, 2, dims, strides, box, es, span)) return 1;
    if (s.in2) {
      cuuint64_t dims2[2] = {(cuuint64_t)s.in2_C, (cuuint64_t)gi.pixels()};
      cuuint64_t strides2[1] = {(cuuint64_t)s.in2_C * 2};
      if (encode(&p.tmA2, s.in2, 2, dims2, strides2, box, es, span)) return 1;
    }
  } else {
    // structured tile: every 128-row accumulator block is a (bw x bh x bn1) box of output pixels (bw*bh*bn1 = 128, the
    // widest bw first), so a staged panel is one 4-D TMA box; a tile stacks mb blocks along the image index
    bool found = false;
    for (int bw = p.W < 128 ? p.W : 128; bw >= 1 && !found; --bw) {
      if (p.W % bw || 128 % bw) continue;
      const int rem = 128 / bw;
      for (int bh = p.H < rem ? p.H : rem; bh >= 1 && !found; --bh) {
        if (p.H % bh || rem % bh) continue;
        const int bn1 = rem / bh;
        if (bn1 * mb > 256) continue;
        p.bw = bw; p.bh = bh; p.bn = bn1 * mb; p.mb = mb;
        found = true;
      }
    }
    if (!found) { set_error("conv: no structured tile for %dx%d output", p.H, p.W); return 1; }
    p.a_shift = 0;
    p.halo = 0;
    p.a_pieces = 1;
    p.a_box_rows = 128 * p.mb;
    p.tiles_w = p.W / p.bw;
    p.tiles_h = p.H / p.bh;
    p.tiles_n = (p.N + p.bn - 1) / p.bn;
    p.total_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_n * p.n_ntiles;
    cuuint64_t dims[4] = {(cuuint64_t)gi.C, (cuuint64_t)gi.Wp(), (cuuint64_t)gi.Hp(), (cuuint64_t)gi.N};
    cuuint64_t strides[3] = {(cuuint64_t)gi.C * 2, (cuuint64_t)gi.C * 2 * gi.Wp(),
                             (cuuint64_t)gi.C * 2 * gi.Wp() * gi.Hp()};
    cuuint32_t box[4] = {(cuuint32_t)p.ck, (cuuint32_t)(2 * p.bw), (cuuint32_t)(2 * p.bh), (cuuint32_t)p.bn};
    cuuint32_t es[4] = {1, 2, 2, 1};
    if (encode(&p.tmA, s.in, 4, dims, strides, box, es, span)) return 1;
  }
  p.fd_Wp.init((uint32_t)p.Wp);
  p.fd_Hp.init((uint32_t)p.Hp);
  p.fd_bw.init((uint32_t)(p.bw > 0 ? p.bw : 1));
  p.fd_bh.init((uint32_t)(p.bh > 0 ? p.bh : 1));
  p.a_tx_bytes = (uint32_t)p.a_pieces * p.a_box_rows * span;
  p.a_stage_bytes = (p.a_tx_bytes + 1023u) & ~1023u;
  // accumulator buffers in the 512 TMEM columns: 2 (tile t+1 is multiplied while tile t drains), else 1.
  // Four buffers (STL_DBG_ACC4) were measured on B200 and change nothing: the narrow layers are not limited by the
  // TMEM hand-off but by operand fetch and HBM.
  p.n_accbuf = (4 * p.mb * p.nt <= 512 && getenv("STL_DBG_ACC4")) ? 4 : ((2 * p.mb * p.nt <= 512) ? 2 : 1);
  if (kwm) p.n_accbuf = 2;   // 2 x 3*nt <= 512 columns by construction
  uint32_t cols = 32;
  while (cols < (uint32_t)(p.n_accbuf * (kwm ? 3 : p.mb) * p.nt)) cols <<= 1;
  p.tmem_cols = cols;

  // weights stay resident in shared memory when every tile of the layer fits next to >= 2 activation stages;
  // otherwise they stream through a ring (one stage per (chunk, tap))
  // staged (TMA) epilogue for flat-mode bf16 outputs: 2 groups x 2 panel buffers x 16 KB
  // (upsampled addends: only the structured staged epilogue gathers them - that is where the fuse rows put them)
  p.epi_tma = (!s.out_nchw && (s.n_up == 0 || (p.mode == 1 && !getenv("STL_DBG_NO_EPI_TMA_UP"))) &&
               !getenv("STL_DBG_NO_EPI_TMA") && (p.mode == 0 || (p.taps == 9 && !getenv("STL_DBG_NO_EPI_TMA_S2")))) ? 1 : 0;
  p.panel_ch = (p.nt % 64 == 0) ? 64 : p.nt;
  p.panel_swz = p.panel_ch == 64 ? 128 : (p.panel_ch == 32 ? 64 : 0);
  p.epi_panel_bytes = (uint32_t)((p.panel_ch * 2 * 128 + 1023) & ~1023);
  if (p.epi_panel_bytes > 16384) p.epi_tma = 0;
  // panels per batch: a whole tile when its panels are small (8 KB), else one; each warp then has at most 3
  // 16-channel units per batch (slices of a panel alternate between the two warps of a TMEM lane quarter)
  p.epi_batch = p.epi_panel_bytes <= 8192 ? mb : 1;
  {
    const int spp = p.panel_ch / 16;
    while (p.epi_batch > 1 && p.epi_batch * ((spp + 1) / 2) > 3) --p.epi_batch;
    if ((spp + 1) / 2 > 3) p.epi_tma = 0;
  }
  const size_t xch_bytes = kwm ? 2 * 2 * 2 * 4 * 192 : 0;   // [group][sub][parity][quarter][3 rows x 16 fp32]
  const size_t epi_bytes = (p.epi_tma ? (size_t)2 * 2 * p.epi_batch * p.epi_panel_bytes : 0) + xch_bytes;
  // control block: 1 KB of barriers + the bias rounded up to 1 KB (2 KB instead of 4 for <= 256 output channels; gives
  // 64 -> 64 3x3 @64x48 a third activation stage - measured neutral, 286 vs 283 us)
  p.ctl_bytes = 1024u + (((uint32_t)s.cout_pad * 4u + 1023u) & ~1023u);
  if (p.ctl_bytes > kCtlBytes) { set_error("conv: cout_pad %d exceeds the bias area", s.cout_pad); return 1; }
  const size_t budget = kMaxSmem - p.ctl_bytes - 1024 - epi_bytes;
  // CTA pairs (cta_group::2): flat mode, staged epilogue, burst-capable; each CTA then keeps half of the weight rows
  // Measured on B200: for layers whose weights stay resident (N <= 64, every 1x1) pairs are slower - the MMA rate is
  // bound by the A-operand fetch (128 rows x 32 B per SM per instruction), which pairing does not reduce.  For
  // layers that stream their weights (C >= 128) each CTA of a pair streams only half of the rows: half the L2 traffic.
  const bool pair_ok = p.mode == 0 && p.epi_tma && p.nt % 32 == 0 && !kwm && !getenv("STL_DBG_NO_PAIR");
  size_t resident = 0;
  p.pair = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    p.b_tx_bytes = (uint32_t)(p.pair ? p.nt / 2 : p.nt) * span;
    p.b_stage_bytes = (p.b_tx_bytes + 1023u) & ~1023u;
    resident = (size_t)p.n_ntiles * p.n_chunks * p.taps * p.b_stage_bytes;
    const bool fits = resident <= 120 * 1024 && resident + 2 * (size_t)p.a_stage_bytes <= budget &&
                      !getenv("STL_DBG_NO_RESIDENT");
    if (fits || !pair_ok || p.pair || (getenv("STL_DBG_PAIR_ALL") == nullptr && resident <= 120 * 1024)) break;
    p.pair = 1;  // weights must stream: share every weight tile between the two CTAs of a pair
  }
  if (getenv("STL_DBG_PAIR_ALL") && pair_ok && !p.pair) {
    p.pair = 1;
    p.b_tx_bytes = (uint32_t)(p.nt / 2) * span;
    p.b_stage_bytes = (p.b_tx_bytes + 1023u) & ~1023u;
    resident = (size_t)p.n_ntiles * p.n_chunks * p.taps * p.b_stage_bytes;
  }
  // Streamed weights (C >= 128 layers): one ring stage holds the three taps of a filter ROW (one 3-D TMA box), so the
  // issuing warp waits for weights and commits the slot back once per 3 x KSTEPS x MB MMAs instead of once per
  // KSTEPS x MB.  Measured on B200 with the per-role counters (gpurun_out/r02_convprobe*.txt): with one tap per stage
  // the leader spends 52 % of its time waiting for the next stage although the ring is full - a fixed ~450 cycles per
  // (wait, MMAs, commit) round that neither the ring depth nor the bytes moved change.  STL_DBG_B_TAPS=1 restores it.
  const bool will_stream = !(resident <= 120 * 1024 && resident + 2 * (size_t)p.a_stage_bytes <= budget &&
                             resident < (1u << 20) && !getenv("STL_DBG_NO_RESIDENT"));
  p.b_taps = 1;
  if (will_stream && p.taps == 9 && p.b_tx_bytes % 1024 == 0 && !(getenv("STL_DBG_B_TAPS") && atoi(getenv("STL_DBG_B_TAPS")) == 1)) {
    p.b_taps = 3;
    p.b_tx_bytes *= 3;
    p.b_stage_bytes = p.b_tx_bytes;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)cin_total, (cuuint64_t)s.cout_pad, (cuuint64_t)p.taps};
    cuuint64_t strides[2] = {(cuuint64_t)cin_total * 2, (cuuint64_t)cin_total * 2 * s.cout_pad};
    cuuint32_t box[3] = {(cuuint32_t)p.ck, (cuuint32_t)(p.pair ? p.nt / 2 : p.nt), (cuuint32_t)p.b_taps};
    cuuint32_t es[3] = {1, 1, 1};
    if (encode(&p.tmB, s.weights, 3, dims, strides, box, es, span)) return 1;
  }
  if (p.pair) p.total_tiles = (((p.P - p.q_lo + 128 * p.mb - 1) / (128 * p.mb) + 1) / 2) * p.n_ntiles;
  const int a_loads_per_tile = p.n_chunks * (p.a_shift || p.taps == 1 ? 1 : p.taps);
  int a_want = a_loads_per_tile >= 4 ? 4 : (p.a_shift ? 3 : 4);
  if (ck_split) a_want = 4;
  if (const char* e = getenv("STL_DBG_A_WANT")) { const int v = atoi(e); if (v >= 2 && v <= kMaxStages) a_want = v; }
  int a_st = 2, b_st = 2;
  p.b_resident = (resident <= 120 * 1024 && resident + 2 * (size_t)p.a_stage_bytes <= budget &&
                  resident < (1u << 20) && !getenv("STL_DBG_NO_RESIDENT")) ? 1 : 0;
  if (p.b_resident) {
    p.b_resident_bytes = (uint32_t)((size_t)p.n_ntiles * p.n_chunks * p.taps * p.b_tx_bytes);
    while (a_st < a_want && resident + (size_t)(a_st + 1) * p.a_stage_bytes <= budget) ++a_st;
    b_st = 1;
    p.b_bytes_total = (uint32_t)resident;
  } else {
    auto used = [&](int a, int b) { return (size_t)a * p.a_stage_bytes + (size_t)b * p.b_stage_bytes; };
    if (used(2, 2) > budget) a_st = 1;
    if (used(a_st, b_st) > budget) { set_error("conv: tile does not fit shared memory"); return 1; }
    const int b_want = p.b_taps == 3 ? 4 : 8;
    bool grew = true;
    while (grew) {
      grew = false;
      if (b_st < b_want && b_st < kMaxStages && used(a_st, b_st + 1) <= budget) { ++b_st; grew = true; }
      if (a_st < a_want && a_st < kMaxStages && used(a_st + 1, b_st) <= budget) { ++a_st; grew = true; }
    }
    if (const char* e = getenv("STL_DBG_B_STAGES")) { int v = atoi(e); if (v >= 1 && v <= b_st) b_st = v; }
    p.b_bytes_total = (uint32_t)((size_t)b_st * p.b_stage_bytes);
  }
  if (const char* e = getenv("STL_DBG_A_STAGES")) { int v = atoi(e); if (v >= 1 && v <= a_st) a_st = v; }
  p.a_stages = a_st;
  p.b_stages = b_st;
  if (kwm && !(p.b_resident && p.epi_tma && !p.pair)) {   // the merged shape needs the whole weight block in place
    ConvSpec s2 = s;
    s2.kw_merge = -1;
    return conv_prepare(s2, pp, grid, smem_bytes);
  }
  // BatchNorm statistics in the epilogue (training): staged epilogues only, at most kStatSlots (n-tile, panel, slice
  // pair) slots per epilogue warp, and the [8][2][cout_pad] reduction table must fit the activation ring it reuses.
  // Correct (tests/test_train_kernels_gpu.py::test_conv_epilogue_statistics) but not faster, in either form: with a
  // separate row-reduction launch 40.2 vs 39.8 ms per step at batch 128 and 17.9 vs 17.3 at batch 32; with the last CTA
  // finalising the statistics itself (stl_conv2d_bn: 292 kernel nodes fewer per step) 41.9 vs 41.5 and 18.45 vs 18.25.
  // The epilogue's 32 shuffles per 16-channel unit make the forward convolutions 24 % slower, which costs what the
  // statistics pass saved.  The variants are therefore compiled only with -DSTL_CONV_STATS; without it stl_conv2d_stats
  // / stl_conv2d_bn report "not for this shape" and the callers use the separate statistics kernel.
  p.stats = nullptr;
  p.stats_ticket = nullptr;
#ifdef STL_CONV_STATS
  if (s.stats) {
    const int spp = p.panel_ch / 16, npanels = p.nt / p.panel_ch;
    if (p.epi_tma && !kwm && s.cout == s.cout_pad && p.n_ntiles * npanels * ((spp + 1) / 2) <= kStatSlots &&
        (size_t)72 * s.cout_pad <= (size_t)a_st * p.a_stage_bytes && (!s.stats_ticket || 2 * s.cout_pad <= 640)) {
      p.stats = s.stats;
      p.stats_ticket = s.stats_ticket;
      p.stats_count = s.stats_count; p.stats_eps = s.stats_eps; p.stats_momentum = s.stats_momentum;
      p.stats_mean = s.stats_mean; p.stats_rstd = s.stats_rstd;
      p.stats_run_mean = s.stats_run_mean; p.stats_run_var = s.stats_run_var;
    }
  }
#endif
  // measured on B200: a second issuer pays off only for N <= 32 (16-cycle MMAs); at N = 64 the two streams interfere
  p.n_mma = (p.b_resident && (p.a_shift || p.taps == 1) && p.mb >= 2 && p.nt <= 32 && !getenv("STL_DBG_SINGLE_MMA")) ? 2 : 1;
  *smem_bytes = p.ctl_bytes + 1024 + (size_t)a_st * p.a_stage_bytes + p.b_bytes_total + epi_bytes;
  p.epi_base_off = (uint32_t)((size_t)a_st * p.a_stage_bytes + p.b_bytes_total);
  p.xch_off = p.epi_base_off + (uint32_t)(epi_bytes - xch_bytes);
  if (p.epi_tma && p.mode == 1) {
    // valid pixels only (W x H, not Wp x Hp): TMA clips partial tiles and never touches the zero cells
    cuuint64_t dims[4] = {(cuuint64_t)s.cout, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
    cuuint64_t strides[3] = {(cuuint64_t)s.cout * 2, (cuuint64_t)s.cout * 2 * p.Wp, (cuuint64_t)s.cout * 2 * p.Wp * p.Hp};
    cuuint32_t box[4] = {(cuuint32_t)p.panel_ch, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)(p.bn / p.mb)};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (encode(&p.tmO, s.out, 4, dims, strides, box, es, (uint32_t)p.panel_swz)) return 1;
    if (s.residual && encode(&p.tmR, s.residual, 4, dims, strides, box, es, (uint32_t)p.panel_swz)) return 1;
  } else if (p.epi_tma) {
    cuuint64_t dims[2] = {(cuuint64_t)s.cout, (cuuint64_t)p.P};
    cuuint64_t strides[1] = {(cuuint64_t)s.cout * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.panel_ch, (cuuint32_t)(kwm ? 126 : 128)};
    cuuint32_t es[2] = {1, 1};
    if (encode(&p.tmO, s.out, 2, dims, strides, box, es, (uint32_t)p.panel_swz)) return 1;
    if (s.residual && encode(&p.tmR, s.residual, 2, dims, strides, box, es, (uint32_t)p.panel_swz)) return 1;
  }

  p.out = s.out;
  p.bias = s.bias;
  p.residual = s.residual;
  p.n_up = s.n_up;
  for (int i = 0; i < kMaxUp; ++i) { p.up_src[i] = s.up_src[i]; p.up_shift[i] = s.up_shift[i]; }
  p.relu = s.relu;
  p.out_nchw = s.out_nchw;
  p.pdl = s.pdl;
  p.dbg_skip_epilogue = getenv("STL_DBG_SKIP_EPILOGUE") ? 1 : (getenv("STL_DBG_SKIP_STORE") ? 2 : 0);
  p.dbg_counters = reinterpret_cast<long long*>(s.dbg_counters);
  // (p.stats was decided above, next to the stage counts)
  p.cout = s.cout;
  p.cout_pad = s.cout_pad;

  long long g = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (s.max_ctas > 0 && g > s.max_ctas) g = s.max_ctas;
  if (p.pair) {  // one cluster of two CTAs per tile pair
    long long pairs = p.total_tiles < num_sms() / 2 ? p.total_tiles : num_sms() / 2;
    if (s.max_ctas > 1 && pairs > s.max_ctas / 2) pairs = s.max_ctas / 2;
    g = 2 * pairs;
  }
  *grid = (int)g;

  // a smaller tile can make the difference between streaming the weights for every tile and keeping them resident
  if (p.mode == 0 && !p.b_resident && !s.force_mb && p.mb > 1 && resident <= 120 * 1024) {
    ConvSpec s2 = s;
    ConvParams p2;
    int g2 = 0;
    size_t sm2 = 0;
    for (int mb2 = p.mb - 1; mb2 >= 1; --mb2) {
      s2.force_mb = mb2;
      if (conv_prepare(s2, &p2, &g2, &sm2) == 0 && p2.b_resident) {
        p = p2;
        *grid = g2;
        *smem_bytes = sm2;
        break;
      }
    }
  }
  return 0;
}

namespace {
typedef void (*ConvKernel)(const ConvParams);
template <int MB, int KSTEPS>
ConvKernel pick_variant(int taps, int epi, bool pair, bool stats) {
  if (stats) {   // training convolutions with BatchNorm statistics in the epilogue: the two staged epilogues only
#ifndef STL_CONV_STATS
    return nullptr;   // measured-negative: only in builds made with -DSTL_CONV_STATS
#else
    if (epi == kEpiStagedS2) return taps == 9 ? conv_tc_kernel<MB, KSTEPS, 9, kEpiStagedS2, false, true> : nullptr;
    if (epi != kEpiStaged) return nullptr;
    if (pair) return taps == 1 ? conv_tc_kernel<MB, KSTEPS, 1, kEpiStaged, true, true> : conv_tc_kernel<MB, KSTEPS, 9, kEpiStaged, true, true>;
    return taps == 1 ? conv_tc_kernel<MB, KSTEPS, 1, kEpiStaged, false, true> : conv_tc_kernel<MB, KSTEPS, 9, kEpiStaged, false, true>;
#endif
  }
  if (epi == kEpiNchw) return taps == 1 ? conv_tc_kernel<MB, KSTEPS, 1, kEpiNchw, false> : nullptr;
  if (epi == kEpiStagedS2) return taps == 9 ? conv_tc_kernel<MB, KSTEPS, 9, kEpiStagedS2, false> : nullptr;
  if (epi == kEpiStagedKW) {
#ifdef STL_KW_MERGE
    if constexpr (MB == 1) return taps == 9 ? conv_tc_kernel<1, KSTEPS, 9, kEpiStagedKW, false> : nullptr;
#endif
    return nullptr;
  }
  if (epi == kEpiStaged) {
    if (pair) return taps == 1 ? conv_tc_kernel<MB, KSTEPS, 1, kEpiStaged, true> : conv_tc_kernel<MB, KSTEPS, 9, kEpiStaged, true>;
    return taps == 1 ? conv_tc_kernel<MB, KSTEPS, 1, kEpiStaged, false> : conv_tc_kernel<MB, KSTEPS, 9, kEpiStaged, false>;
  }
  return taps == 1 ? conv_tc_kernel<MB, KSTEPS, 1, kEpiDirect, false> : conv_tc_kernel<MB, KSTEPS, 9, kEpiDirect, false>;
}
template <int MB>
ConvKernel pick_ksteps(int ksteps, int taps, int epi, bool pair, bool stats) {
  switch (ksteps) {
    case 1: return pick_variant<MB, 1>(taps, epi, pair, stats);
    case 2: return pick_variant<MB, 2>(taps, epi, pair, stats);
    case 4: return pick_variant<MB, 4>(taps, epi, pair, stats);
  }
  return nullptr;
}
ConvKernel pick_kernel(int mb, int ksteps, int taps, int epi, bool pair, bool stats = false) {
  switch (mb) {
    case 1: return pick_ksteps<1>(ksteps, taps, epi, pair, stats);
    case 2: return pick_ksteps<2>(ksteps, taps, epi, pair, stats);
    case 3: return pick_ksteps<3>(ksteps, taps, epi, pair, stats);
  }
  return nullptr;
}
int epi_kind(const ConvParams& p) {
  return p.out_nchw ? kEpiNchw : (p.epi_tma ? (p.mode == 1 ? kEpiStagedS2 : (p.kwm ? kEpiStagedKW : kEpiStaged)) : kEpiDirect);
}
}  // namespace

int conv_launch_prepared(const ConvParams& p, int grid, size_t smem_bytes, cudaStream_t stream) {
  static DeviceOnce attr_once;
  if (attr_once.run([]() {
        for (int mb = 1; mb <= 3; ++mb)
          for (int ks = 1; ks <= 4; ks *= 2)
            for (int epi = 0; epi < 5; ++epi)
              for (int taps = 1; taps <= 9; taps += 8)
                for (int pair = 0; pair < 2; ++pair)
                  for (int stats = 0; stats < 2; ++stats) {
                    if (pair && epi != kEpiStaged) continue;
                    ConvKernel k = pick_kernel(mb, ks, taps, epi, pair != 0, stats != 0);
                    if (!k) continue;
                    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
                    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
                  }
        return 0;
      }))
    return 1;
  if (grid <= 0) return 0;
  ConvKernel kern = pick_kernel(p.mb, p.ck / 16, p.taps, epi_kind(p), p.pair != 0, p.stats != nullptr);
  if (!kern) { set_error("conv: no kernel for mb %d ck %d taps %d nchw %d", p.mb, p.ck, p.taps, p.out_nchw); return 1; }
  cudaError_t e;
  if (p.pair || p.pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(p.kwm ? kThreadsKW : kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (p.pair) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = 2;
      attr[na].val.clusterDim.y = 1;
      attr[na].val.clusterDim.z = 1;
      ++na;
    }
    if (p.pdl) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    e = cudaLaunchKernelEx(&cfg, kern, p);
    if (e != cudaSuccess) { set_error("conv_tc_kernel (attributed) launch: %s", cudaGetErrorString(e)); return 1; }
  } else {
    kern<<<grid, p.kwm ? kThreadsKW : kThreads, smem_bytes, stream>>>(p);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("conv_tc_kernel launch: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

int conv_launch(const ConvSpec& spec, cudaStream_t stream) {
  ConvParams p;
  int grid = 0;
  size_t smem = 0;
  if (conv_prepare(spec, &p, &grid, &smem)) return 1;
  return conv_launch_prepared(p, grid, smem, stream);
}

}  // namespace stl

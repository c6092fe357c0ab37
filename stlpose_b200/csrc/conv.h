// Convolution engine: problem description (host) and kernel parameters (device).
//
// Activation layout ("padded-linear NHWC", bf16): a tensor of N images, H x W pixels, C channels is stored
// as N*(H+1)*(W+1) pixels of C contiguous channels; local row H and local column W of every image are zero.
// One shared zero row/column is enough for 3x3/pad-1 convolutions: pixel q = (n*(H+1)+h)*(W+1)+w has its
// 9 neighbours at q + (kh-1)*(W+1) + (kw-1), and every neighbour that falls outside the image lands on a zero
// cell (or outside the tensor, where TMA zero-fills).  A stride-1 conv is then out[q] = sum_t W_t . in[q+off_t]
// over the flat pixel index, for any batch and any tile boundary.
#pragma once
#include <mutex>
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace stl {

constexpr int kMaxUp = 3;
constexpr int kMaxStages = 12;

struct PaddedGeom {
  int N, H, W, C;
  __host__ __device__ int Hp() const { return H + 1; }
  __host__ __device__ int Wp() const { return W + 1; }
  __host__ __device__ long long pixels() const { return (long long)N * (H + 1) * (W + 1); }
  __host__ __device__ size_t bytes() const { return (size_t)pixels() * C * 2; }
};

// Host-side description of one fused convolution.
struct ConvSpec {
  const __nv_bfloat16* in = nullptr;   // padded-linear NHWC, Cin channels
  PaddedGeom in_geom{};
  // optional second input of a 1x1 convolution (same N, H, W; in2_C channels): the K dimension is the concatenation
  // [in | in2] and the packed weights have in_geom.C + in2_C columns.  Two 1x1 convolutions that are added before the
  // activation (Bottleneck conv3 + downsample, HRnet.py:88-101) become ONE launch whose sum stays in the fp32 accumulator
  const __nv_bfloat16* in2 = nullptr;
  int in2_C = 0;
  void* out = nullptr;                 // padded-linear NHWC bf16 (out_nchw=0) or fp32 NCHW (out_nchw=1)
  int cout = 0;                        // real output channels
  int cout_pad = 0;                    // multiple of 16; packed weights/bias have this many rows
  int ksize = 1;                       // 1 or 3 (pad = ksize/2)
  int stride = 1;                      // 1 or 2
  const __nv_bfloat16* weights = nullptr;  // [taps][cout_pad][cin] bf16 (BN scale folded)
  const float* bias = nullptr;             // [cout_pad] fp32 (BN shift folded)
  const __nv_bfloat16* residual = nullptr; // same geometry as out, added before ReLU (may alias out)
  int n_up = 0;                            // nearest-upsampled addends (HRNet fuse layers, HRnet.py:198-209)
  const __nv_bfloat16* up_src[kMaxUp] = {nullptr, nullptr, nullptr};
  int up_shift[kMaxUp] = {0, 0, 0};        // log2 of the upsampling factor
  int relu = 0;
  int out_nchw = 0;                        // 1: write fp32 [N][cout][H][W] (heatmap head)
  // tuning / debugging knobs (0 = let the engine decide)
  int force_tap_reload = 0;                // 1: one aligned TMA load per filter tap instead of shifted descriptors
  int force_mb = 0;
  int kw_merge = 0;                        // 1: kw-merged MMA shape where supported, -1: never, 0: STLPOSE_KW_MERGE=1 decides
  int max_ctas = 0;
  int pdl = 0;                             // 1: programmatic dependent launch (weights and bias must not be produced
                                           //    by the preceding kernel in the stream; activations may be)
  int img_lo = 0, img_hi = 0;              // stride-1 convs only: restrict the launch to images [img_lo, img_hi) (0,0 = all)
  void* dbg_counters = nullptr;            // optional [grid][3][4] int64 cycle counters (measurement aid)
  // training: per-channel sum / sum of squares of the stored (bf16-rounded) outputs over the valid pixels, accumulated
  // in the epilogue (BatchNorm batch statistics, HRnet.py:48-59 under model.train()).  [kStatsMaxRows][2][cout_pad]
  // floats; the launch fills one row per CTA (conv_stats_rows() of them), a fixed-order second pass adds the rows.
  float* stats = nullptr;
  // ... and, when stats_ticket is set, the LAST CTA to finish adds the rows in a fixed order and finalises the BatchNorm
  // statistics itself (mean, rstd, running statistics with momentum and the unbiased variance), so that no separate
  // launch is needed between the convolution and the normalisation.  stats_ticket: a device word that is zero on entry
  // and left zero.
  unsigned* stats_ticket = nullptr;
  float stats_count = 0.f, stats_eps = 0.f, stats_momentum = 0.f;
  float *stats_mean = nullptr, *stats_rstd = nullptr, *stats_run_mean = nullptr, *stats_run_var = nullptr;
};
constexpr int kStatsMaxRows = 160;           // >= CTAs of any launch (one per SM)

// Exact division of a 31-bit unsigned value by a small constant: q = (x * mul) >> shift  (mul < 2^32, 64-bit product).
struct FastDiv {
  uint32_t d, mul, shift;
  __host__ void init(uint32_t div) {
    d = div;
    uint32_t s = 0;
    while ((1u << s) < div) ++s;
    shift = 31 + s;
    mul = (uint32_t)((((unsigned long long)1 << shift) + div - 1) / div);
  }
  __host__ __device__ __forceinline__ uint32_t div(uint32_t x) const {
    return (uint32_t)(((unsigned long long)x * mul) >> shift);
  }
};

// Kernel parameters (passed __grid_constant__).
struct ConvParams {
  CUtensorMap tmA;
  CUtensorMap tmA2;  // second input tensor (ConvSpec::in2): K chunks >= n_chunks_a are loaded from it
  CUtensorMap tmB;
  CUtensorMap tmO;   // output tensor   (flat mode, bf16 out): epilogue stores whole panels with TMA
  CUtensorMap tmR;   // residual tensor (same geometry): panels are pre-loaded into the staging buffer
  int mode;        // 0: stride-1 flat-pixel tiles, 1: stride-2 structured tiles
  int taps;        // 1 or 9
  int n_chunks;    // Cin / ck (both inputs)
  int n_chunks_a;  // chunks that come from tmA (= n_chunks without a second input)
  int ck;          // channels per K chunk: 16, 32 or 64 (row = 2*ck bytes = swizzle span)
  int ksteps_last; // MMA K steps issued for the LAST chunk (ck / 16 unless its TMA box reaches past the channel count)
  int nt;          // UMMA N
  int n_ntiles;
  int mb;          // 128-row accumulator blocks per tile
  int kwm;         // 1: the three taps of a filter row share one MMA (N = 3*nt); the epilogue adds the three column
                   //    groups at row shifts 2 / 1 / 0 (conv_tc.cu, "kw-merged" mode).  mb = 1, 126 outputs per tile
  int tile_rows;   // output pixels per tile: 128*mb, or 126 in kw-merged mode
  uint32_t xch_off;  // kw-merged mode: byte offset (from the activation stages) of the cross-warp row exchange area
  int a_shift;     // 1: halo'd tile loaded once per chunk, taps addressed by shifted descriptors
  int halo;        // rows in front of the tile in shift mode (Wp+1 for 3x3, 0 for 1x1)
  int a_box_rows, a_pieces;
  int a_stages, b_stages;
  int b_taps;             // filter taps per streamed weight stage: 1, or 3 (a whole filter row per TMA box)
  int pair;        // 1: launched as clusters of two CTAs sharing cta_group::2 MMAs (M = 256)
  int n_mma;       // MMA-issuing warps (2 in burst mode with >= 2 blocks per tile)
  int b_resident;  // 1: all weight tiles of the layer stay in shared memory for the whole kernel
  int epi_tma;     // 1: epilogue stages 128-row x panel_ch panels in shared memory and moves them with TMA
  int panel_ch;    // channels per staged panel (64, or nt when nt is not a multiple of 64)
  int panel_swz;   // swizzle span of a panel row in bytes (128, 64) or 0 for none
  uint32_t epi_base_off;  // byte offset of the staging panels from the start of the activation stages
  int epi_batch;          // panels per staging buffer (a whole tile when panels are small)
  uint32_t epi_panel_bytes;
  uint32_t a_stage_bytes, b_stage_bytes, a_tx_bytes, b_tx_bytes, b_resident_bytes, b_bytes_total;
  int n_accbuf;
  uint32_t ctl_bytes;     // barriers + bias in front of the activation stages (multiple of 1024)
  uint32_t tmem_cols;
  long long total_tiles;
  // input / output geometry
  int in_Wp;
  int N, H, W, Hp, Wp;   // OUTPUT geometry
  long long P;           // output padded pixel count (end of the processed image range)
  int q_lo;              // first padded pixel of the processed image range (flat mode)
  FastDiv fd_Wp, fd_Hp, fd_bw, fd_bh;  // exact dividers for the epilogue's row -> (n, h, w) decode
  int bw, bh, bn, tiles_w, tiles_h, tiles_n;  // structured tiles (mode 1)
  // epilogue
  void* out;
  const float* bias;
  const __nv_bfloat16* residual;
  const __nv_bfloat16* up_src[kMaxUp];
  int up_shift[kMaxUp];
  int n_up;
  int relu;
  int out_nchw;
  int cout, cout_pad;
  int pdl;                // launched with programmatic stream serialization
  int dbg_skip_epilogue;  // measurement only
  float* stats;           // see ConvSpec::stats (null: off)
  unsigned* stats_ticket; // see ConvSpec::stats_ticket (null: rows only)
  float stats_count, stats_eps, stats_momentum;
  float *stats_mean, *stats_rstd, *stats_run_mean, *stats_run_var;
  long long* dbg_counters;  // measurement only: [grid][3 roles][4] cycle counters, or null
};

// Fills ConvParams (tensor maps included) and returns the launch configuration. 0 on success.
int conv_prepare(const ConvSpec& spec, ConvParams* p, int* grid, size_t* smem_bytes);
int conv_launch_prepared(const ConvParams& p, int grid, size_t smem_bytes, cudaStream_t stream);
int conv_launch(const ConvSpec& spec, cudaStream_t stream);

// Fused BasicBlock (block_tc.cu): y = relu(conv2(relu(conv1(x) + b1)) + b2 + x), both 3x3 / stride 1 / 32 -> 32 channels.
bool basic_block_supported(int H, int W, int C);
int basic_block_launch(const __nv_bfloat16* x, __nv_bfloat16* y, const __nv_bfloat16* w1, const float* b1,
                       const __nv_bfloat16* w2, const float* b2, int N, int H, int W, int max_ctas, cudaStream_t stream,
                       int pdl = 0);   // pdl: programmatic dependent launch (parameters must not come from the previous kernel)

// Junction of two layer1 Bottlenecks (link_tc.cu): out = relu(conv3(t) + b3 + x), a = relu(conv1'(out) + b1'), both 1x1
// (64 -> 256 -> 64 channels); `out` crosses HBM once instead of being written and read back.  With a second input t2
// instead of the residual x: out = relu([W3 | Wd] . [t | t2] + b3), the conv3 + downsample pair of layer1.0.
bool bottleneck_link_supported(int ct, int co, int ca);
int bottleneck_link_launch(const __nv_bfloat16* t, const __nv_bfloat16* t2, const __nv_bfloat16* x, __nv_bfloat16* out,
                           __nv_bfloat16* a, const __nv_bfloat16* w3, const float* b3, const __nv_bfloat16* w1,
                           const float* b1, int N, int H, int W, int max_ctas, cudaStream_t stream, int pdl = 0);

// Reference CUDA-core direct convolution with the same fused epilogue (validation only; slow).
int conv_launch_naive(const ConvSpec& spec, cudaStream_t stream);

void set_error(const char* fmt, ...);


// Per-device state of the launchers.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device and
// the SM count differs per device, so "done once" flags and cached counts are indexed by device; a mutex makes the
// first use safe from autograd worker threads.
struct DeviceOnce {
  static constexpr int kMaxDevices = 64;
  std::mutex mu;
  bool done[kMaxDevices] = {};
  // runs fn() (-> 0 on success) the first time it is called for the current device
  template <typename F>
  int run(F&& fn) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return fn();
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev]) return 0;
    const int r = fn();
    if (r == 0) done[dev] = true;
    return r;
  }
};
inline int device_sm_count() {
  static std::mutex mu;
  static int count[DeviceOnce::kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= DeviceOnce::kMaxDevices) return 148;
  std::lock_guard<std::mutex> lock(mu);
  if (!count[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    count[dev] = n > 0 ? n : 148;
  }
  return count[dev];
}

}  // namespace stl

// extern "C" surface of libstlpose_b200.so (declared in include/stlpose_b200.h).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/stlpose_b200.h"
#include "aux_kernels.h"
#include "conv.h"
#include "plan.h"
#include "pose_kernels.h"
#include "train_kernels.h"

namespace stl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

static int have_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device: stlpose_b200 has no CPU fallback");
    return 0;
  }
  return 1;
}

}  // namespace stl

using namespace stl;

struct stl_plan {
  Plan* impl;
};

extern "C" {

int stl_abi_version(void) { return STL_ABI_VERSION; }
const char* stl_last_error(void) { return g_err; }

int stl_flip_avg(const float* heat, const float* heat_flipped, float* out, int B, int J, int h, int w,
                 const int* pairs_host, int n_pairs, void* stream) {
  if (!have_device()) return 1;
  if (!heat || !heat_flipped || !out) { set_error("stl_flip_avg: null pointer"); return 1; }
  return flip_avg(heat, heat_flipped, out, B, J, h, w, pairs_host, n_pairs, (cudaStream_t)stream);
}

int stl_flip_back(const float* in, float* out, int B, int J, int h, int w, const int* pairs_host, int n_pairs,
                  void* stream) {
  if (!have_device()) return 1;
  if (!in || !out) { set_error("stl_flip_back: null pointer"); return 1; }
  return flip_back(in, out, B, J, h, w, pairs_host, n_pairs, (cudaStream_t)stream);
}

int stl_decode(const float* heat, const float* heat_flipped, const float* center, const float* scale, int B, int J,
               int h, int w, const int* pairs_host, int n_pairs, int refine, float* avg_out, float* preds,
               float* maxvals, float* coords, void* stream) {
  if (!have_device()) return 1;
  if (B > 0 && (!heat || !maxvals || !coords)) { set_error("stl_decode: null pointer"); return 1; }
  if (preds && (!center || !scale)) { set_error("stl_decode: preds requested without center/scale"); return 1; }
  return decode(heat, heat_flipped, center, scale, B, J, h, w, pairs_host, n_pairs, refine, avg_out, preds, maxvals,
                coords, (cudaStream_t)stream);
}

size_t stl_mse_workspace_bytes(void) { return mse_workspace_bytes(); }

int stl_oks_nms(const float* keypoints, const double* area, const double* box_score, const int* image_offsets,
                int n_images, int max_persons_per_image, int J, const double* vars, float in_vis_thr, double oks_thr,
                float nms_vis_thr, int rescore, double* score_out, int* keep_rank, void* stream) {
  if (!have_device()) return 1;
  if (n_images > 0 && (!keypoints || !area || !box_score || !image_offsets || !vars || !score_out || !keep_rank)) {
    set_error("stl_oks_nms: null pointer");
    return 1;
  }
  if (J < 1 || J > kMaxJoints) { set_error("stl_oks_nms: 1..%d joints supported (got %d)", kMaxJoints, J); return 1; }
  return oks_nms(keypoints, area, box_score, image_offsets, n_images, max_persons_per_image, J, vars, in_vis_thr, oks_thr, nms_vis_thr, rescore,
                 score_out, keep_rank, (cudaStream_t)stream);
}

int stl_generate_target(const double* joints, const double* joints_vis, const float* joints_weight, int B, int J, int h,
                        int w, int image_h, int image_w, int sigma, float* target, float* target_weight, void* stream) {
  if (!have_device()) return 1;
  if (B > 0 && (!joints || !joints_vis || !target || !target_weight)) { set_error("stl_generate_target: null pointer"); return 1; }
  return generate_target(joints, joints_vis, joints_weight, B, J, h, w, image_h, image_w, sigma, target, target_weight,
                         (cudaStream_t)stream);
}

int stl_upsampled_argmax(const float* heat, int B, int J, int h, int w, int out_h, int out_w, float* coords,
                         float* maxvals, void* stream) {
  if (!have_device()) return 1;
  if (B > 0 && (!heat || !coords || !maxvals)) { set_error("stl_upsampled_argmax: null pointer"); return 1; }
  return upsampled_argmax(heat, B, J, h, w, out_h, out_w, coords, maxvals, (cudaStream_t)stream);
}

int stl_warp_affine_crops(const void* img_u8_hwc, int img_h, int img_w, const double* minv, int N, int out_h, int out_w,
                          void* out_u8_nchw, float* out_f32_nchw, const float* mean3_host, const float* std3_host,
                          void* stream) {
  if (!have_device()) return 1;
  if (N > 0 && (!img_u8_hwc || !minv || (!out_u8_nchw && !out_f32_nchw))) {
    set_error("stl_warp_affine_crops: null pointer");
    return 1;
  }
  return warp_affine_crops((const uint8_t*)img_u8_hwc, img_h, img_w, minv, N, out_h, out_w, (uint8_t*)out_u8_nchw,
                           out_f32_nchw, mean3_host, std3_host, (cudaStream_t)stream);
}

int stl_warp_affine_crops_f32(const float* img_f32_hwc, int img_h, int img_w, const double* minv, int N, int out_h,
                              int out_w, float* out_f32_nchw, void* stream) {
  if (!have_device()) return 1;
  if (N > 0 && (!img_f32_hwc || !minv || !out_f32_nchw)) { set_error("stl_warp_affine_crops_f32: null pointer"); return 1; }
  return warp_affine_crops_f32(img_f32_hwc, img_h, img_w, minv, N, out_h, out_w, out_f32_nchw, (cudaStream_t)stream);
}

int stl_sgd_step_batched(const void* items, const int* block_offsets, int n_items, int total_blocks, float lr,
                         float momentum, float weight_decay, int nesterov, void* stream) {
  if (!have_device()) return 1;
  if (n_items > 0 && (!items || !block_offsets)) { set_error("stl_sgd_step_batched: null pointer"); return 1; }
  static_assert(sizeof(SgdItem) == 32, "stl_sgd_step_batched item layout");
  return sgd_step_batched(reinterpret_cast<const SgdItem*>(items), block_offsets, n_items, total_blocks, lr, momentum,
                          weight_decay, nesterov, (cudaStream_t)stream);
}

int stl_stem_im2col(const float* x_nchw, void* rows, int N, int H, int W, void* stream) {
  if (!have_device()) return 1;
  if (N > 0 && (!x_nchw || !rows)) { set_error("stl_stem_im2col: null pointer"); return 1; }
  if (N <= 0) return 0;
  if ((H | W) & 1) { set_error("stl_stem_im2col: even H, W required"); return 1; }
  return stem_im2col(x_nchw, (__nv_bfloat16*)rows, N, N, H, W, (cudaStream_t)stream);
}

int stl_pck_accuracy(const float* pred_coords, const float* target_coords, int B, int J, int h, int w, float thr,
                     float* acc, float* avg_acc, int* cnt, void* stream) {
  if (!have_device()) return 1;
  if (!pred_coords || !target_coords || !acc || !avg_acc || !cnt) { set_error("stl_pck_accuracy: null pointer"); return 1; }
  return pck_accuracy(pred_coords, target_coords, B, J, h, w, thr, acc, avg_acc, cnt, (cudaStream_t)stream);
}

int stl_scale_inplace(float* x, const float* scale_dev, long long n, void* stream) {
  if (!have_device()) return 1;
  if (!x || !scale_dev) { set_error("stl_scale_inplace: null pointer"); return 1; }
  return scale_inplace(x, scale_dev, n, (cudaStream_t)stream);
}

int stl_mse_loss_fwd_bwd(const float* out, const float* tgt, const float* tw, int B, int J, int hw, float* loss,
                         float* grad, void* workspace, void* stream) {
  if (!have_device()) return 1;
  if (!out || !tgt || !tw || !loss || !workspace) { set_error("stl_mse_loss_fwd_bwd: null pointer"); return 1; }
  if (B <= 0 || J <= 0 || hw <= 0) { set_error("stl_mse_loss_fwd_bwd: empty problem"); return 1; }
  return mse_loss(out, tgt, tw, B, J, hw, loss, grad, workspace, (cudaStream_t)stream);
}

size_t stl_padded_bytes(int N, int C, int H, int W) { return PaddedGeom{N, H, W, C}.bytes(); }

int stl_nchw_to_padded(const float* x, void* y, int N, int C, int H, int W, int C_pad, void* stream) {
  if (!have_device()) return 1;
  return nchw_to_padded(x, reinterpret_cast<__nv_bfloat16*>(y), N, C, H, W, C_pad, (cudaStream_t)stream);
}

int stl_padded_to_nchw(const void* y, float* x, int N, int C, int H, int W, int C_pad, void* stream) {
  if (!have_device()) return 1;
  return padded_to_nchw(reinterpret_cast<const __nv_bfloat16*>(y), x, N, C, H, W, C_pad, (cudaStream_t)stream);
}

int stl_pack_conv_weights(const float* w, const float* g, const float* b, const float* m, const float* v,
                          const float* cbias, float eps, int Cout, int Cin, int ksize, int Cout_pad, int Cin_pad,
                          void* w_packed, float* bias_packed, void* stream) {
  if (!have_device()) return 1;
  if (g && (!b || !m || !v)) { set_error("stl_pack_conv_weights: incomplete BatchNorm parameters"); return 1; }
  return pack_weights(w, g, b, m, v, cbias, eps, Cout, Cin, ksize, Cout_pad, Cin_pad,
                      reinterpret_cast<__nv_bfloat16*>(w_packed), bias_packed, (cudaStream_t)stream);
}

int stl_basic_block(const void* x, void* y, const void* w1_packed, const float* bias1, const void* w2_packed,
                    const float* bias2, int N, int H, int W, int C, void* stream) {
  if (!have_device()) return 1;
  if (!x || !y || !w1_packed || !bias1 || !w2_packed || !bias2) { set_error("stl_basic_block: null pointer"); return 1; }
  if (!basic_block_supported(H, W, C)) { set_error("stl_basic_block: unsupported shape C=%d %dx%d", C, H, W); return 1; }
  return basic_block_launch((const __nv_bfloat16*)x, (__nv_bfloat16*)y, (const __nv_bfloat16*)w1_packed, bias1,
                            (const __nv_bfloat16*)w2_packed, bias2, N, H, W, 0, (cudaStream_t)stream);
}

int stl_bottleneck_link(const void* t, const void* x, void* out, void* a, const void* w3_packed, const float* bias3,
                        const void* w1n_packed, const float* bias1n, int N, int H, int W, int max_ctas, void* stream) {
  if (!have_device()) return 1;
  typedef const __nv_bfloat16* P;
  return bottleneck_link_launch((P)t, nullptr, (P)x, (__nv_bfloat16*)out, (__nv_bfloat16*)a, (P)w3_packed, bias3,
                                (P)w1n_packed, bias1n, N, H, W, max_ctas, (cudaStream_t)stream);
}

int stl_bottleneck_link2(const void* t, const void* t2, void* out, void* a, const void* w3cat_packed, const float* bias3,
                         const void* w1n_packed, const float* bias1n, int N, int H, int W, int max_ctas, void* stream) {
  if (!have_device()) return 1;
  typedef const __nv_bfloat16* P;
  return bottleneck_link_launch((P)t, (P)t2, nullptr, (__nv_bfloat16*)out, (__nv_bfloat16*)a, (P)w3cat_packed, bias3,
                                (P)w1n_packed, bias1n, N, H, W, max_ctas, (cudaStream_t)stream);
}

int stl_pack_conv_weights_dgrad(const float* w, int Cout, int Cin, int ksize, int Rows_pad, int K_pad, void* w_packed,
                                float* bias_packed, void* stream) {
  if (!have_device()) return 1;
  if (!w || !w_packed) { set_error("stl_pack_conv_weights_dgrad: null pointer"); return 1; }
  return pack_weights_dgrad(w, Cout, Cin, ksize, Rows_pad, K_pad, reinterpret_cast<__nv_bfloat16*>(w_packed),
                            bias_packed, (cudaStream_t)stream);
}

int stl_pack_conv_weights_batched(const stl_pack_item* items_dev, const int* block_offsets_dev, int n_items,
                                  int total_blocks, void* stream) {
  if (!have_device()) return 1;
  if (n_items > 0 && (!items_dev || !block_offsets_dev)) { set_error("stl_pack_conv_weights_batched: null pointer"); return 1; }
  static_assert(sizeof(stl_pack_item) == sizeof(PackItem), "stl_pack_item and PackItem must have the same layout");
  return pack_weights_batched(reinterpret_cast<const PackItem*>(items_dev), block_offsets_dev, n_items, total_blocks,
                              (cudaStream_t)stream);
}

struct BnFused {   // stl_conv2d_bn: the convolution's last CTA finalises the BatchNorm statistics
  unsigned* ticket;
  float eps, momentum;
  float *mean, *rstd, *run_mean, *run_var;
};
static int conv2d_impl(const stl_conv_desc* d, float* stats, int* stats_rows, void* stream, const BnFused* bn = nullptr);

int stl_conv2d(const stl_conv_desc* d, void* stream) { return conv2d_impl(d, nullptr, nullptr, stream); }

size_t stl_conv2d_stats_floats(int cout_pad) { return (size_t)kStatsMaxRows * 2 * (size_t)cout_pad; }

int stl_conv2d_stats(const stl_conv_desc* d, float* stats, int* stats_rows, void* stream) {
  if (!stats || !stats_rows) { set_error("stl_conv2d_stats: null pointer"); return 1; }
  return conv2d_impl(d, stats, stats_rows, stream);
}

int stl_conv2d_bn(const stl_conv_desc* d, float* stats, unsigned* ticket, float eps, float momentum, float* mean,
                  float* rstd, float* running_mean, float* running_var, int* done, void* stream) {
  if (!stats || !ticket || !mean || !rstd || !done) { set_error("stl_conv2d_bn: null pointer"); return 1; }
  BnFused bn{ticket, eps, momentum, mean, rstd, running_mean, running_var};
  int rows = 0;
  *done = 0;
  if (conv2d_impl(d, stats, &rows, stream, &bn)) return 1;
  *done = rows > 0 ? 1 : 0;
  return 0;
}

int stl_bn_apply(const void* z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                 const void* residual, int relu, int N, int H, int W, int C, void* y, void* stream) {
  if (!have_device()) return 1;
  if (!z || !mean || !rstd || !gamma || !beta || !y) { set_error("stl_bn_apply: null pointer"); return 1; }
  typedef const __nv_bfloat16* P;
  return bn_apply((P)z, mean, rstd, gamma, beta, (P)residual, relu, N, H, W, C, (__nv_bfloat16*)y, (cudaStream_t)stream);
}

static int conv2d_impl(const stl_conv_desc* d, float* stats, int* stats_rows, void* stream, const BnFused* bn) {
  if (!have_device()) return 1;
  if (!d || !d->in || !d->out || !d->w_packed || !d->bias_packed) { set_error("stl_conv2d: null pointer"); return 1; }
  if (d->n_up < 0 || d->n_up > STL_MAX_UP) { set_error("stl_conv2d: n_up out of range"); return 1; }
  ConvSpec s;
  s.in = reinterpret_cast<const __nv_bfloat16*>(d->in);
  s.in_geom = PaddedGeom{d->N, d->H, d->W, d->Cin};
  s.out = d->out;
  s.cout = d->Cout;
  s.cout_pad = d->Cout_pad;
  s.ksize = d->ksize;
  s.stride = d->stride;
  s.weights = reinterpret_cast<const __nv_bfloat16*>(d->w_packed);
  s.bias = d->bias_packed;
  s.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual);
  s.n_up = d->n_up;
  for (int i = 0; i < d->n_up; ++i) {
    s.up_src[i] = reinterpret_cast<const __nv_bfloat16*>(d->up_src[i]);
    s.up_shift[i] = d->up_shift[i];
  }
  s.relu = d->relu;
  s.out_nchw = d->out_nchw;
  s.force_tap_reload = d->impl == 1;
  s.kw_merge = d->impl == 3 ? 1 : 0;
  s.force_mb = d->force_mb;
  s.max_ctas = d->max_ctas;
  s.dbg_counters = d->dbg_counters;
  s.pdl = d->pdl ? 1 : 0;
  if (d->in2) {
    if (d->impl != 0 || d->Cin2 <= 0) { set_error("stl_conv2d: a second input needs impl 0 and Cin2 > 0"); return 1; }
    s.in2 = reinterpret_cast<const __nv_bfloat16*>(d->in2);
    s.in2_C = d->Cin2;
  }
  if (stats_rows) *stats_rows = 0;
  if (d->impl == 2) return conv_launch_naive(s, (cudaStream_t)stream);
  if (!stats) return conv_launch(s, (cudaStream_t)stream);
  s.stats = stats;
  if (bn) {
    s.stats_ticket = bn->ticket;
    s.stats_count = (float)((long long)d->N * (d->H / d->stride) * (d->W / d->stride));
    s.stats_eps = bn->eps; s.stats_momentum = bn->momentum;
    s.stats_mean = bn->mean; s.stats_rstd = bn->rstd; s.stats_run_mean = bn->run_mean; s.stats_run_var = bn->run_var;
  }
  ConvParams p;
  int grid = 0;
  size_t smem = 0;
  if (conv_prepare(s, &p, &grid, &smem)) return 1;
  if (grid > kStatsMaxRows) { p.stats = nullptr; p.stats_ticket = nullptr; }
  if (conv_launch_prepared(p, grid, smem, (cudaStream_t)stream)) return 1;
  *stats_rows = (p.stats && (!bn || p.stats_ticket)) ? grid : 0;   // 0: no fused statistics for this shape: run the separate reduction
  return 0;
}

stl_plan* stl_plan_create(const stl_hrnet_cfg* cfg) {
  if (!cfg) { set_error("stl_plan_create: null cfg"); return nullptr; }
  Plan* p = Plan::create(*cfg);
  if (!p) return nullptr;
  stl_plan* h = new stl_plan;
  h->impl = p;
  const char* e = getenv("STLPOSE_TAP_RELOAD");
  if (e && e[0] == '1') p->tap_reload = 1;
  return h;
}

void stl_plan_destroy(stl_plan* plan) {
  if (!plan) return;
  delete plan->impl;
  delete plan;
}

int stl_plan_num_convs(const stl_plan* plan) { return plan ? (int)plan->impl->layers.size() : 0; }

int stl_plan_conv_info(const stl_plan* plan, int index, stl_conv_info* info) {
  if (!plan || !info || index < 0 || index >= (int)plan->impl->layers.size()) {
    set_error("stl_plan_conv_info: bad arguments");
    return 1;
  }
  const Plan::Layer& L = plan->impl->layers[index];
  memset(info, 0, sizeof(*info));
  snprintf(info->conv_key, sizeof info->conv_key, "%s", L.conv_key.c_str());
  snprintf(info->bn_key, sizeof info->bn_key, "%s", L.bn_key.c_str());
  info->cout = L.cout; info->cin = L.cin; info->ksize = L.k; info->stride = L.stride;
  return 0;
}

size_t stl_plan_weight_bytes(const stl_plan* plan) { return plan ? plan->impl->weight_bytes : 0; }

int stl_plan_pack_conv(stl_plan* plan, int index, const float* w, const float* g, const float* b, const float* m,
                       const float* v, const float* cbias, float eps, void* arena, void* stream) {
  if (!have_device()) return 1;
  if (!plan || !w || !arena) { set_error("stl_plan_pack_conv: null pointer"); return 1; }
  if (g && (!b || !m || !v)) { set_error("stl_plan_pack_conv: incomplete BatchNorm parameters"); return 1; }
  return plan->impl->pack_conv(index, w, g, b, m, v, cbias, eps, arena, (cudaStream_t)stream);
}

size_t stl_plan_workspace_bytes(const stl_plan* plan, int n_images) {
  return plan && n_images > 0 ? plan->impl->workspace_bytes(n_images) : 0;
}

int stl_plan_forward(stl_plan* plan, const float* x, int B, int flip_pair, float* heat, const void* arena,
                     void* workspace, size_t ws_bytes, void* stream) {
  if (!have_device()) return 1;
  if (!plan || !x || !heat || !arena || !workspace) { set_error("stl_plan_forward: null pointer"); return 1; }
  return plan->impl->forward(x, B, flip_pair, heat, arena, workspace, ws_bytes, (cudaStream_t)stream);
}

int stl_plan_launches_per_forward(const stl_plan* plan) { return plan ? (int)plan->impl->ops.size() : 0; }
int stl_plan_kernel_launches(const stl_plan* plan) {
  if (!plan) return 0;
  return plan->impl->bound ? (int)plan->impl->launches.size() : (int)plan->impl->ops.size();
}

int stl_plan_forward_timed(stl_plan* plan, const float* x, int B, int flip_pair, float* heat, const void* arena,
                           void* workspace, size_t ws_bytes, void* stream, float* op_ms_host) {
  if (!have_device()) return 1;
  if (!plan || !x || !heat || !arena || !workspace || !op_ms_host) {
    set_error("stl_plan_forward_timed: null pointer");
    return 1;
  }
  return plan->impl->forward(x, B, flip_pair, heat, arena, workspace, ws_bytes, (cudaStream_t)stream, op_ms_host);
}

int stl_plan_op_info(const stl_plan* plan, int op_index, stl_op_info* info) {
  if (!plan) { set_error("stl_plan_op_info: null plan"); return 1; }
  return plan->impl->op_info(op_index, info);
}

int stl_bn_train_forward(const void* z, const float* gamma, const float* beta, const void* residual, int relu,
                         float eps, float momentum, int N, int H, int W, int C, void* y, float* sums, float* mean,
                         float* rstd, float* running_mean, float* running_var, void* stream) {
  if (!have_device()) return 1;
  if (!z || !gamma || !beta || !y || !sums || !mean || !rstd) { set_error("stl_bn_train_forward: null pointer"); return 1; }
  typedef const __nv_bfloat16* P;
  return bn_train_forward((P)z, gamma, beta, (P)residual, relu, eps, momentum, N, H, W, C, (__nv_bfloat16*)y, sums, mean,
                          rstd, running_mean, running_var, nullptr, (cudaStream_t)stream);
}

int stl_bn_train_forward_ticket(const void* z, const float* gamma, const float* beta, const void* residual, int relu,
                                float eps, float momentum, int N, int H, int W, int C, void* y, float* sums, float* mean,
                                float* rstd, float* running_mean, float* running_var, unsigned* ticket, void* stream) {
  if (!have_device()) return 1;
  if (!z || !gamma || !beta || !y || !sums || !mean || !rstd || !ticket) {
    set_error("stl_bn_train_forward_ticket: null pointer");
    return 1;
  }
  typedef const __nv_bfloat16* P;
  return bn_train_forward((P)z, gamma, beta, (P)residual, relu, eps, momentum, N, H, W, C, (__nv_bfloat16*)y, sums, mean,
                          rstd, running_mean, running_var, ticket, (cudaStream_t)stream);
}

int stl_bn_train_forward_coop(const void* z, const float* gamma, const float* beta, const void* residual, int relu,
                              float eps, float momentum, int N, int H, int W, int C, void* y, float* sums, float* mean,
                              float* rstd, float* running_mean, float* running_var, unsigned* ticket, unsigned* sync,
                              void* stream) {
  if (!have_device()) return 1;
  if (!z || !gamma || !beta || !y || !sums || !mean || !rstd || !ticket || !sync) {
    set_error("stl_bn_train_forward_coop: null pointer");
    return 1;
  }
  typedef const __nv_bfloat16* P;
  return bn_train_forward_coop((P)z, gamma, beta, (P)residual, relu, eps, momentum, N, H, W, C, (__nv_bfloat16*)y, sums,
                               mean, rstd, running_mean, running_var, ticket, sync, (cudaStream_t)stream);
}

int stl_bn_train_backward_coop(const void* dy, const void* y, const void* z, const float* mean, const float* rstd,
                               const float* gamma, const float* beta, int relu, int N, int H, int W, int C, void* dz,
                               void* dres, float* dbeta_dgamma, float* workspace, unsigned* ticket, unsigned* sync,
                               void* stream) {
  if (!have_device()) return 1;
  if (!dy || !z || !mean || !rstd || !gamma || !dz || !dbeta_dgamma || !workspace || !ticket || !sync ||
      (relu == 1 && !y) || (relu == 2 && !beta)) {
    set_error("stl_bn_train_backward_coop: null pointer");
    return 1;
  }
  typedef const __nv_bfloat16* P;
  return bn_train_backward_coop((P)dy, (P)y, (P)z, mean, rstd, gamma, beta, relu, N, H, W, C, (__nv_bfloat16*)dz,
                                (__nv_bfloat16*)dres, dbeta_dgamma, workspace, ticket, sync, (cudaStream_t)stream);
}

int stl_bn_train_forward_fused(const void* z, const float* stat_rows, int rows, int c_pad, const float* gamma,
                               const float* beta, const void* residual, int relu, float eps, float momentum, int N, int H,
                               int W, int C, void* y, float* mean, float* rstd, float* running_mean, float* running_var,
                               void* stream) {
  if (!have_device()) return 1;
  if (!z || !stat_rows || !gamma || !beta || !y || !mean || !rstd) { set_error("stl_bn_train_forward_fused: null pointer"); return 1; }
  typedef const __nv_bfloat16* P;
  return bn_train_forward_fused((P)z, stat_rows, rows, c_pad, gamma, beta, (P)residual, relu, eps, momentum, N, H, W, C,
                                (__nv_bfloat16*)y, mean, rstd, running_mean, running_var, (cudaStream_t)stream);
}

int stl_bn_train_backward(const void* dy, const void* y, const void* z, const float* mean, const float* rstd,
                          const float* gamma, int relu, int N, int H, int W, int C, void* dz, void* dres, float* sums,
                          void* stream) {
  if (!have_device()) return 1;
  if (!dy || !z || !mean || !rstd || !gamma || !dz || !sums || (relu && !y)) {
    set_error("stl_bn_train_backward: null pointer");
    return 1;
  }
  typedef const __nv_bfloat16* P;
  return bn_train_backward((P)dy, (P)y, (P)z, mean, rstd, gamma, nullptr, relu ? 1 : 0, N, H, W, C, (__nv_bfloat16*)dz,
                           (__nv_bfloat16*)dres, sums, nullptr, nullptr, (cudaStream_t)stream);
}

int stl_bn_train_backward_ticket(const void* dy, const void* y, const void* z, const float* mean, const float* rstd,
                                 const float* gamma, int relu, int N, int H, int W, int C, void* dz, void* dres,
                                 float* dbeta_dgamma, float* workspace, unsigned* ticket, void* stream) {
  if (!have_device()) return 1;
  if (!dy || !z || !mean || !rstd || !gamma || !dz || !dbeta_dgamma || !workspace || !ticket || (relu && !y)) {
    set_error("stl_bn_train_backward_ticket: null pointer");
    return 1;
  }
  typedef const __nv_bfloat16* P;
  return bn_train_backward((P)dy, (P)y, (P)z, mean, rstd, gamma, nullptr, relu ? 1 : 0, N, H, W, C, (__nv_bfloat16*)dz,
                           (__nv_bfloat16*)dres, dbeta_dgamma, workspace, ticket, (cudaStream_t)stream);
}

int stl_bn_train_backward_ticket_z(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma,
                                   const float* beta, int N, int H, int W, int C, void* dz, float* dbeta_dgamma,
                                   float* workspace, unsigned* ticket, void* stream) {
  if (!have_device()) return 1;
  if (!dy || !z || !mean || !rstd || !gamma || !beta || !dz || !dbeta_dgamma || !workspace || !ticket) {
    set_error("stl_bn_train_backward_ticket_z: null pointer");
    return 1;
  }
  typedef const __nv_bfloat16* P;
  return bn_train_backward((P)dy, nullptr, (P)z, mean, rstd, gamma, beta, 2, N, H, W, C, (__nv_bfloat16*)dz, nullptr,
                           dbeta_dgamma, workspace, ticket, (cudaStream_t)stream);
}

int stl_sum_relu_forward(const void* const* same_host, int n_same, const void* const* up_host, const int* shift_host,
                         int n_up, void* y, int N, int H, int W, int C, void* stream) {
  if (!have_device()) return 1;
  if (!y || (n_same > 0 && !same_host) || (n_up > 0 && (!up_host || !shift_host))) {
    set_error("stl_sum_relu_forward: null pointer");
    return 1;
  }
  return sum_relu_forward(reinterpret_cast<const __nv_bfloat16* const*>(same_host), n_same,
                          reinterpret_cast<const __nv_bfloat16* const*>(up_host), shift_host, n_up, (__nv_bfloat16*)y, N,
                          H, W, C, (cudaStream_t)stream);
}

int stl_relu_mask(const void* dy, const void* y, void* g, long long elems, void* stream) {
  if (!have_device()) return 1;
  if (!dy || !y || !g || elems % 8) { set_error("stl_relu_mask: bad arguments"); return 1; }
  return relu_mask((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (__nv_bfloat16*)g, elems, (cudaStream_t)stream);
}

int stl_upsample_backward(const void* g, void* dlow, int N, int H, int W, int C, int shift, void* stream) {
  if (!have_device()) return 1;
  if (!g || !dlow || C % 8 || shift < 1 || shift > 3) { set_error("stl_upsample_backward: bad arguments"); return 1; }
  return upsample_backward((const __nv_bfloat16*)g, (__nv_bfloat16*)dlow, N, H, W, C, shift, (cudaStream_t)stream);
}

int stl_conv_dgrad(const void* dz, const void* w_packed, void* dx, int N, int Hi, int Wi, int Cin, int Cout, int ksize,
                   int stride, void* stream) {
  if (!have_device()) return 1;
  if (!dz || !w_packed || !dx) { set_error("stl_conv_dgrad: null pointer"); return 1; }
  return conv_dgrad_naive((const __nv_bfloat16*)dz, (const __nv_bfloat16*)w_packed, (__nv_bfloat16*)dx, N, Hi, Wi, Cin,
                          Cout, ksize, stride, (cudaStream_t)stream);
}

size_t stl_conv_wgrad_workspace_bytes(int N, int Hi, int Wi, int Cin, int Cout, int ksize, int stride, int cin_real) {
  if (!wgrad_tc_supported(Wi, Cin, Cout, cin_real, ksize, stride))
    return conv_wgrad_naive_workspace_bytes(N, Hi, Wi, Cin, Cout, ksize, stride, cin_real);
  return wgrad_tc_workspace_bytes(N, Hi, Wi, Cin, Cout, ksize, cin_real);
}

int stl_conv_wgrad(const void* x, const void* dz, float* dw, int N, int Hi, int Wi, int Cin, int Cout, int ksize,
                   int stride, int cin_real, void* workspace, size_t workspace_bytes, void* stream) {
  if (!have_device()) return 1;
  if (!x || !dz || !dw) { set_error("stl_conv_wgrad: null pointer"); return 1; }
  if (wgrad_tc_supported(Wi, Cin, Cout, cin_real, ksize, stride))
    return wgrad_tc_launch((const __nv_bfloat16*)x, (const __nv_bfloat16*)dz, dw, N, Hi, Wi, Cin, Cout, ksize, cin_real,
                           workspace, workspace_bytes, (cudaStream_t)stream);
  return conv_wgrad_naive((const __nv_bfloat16*)x, (const __nv_bfloat16*)dz, dw, N, Hi, Wi, Cin, Cout, ksize, stride,
                          cin_real, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t stl_bn_workspace_floats(int C) { return bn_workspace_floats(C); }

int stl_zero_stuff(const void* dz, void* u, int N, int H, int W, int C, void* stream) {
  if (!have_device()) return 1;
  if (!dz || !u) { set_error("stl_zero_stuff: null pointer"); return 1; }
  return zero_stuff((const __nv_bfloat16*)dz, (__nv_bfloat16*)u, N, H, W, C, (cudaStream_t)stream);
}

size_t stl_conv_wgrad_naive_workspace_bytes(int N, int Hi, int Wi, int Cin, int Cout, int ksize, int stride,
                                            int cin_real) {
  return conv_wgrad_naive_workspace_bytes(N, Hi, Wi, Cin, Cout, ksize, stride, cin_real);
}

int stl_conv_wgrad_naive(const void* x, const void* dz, float* dw, int N, int Hi, int Wi, int Cin, int Cout, int ksize,
                         int stride, int cin_real, void* workspace, size_t workspace_bytes, void* stream) {
  if (!have_device()) return 1;
  if (!x || !dz || !dw) { set_error("stl_conv_wgrad_naive: null pointer"); return 1; }
  return conv_wgrad_naive((const __nv_bfloat16*)x, (const __nv_bfloat16*)dz, dw, N, Hi, Wi, Cin, Cout, ksize, stride,
                          cin_real, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"

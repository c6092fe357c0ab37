/* stlpose_b200 -- C ABI of the B200-native HRNet keypoint hot path.
 *
 * Drop-in boundary for the top-down HRNet pipeline of angelvillar96/STLPose.  Each entry point names the
 * reference interface it replaces (paths relative to the reference's src/).  Conventions:
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - the caller owns all buffers; nothing is allocated behind the caller's back;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work on it and never synchronise;
 *   - return value 0 = OK, non-zero = error; stl_last_error() returns the message (thread-local);
 *   - there is no CPU fallback anywhere: without a CUDA device every compute entry point fails.
 */
#ifndef STLPOSE_B200_H_
#define STLPOSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STL_ABI_VERSION 1
#define STL_MAX_UP 3

int stl_abi_version(void);
const char* stl_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Heatmap post-processing (fp32 NCHW heatmaps [B][J][h][w]).
 * ---------------------------------------------------------------------------------------------- */

/* lib/transforms.py:147-164 flip_back + lib/inference.py:25-26 (1-px shift keeping column 0, average):
 *   out = 0.5 * (heat + shift_right_1(flip_W(heat_flipped)[:, swapped joints]))
 * pairs_host: n_pairs x 2 joint indices to swap (CONSTANTS.py:65 FLIP_PAIRS). */
int stl_flip_avg(const float* heat, const float* heat_flipped, float* out, int B, int J, int h, int w,
                 const int* pairs_host, int n_pairs, void* stream);

/* lib/transforms.py:147-164 flip_back alone: out[n,j,y,x] = in[n,swap(j),y,w-1-x]. */
int stl_flip_back(const float* in, float* out, int B, int J, int h, int w, const int* pairs_host, int n_pairs,
                  void* stream);

/* lib/pose_parsing.py:16-55 get_max_preds_hrnet (refine = 0) and :58-92 get_final_preds_hrnet (refine = 1,
 * center/scale/preds given).  When heat_flipped is non-null the flip-test average above is fused in front of
 * the argmax (and written to avg_out if non-null).
 *   coords  [B][J][2]  heat-map space (x, y), zeroed where maxval <= 0, +-0.25 px refinement when refine
 *   maxvals [B][J]
 *   preds   [B][J][2]  image space through the inverse crop affine (lib/transforms.py:184-240; rot = 0, only
 *                      scale[:,0] is used, as in the reference); may be null together with center/scale.
 *   center, scale: [B][2] fp32. */
int stl_decode(const float* heat, const float* heat_flipped, const float* center, const float* scale, int B, int J,
               int h, int w, const int* pairs_host, int n_pairs, int refine, float* avg_out, float* preds,
               float* maxvals, float* coords, void* stream);

/* lib/loss.py:61-94 PersonMSELoss.forward and its gradient:
 *   loss = 0.5/(J*B*hw) * sum (tw*(out-tgt))^2 ;  grad = tw^2*(out-tgt)/(J*B*hw)   (grad may be null)
 * out/tgt [B][J][hw] fp32, tw [B][J] fp32, loss: 1 float.  workspace: stl_mse_workspace_bytes() bytes. */
size_t stl_mse_workspace_bytes(void);
/* OKS rescoring + greedy OKS-NMS: generate_submission_hrnet (lib/metrics.py:232-258) and
 * nms.oks_nms / oks_iou (lib/nms.py:10-74) for all images of an evaluation at once.  Persons are grouped by image:
 * persons [image_offsets[i], image_offsets[i+1]) belong to image i (any number, like lib/nms.py: pass the largest
 * count as max_persons_per_image, it sizes the kernel's shared memory; ~15 000 persons fit).
 * keypoints [M][J][3] fp32 (x, y, score; J <= 64), area / box_score [M] fp64, vars [J] fp64 = (2*sigma_j)^2.
 * rescore != 0: score = mean(joint scores > in_vis_thr) * box_score, else box_score is the score.  nms_vis_thr < 0: all
 * joints enter the OKS (what the reference's call does); otherwise only joints of the candidate with score > nms_vis_thr.
 * Outputs: score_out [M] fp64, keep_rank [M] (position in the image's keep list, -1 = suppressed). */
int stl_oks_nms(const float* keypoints, const double* area, const double* box_score, const int* image_offsets,
                int n_images, int max_persons_per_image, int J, const double* vars, float in_vis_thr, double oks_thr,
                float nms_vis_thr, int rescore, double* score_out, int* keep_rank, void* stream);

/* Training targets: JointsDataset.generate_target (data/JointsDataset.py:230-286) for a batch.
 * joints, joints_vis: [B][J][3] fp64 on the device (x, y in crop pixels; visibility in column 0); joints_weight: [J] fp32
 * or null (use_different_joints_weight); outputs target fp32 [B][J][h][w] (every element written) and target_weight
 * fp32 [B][J].  Heatmap stride = image size / heatmap size; sigma as in the reference's config (2). */
int stl_generate_target(const double* joints, const double* joints_vis, const float* joints_weight, int B, int J, int h,
                        int w, int image_h, int image_w, int sigma, float* target, float* target_weight, void* stream);

/* Second decode path: create_pose_from_outputs (lib/pose_parsing.py:138-151) =
 * F.interpolate(heat, (out_h, out_w), mode="bilinear", align_corners=True) -> get_max_preds_hrnet, used by
 * 04_evaluate_vases_qualitatively.py:216-220 and 05_create_archdata_retrieval_db.py:114.  Fused: the upsampled tensor
 * is never written.  coords [B][J][2] = (x, y) in the upsampled grid (zeroed where the maximum is <= 0), maxvals [B][J]. */
int stl_upsampled_argmax(const float* heat, int B, int J, int h, int w, int out_h, int out_w, float* coords,
                         float* maxvals, void* stream);

/* Crop extraction, the step in front of the network: TransformDetection.__call__ / crop
 * (lib/transforms.py:30-58, 259-268) and JointsDataset.__getitem__ (data/JointsDataset.py:189-197) call
 * cv2.warpAffine(img, M, (out_w, out_h), flags=INTER_LINEAR) per box.  img: uint8 [img_h][img_w][3] on the device;
 * minv: [N][6] float64 on the device, the INVERTED 2x3 matrices (crop -> image) exactly as cv2.warpAffine derives them;
 * outputs (either may be null): uint8 [N][3][out_h][out_w], and the network input fp32 [N][3][out_h][out_w] =
 * (v/255 - mean[c]) / std[c] (ToTensor + Normalize; mean3_host / std3_host: 3 host floats each, null = 0 / 1).
 * OpenCV's fixed-point bilinear arithmetic is reproduced bit for bit. */
int stl_warp_affine_crops(const void* img_u8_hwc, int img_h, int img_w, const double* minv, int N, int out_h, int out_w,
                          void* out_u8_nchw, float* out_f32_nchw, const float* mean3_host, const float* std3_host,
                          void* stream);

/* The same for float32 HWC images (04_evaluate_vases_qualitatively.py:209-213 hands TransformDetection a float array;
 * cv2.warpAffine then interpolates in float32 with its bilinear weight table): fp32 [N][3][out_h][out_w], bit-identical
 * to cv2's result. */
int stl_warp_affine_crops_f32(const float* img_f32_hwc, int img_h, int img_w, const double* minv, int N, int out_h,
                              int out_w, float* out_f32_nchw, void* stream);

/* optimizer.step() of the fine-tuning loop (02_train.py:218; torch.optim.SGD of lib/model_setup.py:138-139 with momentum,
 * weight decay and optional Nesterov momentum, dampening 0) for ALL parameter tensors in one launch.  items: n_items
 * records {float* p; const float* grad; float* momentum_buffer (null = no momentum); int32 numel; int32 pad} on the
 * device; block_offsets[i] = first 1 024-element block of item i, total_blocks = their sum.  Same operations in the same
 * order as torch: g = grad + wd * p; buf = momentum * buf + g (a zeroed buffer gives torch's first step); p -= lr * g. */
int stl_sgd_step_batched(const void* items, const int* block_offsets, int n_items, int total_blocks, float lr,
                         float momentum, float weight_decay, int nesterov, void* stream);

/* conv1 of the stem (models/HRnet.py:286: 3 -> 64 channels, 3x3, stride 2, pad 1) as a 1x1 problem: fp32 NCHW
 * [N][3][H][W] -> its im2col rows, padded-linear bf16 [N][H/2+1][W/2+1][32] with K index ci*9 + kh*3 + kw (the OIHW
 * flatten order, 27 values + 5 zeros).  conv1 is then stl_conv2d with ksize 1 on these rows and the weight read as
 * [64][27][1][1], and its weight gradient a 1x1 stl_conv_wgrad (tensor cores).  The zero cells of `rows` are not
 * written: clear the buffer once. */
int stl_stem_im2col(const float* x_nchw, void* rows, int N, int H, int W, void* stream);

/* PCK accuracy of the training / evaluation loops (lib/metrics.py:268-364 accuracy -> calc_dists -> dist_acc, called at
 * 02_train.py:223,277 and 03_evaluate.py:142 on output.cpu()): pred_coords / target_coords are the [B][J][2] arg-max
 * coordinates of the predicted and the ground-truth heatmaps (stl_decode with refine = 0).  A joint is counted when its
 * target has x > 1 and y > 1 and is a hit when ||(pred - target) / (h/10, w/10)|| < thr.  acc: [J+1] (acc[0] = mean of
 * the per-joint accuracies that are >= 0, acc[1+j] = hits/counted or -1), avg_acc: 1 float, cnt: 1 int. */
int stl_pck_accuracy(const float* pred_coords, const float* target_coords, int B, int J, int h, int w, float thr,
                     float* acc, float* avg_acc, int* cnt, void* stream);

/* x[0:n] *= *scale_dev unless *scale_dev == 1 (then no memory is touched): applies autograd's upstream gradient of the
 * loss to the gradient stl_mse_loss_fwd_bwd already produced, without a host read of the scalar. */
int stl_scale_inplace(float* x, const float* scale_dev, long long n, void* stream);
int stl_mse_loss_fwd_bwd(const float* out, const float* tgt, const float* tw, int B, int J, int hw, float* loss,
                         float* grad, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Single fused convolution (unit-test entry points; replaces one nn.Conv2d + nn.BatchNorm2d [+ add] [+ ReLU]
 * group of models/HRnet.py, e.g. :48-59, :85-100, :198-240).
 *
 * Activations use the engine's padded-linear NHWC bf16 layout: N*(H+1)*(W+1) pixels of C channels, local row H
 * and local column W of every image zero (see stlpose_b200/csrc/conv.h).
 * ---------------------------------------------------------------------------------------------- */
size_t stl_padded_bytes(int N, int C, int H, int W);
int stl_nchw_to_padded(const float* x_nchw, void* y_padded, int N, int C, int H, int W, int C_pad, void* stream);
int stl_padded_to_nchw(const void* y_padded, float* x_nchw, int N, int C, int H, int W, int C_pad, void* stream);

/* OIHW fp32 weights (+ optional eval-mode BatchNorm gamma/beta/mean/var, + optional conv bias) ->
 * [k*k][Cout_pad][Cin_pad] bf16 with the BN scale folded in, and a [Cout_pad] fp32 bias. */
int stl_pack_conv_weights(const float* w_oihw, const float* bn_gamma, const float* bn_beta, const float* bn_mean,
                          const float* bn_var, const float* conv_bias, float eps, int Cout, int Cin, int ksize,
                          int Cout_pad, int Cin_pad, void* w_packed, float* bias_packed, void* stream);
/* One BasicBlock (models/HRnet.py:45-61, eval mode, BatchNorm folded by stl_pack_conv_weights) in one kernel:
 *   y = relu(conv2(relu(conv1(x) + bias1)) + bias2 + x), both convolutions 3x3 / stride 1 / C -> C, C = 32.
 * The intermediate activation stays in shared memory (conv1's accumulators are written as conv2's tensor-core operand).
 * x, y: padded-linear bf16 [N][H+1][W+1][C] (distinct buffers); w*_packed: [9][C][C] bf16; bias*: C fp32.
 * Results are bit-identical to two stl_conv2d calls. */
int stl_basic_block(const void* x, void* y, const void* w1_packed, const float* bias1, const void* w2_packed,
                    const float* bias2, int N, int H, int W, int C, void* stream);

/* The junction of two Bottlenecks (models/HRnet.py:88-101 of one block, :82-84 of the next; eval mode, BatchNorm
 * folded) in one kernel:
 *   out = relu(conv3(t) + bias3 + x)        1x1, 64 -> 256, x = the block's input (the residual)
 *   a   = relu(conv1n(out) + bias1n)        1x1, 256 -> 64, conv1 of the NEXT Bottleneck
 * The finished bf16 tile of `out`, staged in shared memory for its store, is also the tensor-core operand of the second
 * product: the 256-channel tensor crosses HBM once instead of being written and read back.
 * t, a: padded-linear bf16 [N][H+1][W+1][64]; x, out: [..][256] (distinct buffers); w3_packed [256][64],
 * w1n_packed [64][256] bf16; bias3: 256, bias1n: 64 fp32.  `out` is bit-identical to stl_conv2d with the residual, `a`
 * to stl_conv2d on that `out`.  max_ctas: 0 = one CTA per SM. */
int stl_bottleneck_link(const void* t, const void* x, void* out, void* a, const void* w3_packed, const float* bias3,
                        const void* w1n_packed, const float* bias1n, int N, int H, int W, int max_ctas, void* stream);
/* The junction after the FIRST Bottleneck (layer1.0 -> layer1.1), whose shortcut is a convolution: no residual, GEMM1 runs
 * over two inputs instead,  out = relu([W3 | Wd] . [t | t2] + bias3)  with t2 = the block's 64-channel input,
 * w3cat_packed [256][128] and bias3 = the summed folded biases (see stl_conv_desc.in2), then a as above.
 * Bit-identical to the two-input stl_conv2d followed by the 256 -> 64 stl_conv2d. */
int stl_bottleneck_link2(const void* t, const void* t2, void* out, void* a, const void* w3cat_packed, const float* bias3,
                         const void* w1n_packed, const float* bias1n, int N, int H, int W, int max_ctas, void* stream);

/* Weights for the convolution that IS the stride-1 input gradient: dx = stl_conv2d(dz, W'), W'[ci][co][kh][kw] =
 * W[co][ci][k-1-kh][k-1-kw].  w: fp32 OIHW of the forward layer; result [k*k][Rows_pad][K_pad] bf16 with Rows_pad >= Cin
 * (multiple of 16) and K_pad >= Cout (the channel count of dz); bias_packed (Rows_pad floats, may be null) is zeroed. */
int stl_pack_conv_weights_dgrad(const float* w, int Cout, int Cin, int ksize, int Rows_pad, int K_pad, void* w_packed,
                                float* bias_packed, void* stream);

/* All raw (no BatchNorm folding) weight repacks of a training step in ONE launch: the forward layout of
 * stl_pack_conv_weights (dgrad = 0: [k*k][rows_pad >= Cout][cols_pad >= Cin]) and the input-gradient layout of
 * stl_pack_conv_weights_dgrad (dgrad = 1: [k*k][rows_pad >= Cin][cols_pad >= Cout], taps reversed).  items_dev: n_items
 * descriptors in DEVICE memory; block_offsets_dev: n_items + 1 ints (device), item i owns thread blocks
 * [block_offsets[i], block_offsets[i+1]) of 256 threads x 4 elements; total_blocks = block_offsets[n_items]. */
typedef struct stl_pack_item {
  const float* w;        /* fp32 OIHW master weights */
  void* w_packed;        /* bf16 destination */
  int Cout, Cin, ksize;
  int rows_pad, cols_pad;
  int dgrad;
} stl_pack_item;
int stl_pack_conv_weights_batched(const stl_pack_item* items_dev, const int* block_offsets_dev, int n_items,
                                  int total_blocks, void* stream);

typedef struct stl_conv_desc {
  const void* in;          /* padded-linear bf16 [N][H+1][W+1][Cin] */
  int N, H, W, Cin;        /* input geometry */
  void* out;               /* padded-linear bf16 (out_nchw = 0) or fp32 [N][Cout][Ho][Wo] (out_nchw = 1) */
  int Cout, Cout_pad;
  int ksize, stride;       /* 1|3, 1|2 (padding = ksize/2) */
  const void* w_packed;
  const float* bias_packed;
  const void* residual;    /* optional, same geometry as out (may alias out) */
  int n_up;                /* optional nearest-upsampled addends (fuse layers) */
  const void* up_src[STL_MAX_UP];
  int up_shift[STL_MAX_UP];
  int relu;
  int out_nchw;
  int impl;                /* 0 = tcgen05 (shifted-descriptor taps), 1 = tcgen05 (one TMA load per tap),
                              2 = CUDA-core reference kernel (validation only),
                              3 = tcgen05 with the three taps of a filter row merged into one MMA (N = 3*Cout) where
                                  the layer allows it (stride-1 3x3, resident weights), else as 0 */
  int force_mb;            /* 0 = auto */
  int max_ctas;            /* 0 = auto */
  void* dbg_counters;      /* optional int64 [148][3][4]: per-CTA cycle counters of the producer / MMA / epilogue
                              roles (measurement aid; null in normal use) */
  const void* in2;         /* optional second input of a 1x1 / stride-1 convolution, padded-linear bf16 [N][H+1][W+1][Cin2]:
                              K = [in | in2], w_packed has Cin + Cin2 columns.  Two 1x1 convolutions that are summed
                              before the activation (Bottleneck conv3 + downsample, models/HRnet.py:88-101) run as one
                              launch; their sum stays in the fp32 accumulator.  impl 0 only. */
  int Cin2;
  int pdl;                 /* 1: launch with programmatic stream serialization - the kernel's prologue (barriers, tensor
                              memory, resident weights) overlaps the tail of the previous kernel of the stream and
                              everything that touches activations waits for that kernel to complete.  w_packed and
                              bias_packed must not be written by the immediately preceding kernel. */
} stl_conv_desc;

int stl_conv2d(const stl_conv_desc* desc, void* stream);
/* The same convolution, additionally accumulating the BatchNorm batch statistics of its output in the epilogue (training:
 * HRnet.py:48-59 under model.train(), 02_train.py:208): per-channel sum and sum of squares of the STORED bf16 values over
 * the valid pixels.  `stats`: stl_conv2d_stats_floats(Cout_pad) floats; the launch fills *stats_rows rows of
 * [2][Cout_pad] (one per CTA) which stl_bn_train_forward_fused adds in a fixed order (deterministic).  *stats_rows == 0:
 * this shape - or this build: the fused variants were measured slower than the separate statistics pass (DESIGN.md
 * section 4) and are only compiled with -DSTL_CONV_STATS - has no fused statistics (the convolution still ran): use
 * stl_bn_train_forward. */
/* stl_conv2d_bn goes one step further: the LAST CTA of the convolution adds the rows in a fixed order and writes
 * mean[C], rstd[C] and the running statistics (momentum, unbiased variance) itself, so nothing is launched between the
 * convolution and the normalisation (stl_bn_apply).  `ticket`: a device word that is zero on entry and left zero.
 * *done == 0: not for this shape - nothing was finalised, use stl_bn_train_forward. */
size_t stl_conv2d_stats_floats(int cout_pad);
int stl_conv2d_bn(const stl_conv_desc* desc, float* stats, unsigned* ticket, float eps, float momentum, float* mean,
                  float* rstd, float* running_mean, float* running_var, int* done, void* stream);
/* y = [relu](gamma * (z - mean) * rstd + beta [+ residual]) with given statistics (zero cells stay zero). */
int stl_bn_apply(const void* z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                 const void* residual, int relu, int N, int H, int W, int C, void* y, void* stream);
int stl_conv2d_stats(const stl_conv_desc* desc, float* stats, int* stats_rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole-network plan: replaces PoseHighResolutionNet.forward (models/HRnet.py:433-468) in eval mode and the
 * two-pass flip test of lib/inference.py:18-22.
 * ---------------------------------------------------------------------------------------------- */
typedef struct stl_plan stl_plan;

typedef struct stl_hrnet_cfg {
  int width;            /* 32 or 48: branch widths are width * {1,2,4,8}   (MODEL.EXTRA.STAGE*.NUM_CHANNELS) */
  int num_joints;       /* MODEL.NUM_JOINTS (17) */
  int stage_modules[3]; /* NUM_MODULES of STAGE2..4 ({1,4,3}) */
  int blocks;           /* NUM_BLOCKS per branch (4) */
  int image_h, image_w; /* crop size, multiples of 32 */
} stl_hrnet_cfg;

typedef struct stl_conv_info {
  char conv_key[96];  /* state_dict prefix of the conv   ("stage3.1.branches.0.2.conv1") */
  char bn_key[96];    /* state_dict prefix of its BN, "" when the conv has a bias instead (final_layer) */
  int cout, cin, ksize, stride;
} stl_conv_info;

stl_plan* stl_plan_create(const stl_hrnet_cfg* cfg);
void stl_plan_destroy(stl_plan* plan);
int stl_plan_num_convs(const stl_plan* plan);
int stl_plan_conv_info(const stl_plan* plan, int index, stl_conv_info* info);

/* Packed-parameter arena (bf16 folded weights + fp32 biases for every conv). */
size_t stl_plan_weight_bytes(const stl_plan* plan);
/* Fold and pack conv `index` from fp32 device tensors (models/HRnet.py state_dict entries). */
int stl_plan_pack_conv(stl_plan* plan, int index, const float* w_oihw, const float* bn_gamma, const float* bn_beta,
                       const float* bn_mean, const float* bn_var, const float* conv_bias, float eps,
                       void* weight_arena, void* stream);

/* Activation workspace for n_images images (n_images = 2*B when the flip-test pass is batched in). */
size_t stl_plan_workspace_bytes(const stl_plan* plan, int n_images);

/* x_nchw: fp32 [B][3][H][W].  flip_pair != 0 runs the mirrored crops as images B..2B-1 of the same batch.
 * heat_nchw: fp32 [(flip_pair ? 2 : 1) * B][J][H/4][W/4]  (second half = raw flipped-pass heatmaps; feed both
 * halves to stl_decode / stl_flip_avg).  The workspace is zero-initialised by the plan whenever the
 * (workspace, n_images) binding changes. */
int stl_plan_forward(stl_plan* plan, const float* x_nchw, int B, int flip_pair, float* heat_nchw,
                     const void* weight_arena, void* workspace, size_t workspace_bytes, void* stream);

/* Number of ops of the plan (size of the per-op timing array) and number of kernels one stl_plan_forward enqueues
 * for the current binding (an op of a sub-batched group is launched once per sub-batch). */
int stl_plan_launches_per_forward(const stl_plan* plan);
int stl_plan_kernel_launches(const stl_plan* plan);

/* Measurement aid: same as stl_plan_forward, but brackets every launch with CUDA events on `stream`, waits for
 * the stream and writes the per-launch durations (ms) to op_ms_host[stl_plan_launches_per_forward()]. */
int stl_plan_forward_timed(stl_plan* plan, const float* x_nchw, int B, int flip_pair, float* heat_nchw,
                           const void* weight_arena, void* workspace, size_t workspace_bytes, void* stream,
                           float* op_ms_host);

typedef struct stl_op_info {
  int kind;                 /* 0 = stem input packing, 1 = tcgen05 conv, 2 = fuse-sum, 3 = fused BasicBlock (two 3x3
                               convs), 4 = Bottleneck junction (conv3 + the next block's conv1, two 1x1 convs) */
  int layer;                /* conv index for stl_plan_conv_info, -1 for fuse-sum */
  int cin, cout, ksize, stride, out_h, out_w;
  double flops_per_image;   /* 2*MACs */
  double bytes_per_image;   /* algorithmic: input + output (+ residual / upsampled addends) read or written once */
  int grid, smem, mb, nt, ck, a_stages, b_stages, a_shift, tiles;   /* launch shape (valid once the plan has run) */
  int subs;                 /* sub-batches this op is launched in (L2-resident branch groups), 1 otherwise */
} stl_op_info;
int stl_plan_op_info(const stl_plan* plan, int op_index, stl_op_info* info);

/* ------------------------------------------------------------------------------------------------
 * Training path (fine-tuning step, 02_train.py:203-218: model.train() forward, loss.backward()).
 * Activations / activation gradients: padded-linear NHWC bf16.  Statistics and parameter gradients: fp32.
 * ---------------------------------------------------------------------------------------------- */

/* nn.BatchNorm2d in training mode fused with the block's residual add and ReLU (HRnet.py:48-59, 85-100):
 *   batch mean / biased variance of z over N*H*W -> y = [relu](gamma*(z-mean)*rstd + beta [+ residual]);
 *   running_mean / running_var (may be null) updated with `momentum` and the unbiased variance.
 * sums: fp32 workspace of stl_bn_workspace_floats(C) elements (per-block partial sums; the reduction uses no
 * floating-point atomics and is deterministic); sums[0:2C] holds the channel sums on return.
 * mean, rstd: C fp32 outputs kept for the backward. */
size_t stl_bn_workspace_floats(int C);
int stl_bn_train_forward(const void* z, const float* gamma, const float* beta, const void* residual, int relu,
                         float eps, float momentum, int N, int H, int W, int C, void* y, float* sums, float* mean,
                         float* rstd, float* running_mean, float* running_var, void* stream);

/* Backward of the same unit: g = dy masked by (y > 0) when relu; sums (same workspace size as above): sums[0:C] =
 * dbeta = sum g, sums[C:2C] = dgamma = sum g*xhat; dz = gamma*rstd*(g - dbeta/cnt - xhat*dgamma/cnt); dres (may be null) = g. */
int stl_bn_train_backward(const void* dy, const void* y, const void* z, const float* mean, const float* rstd,
                          const float* gamma, int relu, int N, int H, int W, int C, void* dz, void* dres, float* sums,
                          void* stream);

/* The same two calls for a step that is captured once and replayed (TrainStep): no memset node per call and no copy of the
 * parameter gradients.  `ticket` is a caller-owned device word that is ZERO on entry and is left zero; launches that may
 * run concurrently must not share one (one per BatchNorm layer and direction).  backward: the channel sums go to
 * dbeta_dgamma ([2C], a tensor of their own: dbeta | dgamma); workspace holds stl_bn_workspace_floats(C) floats. */
int stl_bn_train_forward_ticket(const void* z, const float* gamma, const float* beta, const void* residual, int relu,
                                float eps, float momentum, int N, int H, int W, int C, void* y, float* sums, float* mean,
                                float* rstd, float* running_mean, float* running_var, unsigned* ticket, void* stream);
/* Train-mode BatchNorm whose statistics come from stl_conv2d_stats: mean / rstd / running statistics from the `rows` rows
 * of [2][c_pad] partial sums, then y = [relu](gamma * (z - mean) * rstd + beta [+ residual]) as above. */
int stl_bn_train_forward_fused(const void* z, const float* stat_rows, int rows, int c_pad, const float* gamma,
                               const float* beta, const void* residual, int relu, float eps, float momentum, int N, int H,
                               int W, int C, void* y, float* mean, float* rstd, float* running_mean, float* running_var,
                               void* stream);
int stl_bn_train_backward_ticket(const void* dy, const void* y, const void* z, const float* mean, const float* rstd,
                                 const float* gamma, int relu, int N, int H, int W, int C, void* dz, void* dres,
                                 float* dbeta_dgamma, float* workspace, unsigned* ticket, void* stream);
/* The same for a ReLU unit WITHOUT residual: the ReLU mask (y > 0) is recomputed from z - gamma * (z - mean) * rstd + beta
 * evaluated with exactly the forward's operations - instead of being read from the stored output, which saves one of
 * the three tensor reads of each backward pass. */
int stl_bn_train_backward_ticket_z(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma,
                                   const float* beta, int N, int H, int W, int C, void* dz, float* dbeta_dgamma,
                                   float* workspace, unsigned* ticket, void* stream);
/* The forward pair (statistics, normalisation) and the backward pair (channel sums, gradient) of a BatchNorm layer as ONE
 * cooperative launch each: the block that finishes the reduction publishes its result, the others wait for it inside the
 * kernel, then all normalise.  At fine-tuning batch sizes the separate launches are latency, not bandwidth.  `sync`: two
 * device words that are zero on entry and left zero (one pair per layer and direction).  backward relu: 0 none, 1 mask
 * from y, 2 mask recomputed from z (needs beta, no residual).  Same results as the separate entry points, bit for bit. */
int stl_bn_train_forward_coop(const void* z, const float* gamma, const float* beta, const void* residual, int relu,
                              float eps, float momentum, int N, int H, int W, int C, void* y, float* sums, float* mean,
                              float* rstd, float* running_mean, float* running_var, unsigned* ticket, unsigned* sync,
                              void* stream);
int stl_bn_train_backward_coop(const void* dy, const void* y, const void* z, const float* mean, const float* rstd,
                               const float* gamma, const float* beta, int relu, int N, int H, int W, int C, void* dz,
                               void* dres, float* dbeta_dgamma, float* workspace, unsigned* ticket, unsigned* sync,
                               void* stream);

/* Fuse-layer row (HRnet.py:255-264): y = relu(sum same[i] + sum nearest_upsample(up[j], 2^shift[j])).
 * same_host / up_host: host arrays of device pointers (n_same <= 4, n_up <= 3). */
int stl_sum_relu_forward(const void* const* same_host, int n_same, const void* const* up_host, const int* shift_host,
                         int n_up, void* y, int N, int H, int W, int C, void* stream);
/* g = dy * (y > 0): gradient of every same-resolution addend of the row above. */
int stl_relu_mask(const void* dy, const void* y, void* g, long long elems, void* stream);
/* gradient of a nearest-upsampled addend: dlow = sum of g over each 2^shift x 2^shift window. */
int stl_upsample_backward(const void* g, void* dlow, int N, int H, int W, int C, int shift, void* stream);

/* u[n,h,w] = dz[n,h/2,w/2] at even (h,w), 0 elsewhere: dz is padded-linear [N][H/2+1][W/2+1][C], u is [N][H+1][W+1][C].
 * Gradients of a stride-2 convolution are the stride-1 gradients of the stuffed dz. */
int stl_zero_stuff(const void* dz, void* u, int N, int H, int W, int C, void* stream);

/* Convolution gradients (autograd of nn.Conv2d).  w_packed: [k*k][Cout][Cin] bf16 as produced by
 * stl_pack_conv_weights without BatchNorm.  dx: padded-linear bf16 [N][Hi+1][Wi+1][Cin];
 * dw: fp32 [Cout][cin_real][k][k] (OIHW), zeroed by the call. */
int stl_conv_dgrad(const void* dz, const void* w_packed, void* dx, int N, int Hi, int Wi, int Cin, int Cout, int ksize,
                   int stride, void* stream);
/* stl_conv_wgrad runs stride-1 problems with channel counts of 32/64/128/256 on the tcgen05 tensor cores (stride-2
 * layers get there through stl_zero_stuff): every CTA accumulates a pixel range in TMEM and writes a slab of partial
 * sums to `workspace` (stl_conv_wgrad_workspace_bytes, 0 when the CUDA-core kernel will be used); a second kernel adds
 * the slabs in a fixed order, so the result is deterministic.  Other shapes (native stride 2, the 3-channel stem,
 * HRNet-W48's channel counts) use a CUDA-core kernel built the same way: per-block slabs of partial sums in `workspace`
 * and a fixed-order second pass - no floating-point atomics anywhere, every gradient is bit-reproducible.
 * stl_conv_wgrad_naive forces the CUDA-core kernel (validation; its workspace size comes from
 * stl_conv_wgrad_naive_workspace_bytes). */
size_t stl_conv_wgrad_workspace_bytes(int N, int Hi, int Wi, int Cin, int Cout, int ksize, int stride, int cin_real);
int stl_conv_wgrad(const void* x, const void* dz, float* dw, int N, int Hi, int Wi, int Cin, int Cout, int ksize,
                   int stride, int cin_real, void* workspace, size_t workspace_bytes, void* stream);
size_t stl_conv_wgrad_naive_workspace_bytes(int N, int Hi, int Wi, int Cin, int Cout, int ksize, int stride,
                                            int cin_real);
int stl_conv_wgrad_naive(const void* x, const void* dz, float* dw, int N, int Hi, int Wi, int Cin, int Cout, int ksize,
                         int stride, int cin_real, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STLPOSE_B200_H_ */

// Bandwidth-bound kernels of the keypoint path: flip-test averaging, heatmap decode, joints-MSE loss.
//
// References (paths relative to /root/reference/src):
//   flip_back + shift + average : lib/transforms.py:147-164, lib/inference.py:21-26
//   argmax / refine / affine    : lib/pose_parsing.py:16-92, lib/transforms.py:184-240
//   PersonMSELoss               : lib/loss.py:61-94
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv.h"
#include "pose_kernels.h"

namespace stl {

void set_error(const char* fmt, ...);

namespace {

struct Perm {
  int src[kMaxJoints];  // channel of the flipped-pass heatmap that feeds joint j after flip_back
};

// ------------------------------------------------------------------ flip-test average (standalone)
// out[n,j,y,x] = 0.5 * (a[n,j,y,x] + f[n,perm[j],y,w - max(x,1)])
//   flip_back reverses W and swaps left/right joints; inference.py:25 then shifts the flipped map right by one
//   pixel keeping column 0, so column x reads reversed column x-1, i.e. source column w-x (w-1 for x = 0).
__global__ void flip_avg_kernel(const float* __restrict__ a, const float* __restrict__ f, float* __restrict__ out,
                                int B, int J, int h, int w, const Perm perm) {
  const long long total = (long long)B * J * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const long long t = i / w;
    const int y = (int)(t % h);
    const long long nj = t / h;
    const int j = (int)(nj % J);
    const long long n = nj / J;
    const int xs = w - (x > 0 ? x : 1);
    const float fv = f[((n * J + perm.src[j]) * h + y) * (long long)w + xs];
    out[i] = (a[i] + fv) * 0.5f;
  }
}

// flip_back alone (transforms.py:147-164): out[n,j,y,x] = in[n,perm[j],y,w-1-x]  (pure permutation, exact)
__global__ void flip_back_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int J, int h, int w,
                                 const Perm perm) {
  const long long total = (long long)B * J * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const long long t = i / w;
    const int y = (int)(t % h);
    const long long nj = t / h;
    const int j = (int)(nj % J);
    const long long n = nj / J;
    out[i] = in[((n * J + perm.src[j]) * h + y) * (long long)w + (w - 1 - x)];
  }
}

// ------------------------------------------------------------------ decode (optionally fused with flip-avg)
__device__ __forceinline__ float hm_value(const float* __restrict__ a, const float* __restrict__ frow_base, int w,
                                          int y, int x) {
  // averaged heatmap value at (y, x); frow_base = flipped-pass map of the swapped joint (or nullptr)
  float v = a[y * w + x];
  if (frow_base) v = (v + frow_base[y * w + (w - (x > 0 ? x : 1))]) * 0.5f;
  return v;
}

// One warp per (crop, joint) map.  Lanes stride over float4 groups (coalesced 512 B per warp request);
// strict '>' keeps the first index within a lane, the shuffle reduction breaks ties towards the lower index,
// which reproduces np.argmax's first-occurrence rule bit for bit.
__global__ void __launch_bounds__(256) decode_kernel(const float* __restrict__ heat, const float* __restrict__ heat_f,
                                                     const float* __restrict__ center, const float* __restrict__ scale,
                                                     int B, int J, int h, int w, const Perm perm, int refine,
                                                     float* __restrict__ avg_out, float* __restrict__ preds,
                                                     float* __restrict__ maxvals, float* __restrict__ coords) {
  const int lane = threadIdx.x & 31;
  const long long map = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (map >= (long long)B * J) return;
  const int j = (int)(map % J);
  const long long n = map / J;
  const int hw = h * w;
  const float* a = heat + map * hw;
  const float* f = heat_f ? heat_f + (n * J + perm.src[j]) * hw : nullptr;
  float* o = avg_out ? avg_out + map * hw : nullptr;

  float best = -INFINITY;
  int best_i = 0x7fffffff;
  if ((w & 3) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(a);
    for (int v = lane; v < hw / 4; v += 32) {
      float4 t = __ldcs(a4 + v);
      const int i0 = v * 4;
      if (f) {
        const int y = i0 / w, x0 = i0 - y * w;
        const float* fr = f + y * w;
        t.x = (t.x + __ldg(fr + (w - (x0 > 0 ? x0 : 1)))) * 0.5f;
        t.y = (t.y + __ldg(fr + (w - (x0 + 1)))) * 0.5f;
        t.z = (t.z + __ldg(fr + (w - (x0 + 2)))) * 0.5f;
        t.w = (t.w + __ldg(fr + (w - (x0 + 3)))) * 0.5f;
        if (o) __stcs(reinterpret_cast<float4*>(o) + v, t);
      }
      if (t.x > best) { best = t.x; best_i = i0; }
      if (t.y > best) { best = t.y; best_i = i0 + 1; }
      if (t.z > best) { best = t.z; best_i = i0 + 2; }
      if (t.w > best) { best = t.w; best_i = i0 + 3; }
    }
  } else {
    for (int i = lane; i < hw; i += 32) {
      const int y = i / w, x = i - y * w;
      const float t = hm_value(a, f, w, y, x);
      if (f && o) o[i] = t;
      if (t > best) { best = t; best_i = i; }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
    if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  if (lane != 0) return;
  if (best_i == 0x7fffffff) best_i = 0;  // all -inf / NaN maps: out of contract, stay in bounds

  // get_max_preds_hrnet (pose_parsing.py:44-53)
  float cx = (float)(best_i % w), cy = (float)(best_i / w);
  if (!(best > 0.f)) { cx = 0.f; cy = 0.f; }
  // quarter-pixel refinement (pose_parsing.py:70-82)
  const int px = (int)floorf(cx + 0.5f), py = (int)floorf(cy + 0.5f);
  if (refine && 1 < px && px < w - 1 && 1 < py && py < h - 1) {
    const float dx = hm_value(a, f, w, py, px + 1) - hm_value(a, f, w, py, px - 1);
    const float dy = hm_value(a, f, w, py + 1, px) - hm_value(a, f, w, py - 1, px);
    cx += dx > 0.f ? 0.25f : (dx < 0.f ? -0.25f : 0.f);
    cy += dy > 0.f ? 0.25f : (dy < 0.f ? -0.25f : 0.f);
  }
  maxvals[map] = best;
  coords[map * 2 + 0] = cx;
  coords[map * 2 + 1] = cy;
  if (preds) {
    // transform_preds with rot = 0 (transforms.py:184-233): isotropic map, only scale[0] is used (:209).
    // The reference solves the 3-point affine in float64; this is its closed form, evaluated in float64.
    const double k = (double)scale[n * 2 + 0] * 200.0 / (double)w;
    preds[map * 2 + 0] = (float)((double)center[n * 2 + 0] + ((double)cx - 0.5 * w) * k);
    preds[map * 2 + 1] = (float)((double)center[n * 2 + 1] + ((double)cy - 0.5 * h) * k);
  }
}

// ------------------------------------------------------------------ PersonMSELoss forward (+ gradient)
// loss = 0.5/(J*B*hw) * sum (tw*(o-t))^2 ; dloss/do = tw^2*(o-t)/(J*B*hw).  Deterministic two-pass reduction.
constexpr int kLossBlock = 256;
__global__ void __launch_bounds__(kLossBlock) mse_partial_kernel(const float* __restrict__ o,
                                                                 const float* __restrict__ t,
                                                                 const float* __restrict__ tw, long long n4, int hw4,
                                                                 float inv_denom, float* __restrict__ grad,
                                                                 double* __restrict__ partial) {
  double acc = 0.0;
  const float4* o4 = reinterpret_cast<const float4*>(o);
  const float4* t4 = reinterpret_cast<const float4*>(t);
  float4* g4 = reinterpret_cast<float4*>(grad);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float wgt = __ldg(tw + i / hw4);
    const float4 a = __ldcs(o4 + i), b = __ldcs(t4 + i);
    const float dx = (a.x - b.x) * wgt, dy = (a.y - b.y) * wgt, dz = (a.z - b.z) * wgt, dw = (a.w - b.w) * wgt;
    acc += (double)(dx * dx + dy * dy) + (double)(dz * dz + dw * dw);
    if (g4) {
      const float s = wgt * inv_denom;
      __stcs(g4 + i, make_float4(dx * s, dy * s, dz * s, dw * s));
    }
  }
  __shared__ double red[kLossBlock / 32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < kLossBlock / 32; ++i) s += red[i];
    partial[blockIdx.x] = s;
  }
}

__global__ void mse_final_kernel(const double* __restrict__ partial, int n, double scale, float* __restrict__ loss) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    *loss = (float)(s * scale);
  }
}

// ------------------------------------------------------------------ OKS rescoring + greedy OKS-NMS per image
// generate_submission_hrnet (lib/metrics.py:232-258) + nms.oks_nms / oks_iou (lib/nms.py:10-74): per person,
// score = mean(joint scores > in_vis_thr) * box score; per image, persons are visited by descending score and a person
// is dropped when its OKS with an already kept one exceeds oks_thr.  OKS = mean_j exp(-(d_j^2 / var_j) / ((a_g+a_d)/2 + eps) / 2)
// over the joints selected by vis_thr (the reference's `list(vg > t) and list(vd > t)` keeps the mask of d only; < 0 = all).
// One block per image (persons [off[i], off[i+1]), any number: lib/nms.py has no limit); float64 like NumPy; joint scores are summed in
// float32 in joint order like the reference's scalar loop.  keep_rank[m] = position in the keep list or -1.
__global__ void __launch_bounds__(128) oks_nms_kernel(const float* __restrict__ kpts, const double* __restrict__ area,
                                                      const double* __restrict__ box_score, const int* __restrict__ off,
                                                      int J, const double* __restrict__ vars, float in_vis_thr,
                                                      double oks_thr, float nms_vis_thr, int rescore, int max_persons,
                                                      double* __restrict__ score_out, int* __restrict__ keep_rank) {
  // dynamic shared memory sized by the caller's max_persons_per_image: [scores f64][order i32][dead u8]
  extern __shared__ __align__(8) unsigned char nms_smem[];
  double* s_score = reinterpret_cast<double*>(nms_smem);
  int* s_order = reinterpret_cast<int*>(s_score + max_persons);
  unsigned char* s_dead = reinterpret_cast<unsigned char*>(s_order + max_persons);
  const int lo = off[blockIdx.x], n_all = off[blockIdx.x + 1] - lo;
  if (n_all <= 0) return;
  // (should the offsets exceed the declared maximum, the surplus persons are reported as suppressed rather than
  // indexing past the shared arrays)
  const int n = n_all < max_persons ? n_all : max_persons;
  for (int i = n + threadIdx.x; i < n_all; i += blockDim.x) { keep_rank[lo + i] = -1; score_out[lo + i] = box_score[lo + i]; }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float* kp = kpts + (size_t)(lo + i) * J * 3;
    double sc = box_score[lo + i];
    if (rescore) {
      float acc = 0.f;
      int valid = 0;
      for (int j = 0; j < J; ++j) {
        const float js = kp[j * 3 + 2];
        if (js > in_vis_thr) { acc = acc + js; ++valid; }
      }
      if (valid) acc = acc / (float)valid;
      sc = (double)acc * sc;
    }
    s_score[i] = sc;
    score_out[lo + i] = sc;
    s_dead[i] = 0;
    keep_rank[lo + i] = -1;
  }
  __syncthreads();
  // descending order by rank counting (ties: the later index first, like argsort()[::-1] on distinct scores)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int r = 0;
    for (int k = 0; k < n; ++k) r += (s_score[k] > s_score[i]) || (s_score[k] == s_score[i] && k > i);
    s_order[r] = i;
  }
  __syncthreads();
  int kept = 0;
  for (int a = 0; a < n; ++a) {
    const int g = s_order[a];
    if (s_dead[g]) { __syncthreads(); continue; }        // uniform: s_dead was published by the barrier below
    if (threadIdx.x == 0) keep_rank[lo + g] = kept;
    ++kept;
    const float* kg = kpts + (size_t)(lo + g) * J * 3;
    for (int b = a + 1 + (int)threadIdx.x; b < n; b += blockDim.x) {
      const int d = s_order[b];
      if (s_dead[d]) continue;
      const float* kd = kpts + (size_t)(lo + d) * J * 3;
      const double denom = (area[lo + g] + area[lo + d]) / 2 + 2.220446049250313e-16;
      // e_j as the reference's array expression: differences, squares and their sum in float32 (keypoints are float32,
      // no FMA contraction), the divisions in float64; joints dropped by the visibility mask are compacted away first
      double e[kMaxJoints];
      int cnt = 0;
      for (int j = 0; j < J; ++j) {
        if (nms_vis_thr >= 0.f && !(kd[j * 3 + 2] > nms_vis_thr)) continue;
        const float dx = __fsub_rn(kd[j * 3], kg[j * 3]), dy = __fsub_rn(kd[j * 3 + 1], kg[j * 3 + 1]);
        const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        e[cnt++] = exp(-((double)d2 / vars[j] / denom / 2));
      }
      // np.sum of a contiguous float64 array of cnt <= J <= 64 elements: 8 running partial sums over the multiple-of-8 prefix,
      // combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail in order (plain loop below 8 elements)
      double sum = 0.0;
      if (cnt < 8) {
        for (int j = 0; j < cnt; ++j) sum += e[j];
      } else {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = e[k];
        int j = 8;
        for (; j < cnt - (cnt % 8); j += 8)
          for (int k = 0; k < 8; ++k) r[k] += e[j + k];
        sum = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; j < cnt; ++j) sum += e[j];
      }
      const double oks = cnt ? sum / cnt : 0.0;
      if (oks > oks_thr) s_dead[d] = 1;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ training targets
// JointsDataset.generate_target (data/JointsDataset.py:230-286) for a whole batch: joints / joints_vis [B][J][3] fp64
// (crop pixels) -> target fp32 [B][J][h][w], target_weight fp32 [B][J].  mu = int(joint / stride + 0.5) in fp64 like
// the reference; a (6*sigma+1)^2 unnormalised Gaussian patch exp(-(dx^2+dy^2)/(2 sigma^2)) in fp32; a joint whose patch
// lies entirely outside the map gets weight 0; invisible joints (weight <= 0.5) get an all-zero map.  One block per map.
__global__ void __launch_bounds__(256) generate_target_kernel(const double* __restrict__ joints,
                                                              const double* __restrict__ joints_vis,
                                                              const float* __restrict__ joints_weight, int J, int h,
                                                              int w, double stride_x, double stride_y, int sigma,
                                                              float* __restrict__ target, float* __restrict__ weight) {
  const long long map = blockIdx.x;
  const int j = (int)(map % J);
  const int rad = sigma * 3;
  const int mx = (int)(joints[map * 3 + 0] / stride_x + 0.5);       // C cast = Python int(): truncation towards zero
  const int my = (int)(joints[map * 3 + 1] / stride_y + 0.5);
  const int x0 = mx - rad, y0 = my - rad, x1 = mx + rad + 1, y1 = my + rad + 1;
  float wgt = (float)joints_vis[map * 3 + 0];
  const bool outside = x0 >= w || y0 >= h || x1 < 0 || y1 < 0;
  if (outside) wgt = 0.f;
  const bool draw = !outside && wgt > 0.5f;
  const float inv = 1.0f / (float)(2 * sigma * sigma);
  float* t = target + map * (long long)h * w;
  for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
    const int y = i / w, x = i - y * w;
    float v = 0.f;
    if (draw && x >= x0 && x < x1 && y >= y0 && y < y1) {
      const float dx = (float)(x - mx), dy = (float)(y - my);
      v = expf(-(dx * dx + dy * dy) * inv);
    }
    t[i] = v;
  }
  if (threadIdx.x == 0) weight[map] = joints_weight ? wgt * joints_weight[j] : wgt;
}

// ------------------------------------------------------------------ arg-max of the bilinearly upsampled heatmaps
// create_pose_from_outputs (lib/pose_parsing.py:138-151; 04_evaluate_vases_qualitatively.py:216-220,
// 05_create_archdata_retrieval_db.py:114,131-147): F.interpolate(dets, (256, 192), mode="bilinear", align_corners=True)
// followed by get_max_preds_hrnet.  One warp per (crop, joint) map evaluates the out_h x out_w samples on the fly --
// the 16x larger tensor is never materialised -- with torch's upsample_bilinear2d arithmetic (fp32: source index =
// o * (in-1)/(out-1), lambda = index - floor, value = l0h*(l0w*v00 + l1w*v01) + l1h*(l0w*v10 + l1w*v11), no fused
// multiply-adds) and np.argmax's first-index tie rule.
__global__ void __launch_bounds__(256) upsampled_argmax_kernel(const float* __restrict__ heat, int B, int J, int h, int w,
                                                               int out_h, int out_w, float* __restrict__ coords,
                                                               float* __restrict__ maxvals) {
  const int lane = threadIdx.x & 31;
  const long long map = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (map >= (long long)B * J) return;
  const float* a = heat + map * (long long)h * w;
  const float sh = out_h > 1 ? (float)(h - 1) / (float)(out_h - 1) : 0.f;
  const float sw = out_w > 1 ? (float)(w - 1) / (float)(out_w - 1) : 0.f;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int oy = 0; oy < out_h; ++oy) {
    const float h1r = __fmul_rn(sh, (float)oy);
    int h1 = (int)h1r;
    if (h1 > h - 1) h1 = h - 1;
    const int h1p = h1 < h - 1 ? 1 : 0;
    float l1h = __fsub_rn(h1r, (float)h1);
    l1h = fminf(fmaxf(l1h, 0.f), 1.f);
    const float l0h = __fsub_rn(1.f, l1h);
    const float* r0 = a + h1 * w;
    const float* r1 = r0 + h1p * w;
    for (int ox = lane; ox < out_w; ox += 32) {
      const float w1r = __fmul_rn(sw, (float)ox);
      int w1 = (int)w1r;
      if (w1 > w - 1) w1 = w - 1;
      const int w1p = w1 < w - 1 ? 1 : 0;
      float l1w = __fsub_rn(w1r, (float)w1);
      l1w = fminf(fmaxf(l1w, 0.f), 1.f);
      const float l0w = __fsub_rn(1.f, l1w);
      const float top = __fadd_rn(__fmul_rn(l0w, __ldg(r0 + w1)), __fmul_rn(l1w, __ldg(r0 + w1 + w1p)));
      const float bot = __fadd_rn(__fmul_rn(l0w, __ldg(r1 + w1)), __fmul_rn(l1w, __ldg(r1 + w1 + w1p)));
      const float v = __fadd_rn(__fmul_rn(l0h, top), __fmul_rn(l1h, bot));
      const int i = oy * out_w + ox;
      if (v > best) { best = v; best_i = i; }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
    if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  if (lane != 0) return;
  if (best_i == 0x7fffffff) best_i = 0;
  float cx = (float)(best_i % out_w), cy = (float)(best_i / out_w);      // pose_parsing.py:44-53
  if (!(best > 0.f)) { cx = 0.f; cy = 0.f; }
  coords[map * 2 + 0] = cx;
  coords[map * 2 + 1] = cy;
  maxvals[map] = best;
}

// ------------------------------------------------------------------ crop extraction (the step before the network)
// cv2.warpAffine(img, M, (out_w, out_h), flags=INTER_LINEAR) of lib/transforms.py:38-43 / JointsDataset.py:189-197 for
// uint8 HWC images, one launch for all boxes of an image: OpenCV's fixed-point algorithm -- source coordinates in
// 1/1024 px from the inverted matrix (float64, rounded to nearest even like cvRound), reduced to 1/32 px, bilinear
// weights (32-fx)(32-fy) in 15-bit fixed point, (sum + 2^14) >> 15, constant 0 outside the image -- so the crops are
// bit-identical to the reference's.  minv: [N][6] float64 (dst -> src).  Outputs (each optional): the uint8 crops as
// [N][3][out_h][out_w] (what TransformDetection returns) and the network input fp32 [N][3][out_h][out_w] =
// (v / 255 - mean[c]) / std[c] (ToTensor + Normalize, data_loaders.py:59-61) in the same float32 operation order.
struct Norm3 { float mean[3], stdv[3]; };
__global__ void __launch_bounds__(256) warp_affine_kernel(const uint8_t* __restrict__ img, int ih, int iw,
                                                          const double* __restrict__ minv, int out_h, int out_w,
                                                          uint8_t* __restrict__ out_u8, float* __restrict__ out_f,
                                                          const Norm3 nm) {
  const int n = blockIdx.z, y = blockIdx.y;
  const double* m = minv + (size_t)n * 6;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < out_w; x += gridDim.x * blockDim.x) {
    const long long adelta = __double2ll_rn(__dmul_rn(__dmul_rn(m[0], (double)x), 1024.0));
    const long long bdelta = __double2ll_rn(__dmul_rn(__dmul_rn(m[3], (double)x), 1024.0));
    const long long x0 = __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], (double)y), m[2]), 1024.0)) + 16;
    const long long y0 = __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], (double)y), m[5]), 1024.0)) + 16;
    const long long X = (x0 + adelta) >> 5, Y = (y0 + bdelta) >> 5;
    // OpenCV stores the integer part as a saturated short
    long long sxl = X >> 5, syl = Y >> 5;
    sxl = sxl < -32768 ? -32768 : (sxl > 32767 ? 32767 : sxl);
    syl = syl < -32768 ? -32768 : (syl > 32767 ? 32767 : syl);
    const int sx = (int)sxl, sy = (int)syl, fx = (int)(X & 31), fy = (int)(Y & 31);
    const int w00 = (32 - fx) * (32 - fy), w01 = fx * (32 - fy), w10 = (32 - fx) * fy, w11 = fx * fy;
    const bool r0 = sy >= 0 && sy < ih, r1 = sy + 1 >= 0 && sy + 1 < ih;
    const bool c0 = sx >= 0 && sx < iw, c1 = sx + 1 >= 0 && sx + 1 < iw;
    const uint8_t* p00 = img + ((size_t)sy * iw + sx) * 3;
    const uint8_t* p10 = p00 + (size_t)iw * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int v00 = (r0 && c0) ? p00[c] : 0, v01 = (r0 && c1) ? p00[3 + c] : 0;
      const int v10 = (r1 && c0) ? p10[c] : 0, v11 = (r1 && c1) ? p10[3 + c] : 0;
      const int v = ((v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11) * 32 + 16384) >> 15;
      const size_t o = (((size_t)n * 3 + c) * out_h + y) * out_w + x;
      if (out_u8) out_u8[o] = (uint8_t)v;
      if (out_f) out_f[o] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), nm.mean[c]), nm.stdv[c]);
    }
  }
}

// The same warp for float32 HWC images (04_evaluate_vases_qualitatively.py:209-213 passes float arrays to
// TransformDetection): OpenCV keeps the fixed-point SOURCE COORDINATES (1/32 px) but interpolates in float32 with the
// weight table w[fy][fx] = {(1-fy/32)(1-fx/32), (1-fy/32)(fx/32), (fy/32)(1-fx/32), (fy/32)(fx/32)} (float products of
// float table entries) and the sum evaluated left to right without contraction - reproduced operation by operation, so
// the crops are bit-identical to cv2's.  Output fp32 [N][3][out_h][out_w].
__global__ void __launch_bounds__(256) warp_affine_f32_kernel(const float* __restrict__ img, int ih, int iw,
                                                              const double* __restrict__ minv, int out_h, int out_w,
                                                              float* __restrict__ out) {
  const int n = blockIdx.z, y = blockIdx.y;
  const double* m = minv + (size_t)n * 6;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < out_w; x += gridDim.x * blockDim.x) {
    const long long adelta = __double2ll_rn(__dmul_rn(__dmul_rn(m[0], (double)x), 1024.0));
    const long long bdelta = __double2ll_rn(__dmul_rn(__dmul_rn(m[3], (double)x), 1024.0));
    const long long x0 = __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], (double)y), m[2]), 1024.0)) + 16;
    const long long y0 = __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], (double)y), m[5]), 1024.0)) + 16;
    const long long X = (x0 + adelta) >> 5, Y = (y0 + bdelta) >> 5;
    long long sxl = X >> 5, syl = Y >> 5;
    sxl = sxl < -32768 ? -32768 : (sxl > 32767 ? 32767 : sxl);
    syl = syl < -32768 ? -32768 : (syl > 32767 ? 32767 : syl);
    const int sx = (int)sxl, sy = (int)syl, fx = (int)(X & 31), fy = (int)(Y & 31);
    const float wx1 = __fmul_rn((float)fx, 0.03125f), wx0 = __fsub_rn(1.f, wx1);     // i * (1/32) is exact in float
    const float wy1 = __fmul_rn((float)fy, 0.03125f), wy0 = __fsub_rn(1.f, wy1);
    const float w00 = __fmul_rn(wy0, wx0), w01 = __fmul_rn(wy0, wx1), w10 = __fmul_rn(wy1, wx0), w11 = __fmul_rn(wy1, wx1);
    const bool r0 = sy >= 0 && sy < ih, r1 = sy + 1 >= 0 && sy + 1 < ih;
    const bool c0 = sx >= 0 && sx < iw, c1 = sx + 1 >= 0 && sx + 1 < iw;
    const float* p00 = img + ((long long)sy * iw + sx) * 3;
    const float* p10 = p00 + (long long)iw * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v00 = (r0 && c0) ? p00[c] : 0.f, v01 = (r0 && c1) ? p00[3 + c] : 0.f;
      const float v10 = (r1 && c0) ? p10[c] : 0.f, v11 = (r1 && c1) ? p10[3 + c] : 0.f;
      const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00, w00), __fmul_rn(v01, w01)), __fmul_rn(v10, w10)),
                                __fmul_rn(v11, w11));
      out[(((size_t)n * 3 + c) * out_h + y) * out_w + x] = v;
    }
  }
}

// ------------------------------------------------------------------ PCK accuracy on heatmap arg-max coordinates
// metrics.py:268-364 (calc_dists, dist_acc, accuracy): a joint counts when its target arg-max has x > 1 and y > 1; it is
// a hit when || (pred - target) / (h/10, w/10) || < thr (the reference divides x by h/10 and y by w/10); acc[1+j] =
// hits / counted or -1; acc[0] = mean over joints with acc >= 0.  One block, one warp per joint.
__global__ void pck_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int B, int J, double nx,
                           double ny, double thr, float* __restrict__ acc, float* __restrict__ avg_acc,
                           int* __restrict__ cnt) {
  __shared__ float s_acc[kMaxJoints];
  const int lane = threadIdx.x & 31;
  for (int j = threadIdx.x >> 5; j < J; j += blockDim.x >> 5) {
    int counted = 0, hits = 0;
    for (int n = lane; n < B; n += 32) {
      const float tx = tgt[((size_t)n * J + j) * 2], ty = tgt[((size_t)n * J + j) * 2 + 1];
      if (tx > 1.f && ty > 1.f) {
        const double dx = (double)pred[((size_t)n * J + j) * 2] / nx - (double)tx / nx;
        const double dy = (double)pred[((size_t)n * J + j) * 2 + 1] / ny - (double)ty / ny;
        ++counted;
        hits += sqrt(dx * dx + dy * dy) < thr;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      counted += __shfl_xor_sync(0xffffffffu, counted, off);
      hits += __shfl_xor_sync(0xffffffffu, hits, off);
    }
    if (lane == 0) s_acc[j] = counted > 0 ? (float)((double)hits / (double)counted) : -1.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sum = 0.0;
    int c = 0;
    for (int j = 0; j < J; ++j) {
      acc[j + 1] = s_acc[j];
      if (s_acc[j] >= 0.f) { sum += (double)s_acc[j]; ++c; }
    }
    const float avg = c ? (float)(sum / c) : 0.f;
    acc[0] = c ? avg : 0.f;
    *avg_acc = avg;
    *cnt = c;
  }
}

// x *= *scale unless *scale == 1 (the usual upstream gradient of a loss): then the kernel touches no memory.
__global__ void __launch_bounds__(256) scale_unless_one_kernel(float* __restrict__ x, const float* __restrict__ scale,
                                                               long long n4, long long n) {
  const float s = __ldg(scale);
  if (s == 1.0f) return;
  float4* x4 = reinterpret_cast<float4*>(x);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = x4[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    x4[i] = v;
  }
  if (blockIdx.x == 0)
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) x[i] *= s;
}

int make_perm(int J, const int* pairs, int n_pairs, Perm* perm) {
  if (J > kMaxJoints || J <= 0) { set_error("decode: J=%d out of range (max %d)", J, kMaxJoints); return 1; }
  for (int j = 0; j < kMaxJoints; ++j) perm->src[j] = j;
  for (int i = 0; i < n_pairs; ++i) {
    const int a = pairs[2 * i], b = pairs[2 * i + 1];
    if (a < 0 || a >= J || b < 0 || b >= J) { set_error("decode: flip pair (%d,%d) out of range", a, b); return 1; }
    const int t = perm->src[a];
    perm->src[a] = perm->src[b];
    perm->src[b] = t;
  }
  return 0;
}

int check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace

int flip_avg(const float* heat, const float* heat_f, float* out, int B, int J, int h, int w, const int* pairs,
             int n_pairs, cudaStream_t st) {
  Perm perm;
  if (make_perm(J, pairs, n_pairs, &perm)) return 1;
  const long long total = (long long)B * J * h * w;
  if (total == 0) return 0;
  long long g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  flip_avg_kernel<<<(int)g, 256, 0, st>>>(heat, heat_f, out, B, J, h, w, perm);
  return check("flip_avg");
}

int flip_back(const float* in, float* out, int B, int J, int h, int w, const int* pairs, int n_pairs,
              cudaStream_t st) {
  Perm perm;
  if (make_perm(J, pairs, n_pairs, &perm)) return 1;
  const long long total = (long long)B * J * h * w;
  if (total == 0) return 0;
  long long g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  flip_back_kernel<<<(int)g, 256, 0, st>>>(in, out, B, J, h, w, perm);
  return check("flip_back");
}

int decode(const float* heat, const float* heat_f, const float* center, const float* scale, int B, int J, int h,
           int w, const int* pairs, int n_pairs, int refine, float* avg_out, float* preds, float* maxvals,
           float* coords, cudaStream_t st) {
  Perm perm;
  if (make_perm(J, pairs, n_pairs, &perm)) return 1;
  if (h < 1 || w < 1) { set_error("decode: empty heatmap"); return 1; }
  const long long maps = (long long)B * J;
  if (maps == 0) return 0;
  const int warps_per_block = 8;
  decode_kernel<<<(unsigned)((maps + warps_per_block - 1) / warps_per_block), warps_per_block * 32, 0, st>>>(
      heat, heat_f, center, scale, B, J, h, w, perm, refine, avg_out, preds, maxvals, coords);
  return check("decode");
}

int oks_nms(const float* kpts, const double* area, const double* box_score, const int* offsets, int n_images,
            int max_persons, int J, const double* vars, float in_vis_thr, double oks_thr, float nms_vis_thr, int rescore,
            double* score_out, int* keep_rank, cudaStream_t st) {
  if (n_images <= 0) return 0;
  if (max_persons < 1) max_persons = 1;
  max_persons = (max_persons + 7) & ~7;                                   // keeps the int / byte arrays aligned
  const size_t smem = (size_t)max_persons * (sizeof(double) + sizeof(int) + 1);
  if (smem > 200 * 1024) { set_error("oks_nms: %d persons in one image exceed the shared-memory budget", max_persons); return 1; }
  if (smem > 48 * 1024) {
    static DeviceOnce attr_once;
    if (attr_once.run([]() {
          cudaError_t e = cudaFuncSetAttribute(oks_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
          if (e != cudaSuccess) { set_error("oks_nms attribute: %s", cudaGetErrorString(e)); return 1; }
          return 0;
        }))
      return 1;
  }
  oks_nms_kernel<<<n_images, 128, smem, st>>>(kpts, area, box_score, offsets, J, vars, in_vis_thr, oks_thr, nms_vis_thr,
                                             rescore, max_persons, score_out, keep_rank);
  return check("oks_nms");
}

int generate_target(const double* joints, const double* joints_vis, const float* joints_weight, int B, int J, int h, int w,
                    int image_h, int image_w, int sigma, float* target, float* weight, cudaStream_t st) {
  const long long maps = (long long)B * J;
  if (maps <= 0) return 0;
  if (h <= 0 || w <= 0 || sigma <= 0) { set_error("generate_target: bad geometry"); return 1; }
  generate_target_kernel<<<(unsigned)maps, 256, 0, st>>>(joints, joints_vis, joints_weight, J, h, w, (double)image_w / w,
                                                       (double)image_h / h, sigma, target, weight);
  return check("generate_target");
}

int upsampled_argmax(const float* heat, int B, int J, int h, int w, int out_h, int out_w, float* coords, float* maxvals,
                     cudaStream_t st) {
  const long long maps = (long long)B * J;
  if (maps <= 0) return 0;
  if (h <= 0 || w <= 0 || out_h <= 0 || out_w <= 0) { set_error("upsampled_argmax: bad geometry"); return 1; }
  upsampled_argmax_kernel<<<(unsigned)((maps + 7) / 8), 256, 0, st>>>(heat, B, J, h, w, out_h, out_w, coords, maxvals);
  return check("upsampled_argmax");
}

int warp_affine_crops(const uint8_t* img, int ih, int iw, const double* minv, int N, int out_h, int out_w,
                      uint8_t* out_u8, float* out_f, const float* mean3, const float* std3, cudaStream_t st) {
  if (N <= 0) return 0;
  if (ih <= 0 || iw <= 0 || out_h <= 0 || out_w <= 0 || out_h > 65535 || N > 65535) {
    set_error("warp_affine_crops: bad geometry");
    return 1;
  }
  Norm3 nm{};
  for (int c = 0; c < 3; ++c) { nm.mean[c] = mean3 ? mean3[c] : 0.f; nm.stdv[c] = std3 ? std3[c] : 1.f; }
  dim3 grid((out_w + 255) / 256, out_h, N);
  warp_affine_kernel<<<grid, 256, 0, st>>>(img, ih, iw, minv, out_h, out_w, out_u8, out_f, nm);
  return check("warp_affine_crops");
}

int warp_affine_crops_f32(const float* img, int ih, int iw, const double* minv, int N, int out_h, int out_w, float* out,
                          cudaStream_t st) {
  if (N <= 0) return 0;
  if (ih <= 0 || iw <= 0 || out_h <= 0 || out_w <= 0 || out_h > 65535 || N > 65535) {
    set_error("warp_affine_crops_f32: bad geometry");
    return 1;
  }
  dim3 grid((out_w + 255) / 256, out_h, N);
  warp_affine_f32_kernel<<<grid, 256, 0, st>>>(img, ih, iw, minv, out_h, out_w, out);
  return check("warp_affine_crops_f32");
}

int pck_accuracy(const float* pred, const float* tgt, int B, int J, int h, int w, float thr, float* acc, float* avg_acc,
                 int* cnt, cudaStream_t st) {
  if (J <= 0 || J > kMaxJoints) { set_error("pck_accuracy: J=%d out of range (max %d)", J, kMaxJoints); return 1; }
  pck_kernel<<<1, 32 * (J < 32 ? J : 32), 0, st>>>(pred, tgt, B, J, (double)h / 10.0, (double)w / 10.0, (double)thr, acc,
                                                  avg_acc, cnt);
  return check("pck_accuracy");
}

int scale_inplace(float* x, const float* scale_dev, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  scale_unless_one_kernel<<<(int)blocks, 256, 0, st>>>(x, scale_dev, n4, n);
  return check("scale_inplace");
}

size_t mse_workspace_bytes() { return sizeof(double) * 148 * 8; }

int mse_loss(const float* out, const float* tgt, const float* tw, int B, int J, int hw, float* loss, float* grad,
             void* workspace, cudaStream_t st) {
  if (hw % 4) { set_error("mse_loss: h*w must be a multiple of 4 (got %d)", hw); return 1; }
  const long long n4 = (long long)B * J * hw / 4;
  const double denom = (double)J * B * hw;
  int blocks = 148 * 8;
  if (n4 < (long long)blocks * kLossBlock) blocks = (int)((n4 + kLossBlock - 1) / kLossBlock);
  if (blocks < 1) blocks = 1;
  double* partial = reinterpret_cast<double*>(workspace);
  mse_partial_kernel<<<blocks, kLossBlock, 0, st>>>(out, tgt, tw, n4, hw / 4, (float)(1.0 / denom), grad, partial);
  if (check("mse_partial")) return 1;
  mse_final_kernel<<<1, 256, 0, st>>>(partial, blocks, 0.5 / denom, loss);
  return check("mse_final");
}

}  // namespace stl

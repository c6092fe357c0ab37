// Bandwidth-bound kernels of the keypoint path: flip-test averaging, heatmap decode, joints-MSE loss.
//
// References (paths relative to /root/reference/src):
//   flip_back + shift + average : lib/transforms.py:147-164, lib/inference.py:21-26
//   argmax / refine / affine    : lib/pose_parsing.py:16-92, lib/transforms.py:184-240
//   PersonMSELoss               : lib/loss.py:61-94
#include <cuda_runtime.h>
#include <stdint.h>

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

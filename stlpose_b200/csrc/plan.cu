// Whole-network plan: HRNet topology -> static schedule of fused convolution launches.
//
// Mirrors the wiring of /root/reference/src/models/HRnet.py (PoseHighResolutionNet.forward :433-468,
// HighResolutionModule.forward :248-266, transitions :341-380) but as a flat list of device ops over pooled
// activation buffers; BatchNorm is folded into the packed weights, ReLU / residual / fuse-layer additions live in
// the convolution epilogues.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "aux_kernels.h"
#include "conv.h"
#include "plan.h"

namespace stl {

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct Plan::Builder {
  Plan& P;
  std::map<std::tuple<int, int, int>, std::vector<int>> free_slots;

  explicit Builder(Plan& p) : P(p) {}

  int layer(const std::string& conv_key, const std::string& bn_key, int cout, int cin, int k, int stride,
            bool im2col = false) {
    Layer L;
    L.conv_key = conv_key;
    L.bn_key = bn_key;
    L.cout = cout; L.cin = cin; L.k = k; L.stride = stride;
    L.im2col = im2col;
    L.cout_pad = (int)align_up(cout, 16);
    L.cin_pad = im2col ? 32 : (int)align_up(cin, 16);   // (without im2col the 3-channel input is padded to one K-step)
    L.w_off = P.weight_bytes;
    P.weight_bytes += align_up((size_t)(im2col ? 1 : k * k) * L.cout_pad * L.cin_pad * 2, 256);
    L.b_off = P.weight_bytes;
    P.weight_bytes += align_up(sizeof(float) * L.cout_pad, 256);
    P.layers.push_back(L);
    return (int)P.layers.size() - 1;
  }

  int acquire(int C, int H, int W) {
    auto key = std::make_tuple(C, H, W);
    auto& fl = free_slots[key];
    if (!fl.empty()) {
      int id = fl.back();
      fl.pop_back();
      return id;
    }
    Plan::Slot s;
    s.C = C; s.H = H; s.W = W;
    P.slots.push_back(s);
    return (int)P.slots.size() - 1;
  }
  void release(int id) {
    const Plan::Slot& s = P.slots[id];
    free_slots[std::make_tuple(s.C, s.H, s.W)].push_back(id);
  }

  void conv(int layer, int in, int out, int res, bool relu, int n_up = 0, const int* up = nullptr,
            const int* up_shift = nullptr, bool out_nchw = false) {
    Plan::Op op;
    op.kind = Plan::OP_CONV;
    op.layer = layer; op.in = in; op.out = out; op.res = res; op.relu = relu; op.out_nchw = out_nchw;
    op.n_up = n_up;
    for (int i = 0; i < n_up; ++i) { op.up[i] = up[i]; op.up_shift[i] = up_shift[i]; }
    P.ops.push_back(op);
  }
  void fuse(int x, int out, int n_up, const int* up, const int* up_shift) {
    Plan::Op op;
    op.kind = Plan::OP_FUSE;
    op.in = x; op.out = out; op.n_up = n_up; op.relu = true;
    for (int i = 0; i < n_up; ++i) { op.up[i] = up[i]; op.up_shift[i] = up_shift[i]; }
    P.ops.push_back(op);
  }

  void build() {
    const stl_hrnet_cfg& c = P.cfg;
    const int H4 = c.image_h / 4, W4 = c.image_w / 4;
    const int ch[4] = {c.width, 2 * c.width, 4 * c.width, 8 * c.width};
    auto bh = [&](int b) { return H4 >> b; };
    auto bw = [&](int b) { return W4 >> b; };
    char k1[128], k2[128];

    // stem (HRnet.py:434-439)
    const bool im2col = P.stem_im2col != 0;
    int xin = im2col ? acquire(32, c.image_h / 2, c.image_w / 2) : acquire(16, c.image_h, c.image_w);
    {
      Plan::Op op;
      op.kind = Plan::OP_STEM;   // fp32 NCHW network input -> im2col rows of conv1 (or 16-channel padded NHWC), flip half mirrored
      op.out = xin;
      P.ops.push_back(op);
    }
    int t0 = acquire(64, c.image_h / 2, c.image_w / 2);
    conv(layer("conv1", "bn1", 64, 3, 3, 2, im2col), xin, t0, -1, true);
    release(xin);
    int x = acquire(64, H4, W4);
    conv(layer("conv2", "bn2", 64, 64, 3, 2), t0, x, -1, true);
    release(t0);

    // layer1: 4 Bottlenecks (HRnet.py:297, 64-102)
    int a_next = -1;   // conv1 output of the coming block when the previous junction already computed it (OP_LINK)
    int l1_next = -1;
    for (int b = 0; b < 4; ++b) {
      const int cin = b == 0 ? 64 : 256;
      snprintf(k1, sizeof k1, "layer1.%d", b);
      const std::string p = k1;
      const int l1 = l1_next >= 0 ? l1_next : layer(p + ".conv1", p + ".bn1", 64, cin, 1, 1);
      int a = a_next;
      if (a < 0) {
        a = acquire(64, H4, W4);
        conv(l1, x, a, -1, true);
      }
      a_next = l1_next = -1;
      int bb = acquire(64, H4, W4);
      conv(layer(p + ".conv2", p + ".bn2", 64, 64, 3, 1), a, bb, -1, true);
      release(a);
      int res = x;
      if (b == 0 && P.fuse_downsample) {
        // out = relu(bn3(conv3(bb)) + bn_d(downsample(x))) as one 1x1 convolution over K = [bb | x]: the 256-channel
        // downsample tensor is neither written nor read back (3.3 GB of HBM traffic per 1 024 images at 64x48)
        const int ld = layer(p + ".downsample.0", p + ".downsample.1", 256, 64, 1, 1);
        const int l3 = layer(p + ".conv3", p + ".bn3", 256, 64, 1, 1);
        Plan::Cat cat;
        cat.la = l3; cat.lb = ld;
        cat.w_off = P.weight_bytes;
        P.weight_bytes += align_up((size_t)P.layers[l3].cout_pad * (P.layers[l3].cin_pad + P.layers[ld].cin_pad) * 2, 256);
        cat.b_off = P.weight_bytes;
        P.weight_bytes += align_up(sizeof(float) * P.layers[l3].cout_pad, 256);
        P.cats.push_back(cat);
        int o = acquire(256, H4, W4);
        if (P.fuse_links && bottleneck_link_supported(64, 256, 64)) {
          // ... and conv1 of layer1.1 in the same kernel (link_tc.cu, two-input variant)
          l1_next = layer("layer1.1.conv1", "layer1.1.bn1", 64, 256, 1, 1);
          a_next = acquire(64, H4, W4);
          Plan::Op op;
          op.kind = Plan::OP_LINK;
          op.layer = l3; op.layer2 = l1_next; op.in = bb; op.in2 = x; op.cat = (int)P.cats.size() - 1;
          op.out = o; op.out2 = a_next; op.relu = true;
          P.ops.push_back(op);
        } else {
          conv(l3, bb, o, -1, true);
          P.ops.back().in2 = x;
          P.ops.back().cat = (int)P.cats.size() - 1;
        }
        release(bb);
        release(x);
        x = o;
        continue;
      }
      if (b == 0) {
        res = acquire(256, H4, W4);
        conv(layer(p + ".downsample.0", p + ".downsample.1", 256, 64, 1, 1), x, res, -1, false);
        release(x);
      }
      int o = acquire(256, H4, W4);
      const int l3 = layer(p + ".conv3", p + ".bn3", 256, 64, 1, 1);
      if (b >= 1 && b < 3 && P.fuse_links && bottleneck_link_supported(64, 256, 64)) {
        // conv3 of this block and conv1 of the next in one kernel (link_tc.cu): the 256-channel output is written once
        // and not read back by the next block's conv1
        snprintf(k2, sizeof k2, "layer1.%d", b + 1);
        const std::string pn = k2;
        l1_next = layer(pn + ".conv1", pn + ".bn1", 64, 256, 1, 1);
        a_next = acquire(64, H4, W4);
        Plan::Op op;
        op.kind = Plan::OP_LINK;
        op.layer = l3; op.layer2 = l1_next; op.in = bb; op.res = res; op.out = o; op.out2 = a_next; op.relu = true;
        P.ops.push_back(op);
      } else {
        conv(l3, bb, o, res, true);
      }
      release(bb);
      release(res);
      x = o;
    }

    // transition1 (HRnet.py:442-447)
    std::vector<int> xs(2);
    xs[0] = acquire(ch[0], bh(0), bw(0));
    conv(layer("transition1.0.0", "transition1.0.1", ch[0], 256, 3, 1), x, xs[0], -1, true);
    xs[1] = acquire(ch[1], bh(1), bw(1));
    conv(layer("transition1.1.0.0", "transition1.1.0.1", ch[1], 256, 3, 2), x, xs[1], -1, true);
    release(x);

    for (int stage = 2; stage <= 4; ++stage) {
      const int nb = stage;
      if (stage > 2) {  // new lowest-resolution branch from the previous stage's last output (HRnet.py:450-463)
        snprintf(k1, sizeof k1, "transition%d.%d.0.0", stage - 1, nb - 1);
        snprintf(k2, sizeof k2, "transition%d.%d.0.1", stage - 1, nb - 1);
        int nbuf = acquire(ch[nb - 1], bh(nb - 1), bw(nb - 1));
        conv(layer(k1, k2, ch[nb - 1], ch[nb - 2], 3, 2), xs[nb - 2], nbuf, -1, true);
        xs.push_back(nbuf);
      }
      const int n_mod = c.stage_modules[stage - 2];
      for (int m = 0; m < n_mod; ++m) {
        const int n_out = (stage == 4 && m == n_mod - 1) ? 1 : nb;  // HRnet.py:413-416
        snprintf(k1, sizeof k1, "stage%d.%d", stage, m);
        const std::string mp = k1;
        // branches: `blocks` BasicBlocks each (HRnet.py:32-61, 252-253).  They are independent of each other and CAN run
        // concurrently, each on its share of the SMs (Plan::branch_streams, measured slower and off by default)
        static const int kShares[5][4] = {{0, 0, 0, 0}, {148, 0, 0, 0}, {84, 64, 0, 0}, {62, 48, 38, 0}, {50, 36, 30, 32}};
        const int par_id = P.n_par_groups++;
        for (int b = 0; b < nb; ++b) {
          const size_t branch_first_op = P.ops.size();
          // the 2 x blocks convs of a branch form a group: executed sub-batch by sub-batch so that the three
          // activation buffers they cycle through stay resident in L2 (bind() picks the sub-batch size)
          // 32-channel branches: each BasicBlock is ONE kernel (block_tc.cu), the intermediate never leaves the SM
          const bool fused = P.fuse_blocks && basic_block_supported(bh(b), bw(b), ch[b]);
          Plan::Group grp;
          grp.first = (int)P.ops.size();
          for (int k = 0; k < c.blocks; ++k) {
            snprintf(k1, sizeof k1, "%s.branches.%d.%d", mp.c_str(), b, k);
            const std::string bp = k1;
            const int l1 = layer(bp + ".conv1", bp + ".bn1", ch[b], ch[b], 3, 1);
            const int l2 = layer(bp + ".conv2", bp + ".bn2", ch[b], ch[b], 3, 1);
            int o = acquire(ch[b], bh(b), bw(b));
            if (fused) {
              Plan::Op op;
              op.kind = Plan::OP_BLOCK;
              op.layer = l1; op.layer2 = l2; op.in = xs[b]; op.out = o; op.res = xs[b]; op.relu = true;
              P.ops.push_back(op);
            } else {
              int tmp = acquire(ch[b], bh(b), bw(b));
              conv(l1, xs[b], tmp, -1, true);
              conv(l2, tmp, o, xs[b], true);
              release(tmp);
            }
            release(xs[b]);
            xs[b] = o;
          }
          if (!fused) {
            grp.count = (int)P.ops.size() - grp.first;
            for (int i = 0; i < grp.count; ++i) P.ops[grp.first + i].group = (int)P.groups.size();
            P.groups.push_back(grp);
          }
          for (size_t i = branch_first_op; i < P.ops.size(); ++i) {
            P.ops[i].par_group = par_id;
            P.ops[i].stream = b;
            P.ops[i].sm_share = kShares[nb][b];
          }
        }
        // fuse layers (HRnet.py:188-243, 255-264)
        std::vector<int> ys(n_out);
        for (int i = 0; i < n_out; ++i) {
          int z[kMaxUp], zs[kMaxUp], nz = 0;
          for (int j = i + 1; j < nb; ++j) {
            snprintf(k1, sizeof k1, "%s.fuse_layers.%d.%d.0", mp.c_str(), i, j);
            snprintf(k2, sizeof k2, "%s.fuse_layers.%d.%d.1", mp.c_str(), i, j);
            z[nz] = acquire(ch[i], bh(j), bw(j));
            zs[nz] = j - i;
            conv(layer(k1, k2, ch[i], ch[j], 1, 1), xs[j], z[nz], -1, false);
            ++nz;
          }
          int y = acquire(ch[i], bh(i), bw(i));
          if (i == 0) {
            fuse(xs[0], y, nz, z, zs);
          } else {
            int running = xs[i];
            for (int j = 0; j < i; ++j) {
              int t = xs[j];
              for (int k = 0; k < i - j; ++k) {
                const bool last = k == i - j - 1;
                snprintf(k1, sizeof k1, "%s.fuse_layers.%d.%d.%d.0", mp.c_str(), i, j, k);
                snprintf(k2, sizeof k2, "%s.fuse_layers.%d.%d.%d.1", mp.c_str(), i, j, k);
                if (!last) {
                  int t2 = acquire(ch[j], bh(j + k + 1), bw(j + k + 1));
                  conv(layer(k1, k2, ch[j], ch[j], 3, 2), t, t2, -1, true);
                  if (t != xs[j]) release(t);
                  t = t2;
                } else {
                  const bool final_term = j == i - 1;
                  conv(layer(k1, k2, ch[i], ch[j], 3, 2), t, y, running, final_term, final_term ? nz : 0, z, zs);
                  if (t != xs[j]) release(t);
                  running = y;
                }
              }
            }
          }
          for (int u = 0; u < nz; ++u) release(z[u]);
          ys[i] = y;
        }
        for (int b = 0; b < nb; ++b) release(xs[b]);
        if (n_out == nb) {
          xs = ys;
        } else {
          xs.assign(1, ys[0]);
        }
      }
    }
    // head (HRnet.py:331-337, 466): 1x1 conv with bias, no activation, fp32 NCHW out
    const int lh = layer("final_layer", "", c.num_joints, ch[0], 1, 1);
    Plan::Op& last = P.ops.back();
    if (P.fuse_head && last.kind == Plan::OP_FUSE && last.out == xs[0] && ch[0] == 32 && c.num_joints <= 32 && last.n_up >= 1) {
      // the last fuse row feeds nothing but the head: both in one pass, the fused map is never written (aux_kernels.cu)
      last.layer = lh;
      last.out_nchw = true;
    } else {
      conv(lh, xs[0], -1, -1, false, 0, nullptr, nullptr, true);
    }
    release(xs[0]);
  }
};

Plan* Plan::create(const stl_hrnet_cfg& cfg) {
  if (cfg.width < 16 || cfg.width % 16 || cfg.num_joints < 1 || cfg.num_joints > 64 || cfg.blocks < 1 ||
      cfg.image_h % 32 || cfg.image_w % 32 || cfg.image_h < 32 || cfg.image_w < 32) {
    set_error("plan: unsupported config (width %d joints %d blocks %d image %dx%d)", cfg.width, cfg.num_joints,
              cfg.blocks, cfg.image_h, cfg.image_w);
    return nullptr;
  }
  for (int i = 0; i < 3; ++i)
    if (cfg.stage_modules[i] < 1) { set_error("plan: stage_modules must be >= 1"); return nullptr; }
  Plan* p = new Plan();
  p->cfg = cfg;
  if (const char* e = getenv("STLPOSE_FUSE_BLOCK")) p->fuse_blocks = atoi(e);
  if (const char* e = getenv("STLPOSE_FUSE_DOWNSAMPLE")) p->fuse_downsample = atoi(e);
  if (const char* e = getenv("STLPOSE_FUSE_LINK")) p->fuse_links = atoi(e);
  if (const char* e = getenv("STLPOSE_FUSE_HEAD")) p->fuse_head = atoi(e);
  if (const char* e = getenv("STLPOSE_BRANCH_STREAMS")) p->branch_streams = atoi(e);
  if (const char* e = getenv("STLPOSE_STEM_IM2COL")) p->stem_im2col = atoi(e);
  Builder b(*p);
  b.build();
  return p;
}

size_t Plan::workspace_bytes(int n_images) const {
  size_t total = 0;
  for (const Slot& s : slots) {
    PaddedGeom g{n_images, s.H, s.W, s.C};
    total += align_up(g.bytes(), 1024);
  }
  return total;
}

int Plan::pack_conv(int index, const float* w, const float* gamma, const float* beta, const float* mean,
                    const float* var, const float* cbias, float eps, void* arena, cudaStream_t st) {
  if (index < 0 || index >= (int)layers.size()) { set_error("pack_conv: index %d out of range", index); return 1; }
  const Layer& L = layers[index];
  uint8_t* base = reinterpret_cast<uint8_t*>(arena);
  bound = false;  // packed parameters changed; tensor maps stay valid but be conservative
  if (L.im2col)   // OIHW [cout][cin][k][k] read as a 1x1 filter over cin*k*k "channels" (the im2col K order)
    return pack_weights(w, gamma, beta, mean, var, cbias, eps, L.cout, L.cin * L.k * L.k, 1, L.cout_pad, L.cin_pad,
                        reinterpret_cast<__nv_bfloat16*>(base + L.w_off), reinterpret_cast<float*>(base + L.b_off), st);
  if (pack_weights(w, gamma, beta, mean, var, cbias, eps, L.cout, L.cin, L.k, L.cout_pad, L.cin_pad,
                   reinterpret_cast<__nv_bfloat16*>(base + L.w_off), reinterpret_cast<float*>(base + L.b_off), st))
    return 1;
  // member of a concatenated pair: rebuild [cout_pad][cin_a | cin_b] and the summed bias from both members' packed
  // parameters (stream-ordered behind the packing above; the other member is (re)packed by its own call)
  for (const Cat& c : cats) {
    if (index != c.la && index != c.lb) continue;
    const Layer& A = layers[c.la];
    const Layer& B = layers[c.lb];
    const size_t pitch = (size_t)(A.cin_pad + B.cin_pad) * 2;
    cudaError_t e = cudaMemcpy2DAsync(base + c.w_off, pitch, base + A.w_off, (size_t)A.cin_pad * 2, (size_t)A.cin_pad * 2,
                                      A.cout_pad, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(base + c.w_off + (size_t)A.cin_pad * 2, pitch, base + B.w_off, (size_t)B.cin_pad * 2,
                            (size_t)B.cin_pad * 2, B.cout_pad, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { set_error("pack_conv: concatenating weights: %s", cudaGetErrorString(e)); return 1; }
    if (add_f32(reinterpret_cast<const float*>(base + A.b_off), reinterpret_cast<const float*>(base + B.b_off),
                reinterpret_cast<float*>(base + c.b_off), A.cout_pad, st))
      return 1;
  }
  return 0;
}

int Plan::bind(int n_images, const void* arena, void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (bound && n_images == bound_images && arena == bound_arena && workspace == bound_ws) return 0;
  const size_t need = workspace_bytes(n_images);
  if (ws_bytes < need) { set_error("plan: workspace too small (%zu < %zu)", ws_bytes, need); return 1; }
  if (reinterpret_cast<uintptr_t>(workspace) % 1024 || reinterpret_cast<uintptr_t>(arena) % 256) {
    set_error("plan: workspace must be 1024-byte aligned and the weight arena 256-byte aligned");
    return 1;
  }
  size_t off = 0;
  slot_ptr.resize(slots.size());
  for (size_t i = 0; i < slots.size(); ++i) {
    PaddedGeom g{n_images, slots[i].H, slots[i].W, slots[i].C};
    slot_ptr[i] = reinterpret_cast<uint8_t*>(workspace) + off;
    off += align_up(g.bytes(), 1024);
  }
  // zero cells of the padded layout are established once; no kernel ever writes a non-zero there
  cudaError_t e = cudaMemsetAsync(workspace, 0, need, st);
  if (e != cudaSuccess) { set_error("plan: memset workspace: %s", cudaGetErrorString(e)); return 1; }

  // sub-batch count of every group: keep one activation tensor of the group near `sub_bytes` so that the input,
  // intermediate and output buffers of a branch (3 tensors) fit in the 126 MB L2 together
  // Measured on B200 (round 1): with one launch per layer the fixed cost of a launch (pipeline fill, TMEM/barrier
  // set-up, weight load: ~8 us) outweighs the L2 hits for every sub-batch size tried (14/28/56 MB: 35.2/34.5/33.4 ms
  // per step vs 32.3 ms unsplit), so the default is off; the mechanism stays for the multi-layer kernels to come.
  size_t sub_bytes = 0;
  if (const char* e = getenv("STLPOSE_SUB_MB")) { long v = atol(e); sub_bytes = v > 0 ? (size_t)v << 20 : 0; }
  std::vector<int> group_subs(groups.size(), 1);
  for (size_t g = 0; g < groups.size(); ++g) {
    const Slot& so = slots[ops[groups[g].first].out];
    const size_t tensor = PaddedGeom{n_images, so.H, so.W, so.C}.bytes();
    int subs = sub_bytes ? (int)((tensor + sub_bytes / 2) / sub_bytes) : 1;
    if (subs > 8) subs = 8;
    // a sub-batch must still fill the machine a few times over: >= 4 tiles of 384 rows per SM
    const long long rows = (long long)n_images * (so.H + 1) * (so.W + 1);
    while (subs > 1 && rows / subs < 4ll * 148 * 384) --subs;
    if (subs < 1) subs = 1;
    if (subs > n_images) subs = n_images;
    group_subs[g] = subs;
  }
  const uint8_t* wbase = reinterpret_cast<const uint8_t*>(arena);
  // programmatic dependent launch between consecutive layers (the weight arena is static during a forward)
  use_pdl = getenv("STLPOSE_PDL") ? atoi(getenv("STLPOSE_PDL")) : 1;
  prepared.assign(ops.size(), std::vector<Prepared>());
  for (size_t i = 0; i < ops.size(); ++i) {
    const Op& op = ops[i];
    if (op.kind != OP_CONV) continue;
    const Layer& L = layers[op.layer];
    const Slot& si = slots[op.in];
    ConvSpec s;
    s.in = reinterpret_cast<const __nv_bfloat16*>(slot_ptr[op.in]);
    s.in_geom = PaddedGeom{n_images, si.H, si.W, si.C};
    s.out = op.out >= 0 ? slot_ptr[op.out] : nullptr;  // head output is patched per forward
    s.cout = L.cout;
    s.cout_pad = L.cout_pad;
    s.ksize = L.im2col ? 1 : L.k;
    s.stride = L.im2col ? 1 : L.stride;
    s.weights = reinterpret_cast<const __nv_bfloat16*>(wbase + L.w_off);
    s.bias = reinterpret_cast<const float*>(wbase + L.b_off);
    if (op.cat >= 0) {
      s.weights = reinterpret_cast<const __nv_bfloat16*>(wbase + cats[op.cat].w_off);
      s.bias = reinterpret_cast<const float*>(wbase + cats[op.cat].b_off);
      s.in2 = reinterpret_cast<const __nv_bfloat16*>(slot_ptr[op.in2]);
      s.in2_C = slots[op.in2].C;
    }
    s.residual = op.res >= 0 ? reinterpret_cast<const __nv_bfloat16*>(slot_ptr[op.res]) : nullptr;
    s.n_up = op.n_up;
    for (int u = 0; u < op.n_up; ++u) {
      s.up_src[u] = reinterpret_cast<const __nv_bfloat16*>(slot_ptr[op.up[u]]);
      s.up_shift[u] = op.up_shift[u];
    }
    s.relu = op.relu;
    s.out_nchw = op.out_nchw;
    s.force_tap_reload = tap_reload;
    s.pdl = use_pdl;
    const int subs = op.group >= 0 ? group_subs[op.group] : 1;
    prepared[i].resize(subs);
    const int chunk = (n_images + subs - 1) / subs;
    for (int sb = 0; sb < subs; ++sb) {
      if (subs > 1) {
        s.img_lo = sb * chunk;
        s.img_hi = s.img_lo + chunk < n_images ? s.img_lo + chunk : n_images;
      }
      Prepared& pr = prepared[i][sb];
      if (s.img_lo >= n_images && subs > 1) { pr.grid = 0; continue; }
      if (conv_prepare(s, &pr.params, &pr.grid, &pr.smem)) return 1;
      pr.grid_par = pr.grid;
      if (op.sm_share > 0 && pr.grid > op.sm_share)
        pr.grid_par = pr.params.pair ? 2 * (op.sm_share / 2) : op.sm_share;
    }
  }
  launches.clear();
  for (size_t i = 0; i < ops.size();) {
    const Op& op = ops[i];
    if (op.group >= 0 && groups[op.group].first == (int)i) {
      const Group& g = groups[op.group];
      for (int sb = 0; sb < group_subs[op.group]; ++sb)
        for (int j = 0; j < g.count; ++j) launches.push_back(Launch{g.first + j, sb});
      i += g.count;
    } else {
      launches.push_back(Launch{(int)i, 0});
      ++i;
    }
  }
  bound = true;
  bound_images = n_images;
  bound_arena = arena;
  bound_ws = workspace;
  return 0;
}

int Plan::forward(const float* x, int B, int flip_pair, float* heat, const void* arena, void* workspace,
                  size_t ws_bytes, cudaStream_t st, float* op_ms_host) {
  if (B <= 0) { set_error("plan: batch must be positive"); return 1; }
  const int n_images = flip_pair ? 2 * B : B;
  if (bind(n_images, arena, workspace, ws_bytes, st)) return 1;
  const uint8_t* wbase = reinterpret_cast<const uint8_t*>(arena);
  std::vector<cudaEvent_t> ev;
  if (op_ms_host) {
    ev.resize(launches.size() + 1);
    for (auto& e : ev) cudaEventCreate(&e);
    cudaEventRecord(ev[0], st);
  }
  struct EvGuard {
    std::vector<cudaEvent_t>& v;
    ~EvGuard() { for (auto e : v) cudaEventDestroy(e); }
  } guard{ev};
  // branches of a HighResolutionModule run concurrently on side streams (forked from / joined back into `st`, which
  // also works under stream capture); a timed forward stays sequential so that per-launch times mean something
  const bool par = branch_streams && !op_ms_host;
  if (par && !ev_fork) {
    cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
    for (int s = 1; s < 4; ++s) {
      cudaStreamCreateWithFlags(&side[s], cudaStreamNonBlocking);
      cudaEventCreateWithFlags(&ev_join[s], cudaEventDisableTiming);
    }
  }
  // (STLPOSE_PDL_FUSED=0: programmatic dependent launch only between the plain convolutions, not for the fused kernels)
  const bool pdl_fused = use_pdl && !op_ms_host && !(getenv("STLPOSE_PDL_FUSED") && atoi(getenv("STLPOSE_PDL_FUSED")) == 0);
  int cur_par = -1;
  unsigned used = 0;   // side streams used by the current module
  auto join = [&]() {
    for (int s = 1; s < 4; ++s)
      if (used & (1u << s)) { cudaEventRecord(ev_join[s], side[s]); cudaStreamWaitEvent(st, ev_join[s], 0); }
    used = 0;
    cur_par = -1;
  };
  cudaStream_t main_st = st;
  for (size_t li = 0; li < launches.size(); ++li) {
    const int i = launches[li].op;
    const Op& op = ops[i];
    st = main_st;
    if (par) {
      if (op.par_group != cur_par) {
        if (cur_par >= 0) join();
        if (op.par_group >= 0) { cudaEventRecord(ev_fork, main_st); cur_par = op.par_group; }
      }
      if (op.par_group >= 0 && op.stream > 0) {
        if (!(used & (1u << op.stream))) { cudaStreamWaitEvent(side[op.stream], ev_fork, 0); used |= 1u << op.stream; }
        st = side[op.stream];
      }
    }
    const bool limited = par && op.par_group >= 0;
    switch (op.kind) {
      case OP_STEM: {
        if (stem_im2col ? stl::stem_im2col(x, reinterpret_cast<__nv_bfloat16*>(slot_ptr[op.out]), n_images, B, cfg.image_h,
                                           cfg.image_w, st)
                        : stem_pack_input(x, reinterpret_cast<__nv_bfloat16*>(slot_ptr[op.out]), n_images, B, cfg.image_h,
                                          cfg.image_w, st))
          return 1;
        break;
      }
      case OP_CONV: {
        Prepared& pr = prepared[i][launches[li].sub];
        if (op.out_nchw) pr.params.out = heat;
        if (conv_launch_prepared(pr.params, limited ? pr.grid_par : pr.grid, pr.smem, st)) return 1;
        break;
      }
      case OP_BLOCK: {
        const Slot& so = slots[op.out];
        const Layer& L1 = layers[op.layer];
        const Layer& L2 = layers[op.layer2];
        if (basic_block_launch(reinterpret_cast<const __nv_bfloat16*>(slot_ptr[op.in]),
                               reinterpret_cast<__nv_bfloat16*>(slot_ptr[op.out]),
                               reinterpret_cast<const __nv_bfloat16*>(wbase + L1.w_off),
                               reinterpret_cast<const float*>(wbase + L1.b_off),
                               reinterpret_cast<const __nv_bfloat16*>(wbase + L2.w_off),
                               reinterpret_cast<const float*>(wbase + L2.b_off), n_images, so.H, so.W,
                               limited ? op.sm_share : 0, st, pdl_fused && !limited))
          return 1;
        break;
      }
      case OP_LINK: {
        const Slot& so = slots[op.out];
        const Layer& L3 = layers[op.layer];
        const Layer& L1 = layers[op.layer2];
        const bool two_in = op.cat >= 0;   // layer1.0 -> .1: concatenated conv3 | downsample weights, no residual
        if (bottleneck_link_launch(reinterpret_cast<const __nv_bfloat16*>(slot_ptr[op.in]),
                                   two_in ? reinterpret_cast<const __nv_bfloat16*>(slot_ptr[op.in2]) : nullptr,
                                   two_in ? nullptr : reinterpret_cast<const __nv_bfloat16*>(slot_ptr[op.res]),
                                   reinterpret_cast<__nv_bfloat16*>(slot_ptr[op.out]),
                                   reinterpret_cast<__nv_bfloat16*>(slot_ptr[op.out2]),
                                   reinterpret_cast<const __nv_bfloat16*>(wbase + (two_in ? cats[op.cat].w_off : L3.w_off)),
                                   reinterpret_cast<const float*>(wbase + (two_in ? cats[op.cat].b_off : L3.b_off)),
                                   reinterpret_cast<const __nv_bfloat16*>(wbase + L1.w_off),
                                   reinterpret_cast<const float*>(wbase + L1.b_off), n_images, so.H, so.W, 0, st,
                                   pdl_fused))
          return 1;
        break;
      }
      case OP_FUSE: {
        const Slot& so = slots[op.out];
        const __nv_bfloat16* z[kMaxUp];
        for (int u = 0; u < op.n_up; ++u) z[u] = reinterpret_cast<const __nv_bfloat16*>(slot_ptr[op.up[u]]);
        if (op.layer >= 0) {   // + the heatmap head
          const Layer& Lh = layers[op.layer];
          if (stl::fuse_head(reinterpret_cast<const __nv_bfloat16*>(slot_ptr[op.in]), z, op.up_shift, op.n_up,
                        reinterpret_cast<const __nv_bfloat16*>(wbase + Lh.w_off), reinterpret_cast<const float*>(wbase + Lh.b_off),
                        heat, n_images, so.H, so.W, so.C, Lh.cout, st))
            return 1;
          break;
        }
        if (fuse_sum(reinterpret_cast<const __nv_bfloat16*>(slot_ptr[op.in]), z, op.up_shift, op.n_up,
                     reinterpret_cast<__nv_bfloat16*>(slot_ptr[op.out]), n_images, so.H, so.W, so.C, st))
          return 1;
        break;
      }
    }
    if (op_ms_host) cudaEventRecord(ev[li + 1], st);
  }
  st = main_st;
  if (par && cur_par >= 0) join();
  if (op_ms_host) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("plan: timed forward failed: %s", cudaGetErrorString(e)); return 1; }
    for (size_t i = 0; i < ops.size(); ++i) op_ms_host[i] = 0.f;
    for (size_t li = 0; li < launches.size(); ++li) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[li], ev[li + 1]);
      op_ms_host[launches[li].op] += ms;   // a sub-batched op reports the sum over its sub-batches
    }
  }
  return 0;
}

int Plan::op_info(int i, stl_op_info* info) const {
  if (i < 0 || i >= (int)ops.size() || !info) { set_error("op_info: bad arguments"); return 1; }
  memset(info, 0, sizeof(*info));
  const Op& op = ops[i];
  info->kind = (int)op.kind;
  info->layer = op.layer;
  const Slot* so = op.out >= 0 ? &slots[op.out] : nullptr;
  const Slot* si = op.in >= 0 ? &slots[op.in] : nullptr;
  if (op.kind == OP_STEM) {
    info->out_h = cfg.image_h; info->out_w = cfg.image_w; info->cin = 3; info->cout = so->C;
    info->bytes_per_image = (double)cfg.image_h * cfg.image_w * 3 * 4 + (double)so->H * so->W * so->C * 2;
    return 0;
  }
  if (op.kind == OP_FUSE) {
    info->out_h = so->H; info->out_w = so->W; info->cout = so->C; info->cin = so->C;
    info->bytes_per_image = 2.0 * so->H * so->W * so->C * 2;
    for (int u = 0; u < op.n_up; ++u) info->bytes_per_image += (double)slots[op.up[u]].H * slots[op.up[u]].W * so->C * 2;
    if (op.layer >= 0) {   // + heatmap head: the fused map is not written, fp32 heatmaps are
      const Layer& Lh = layers[op.layer];
      info->cout = Lh.cout; info->ksize = 1; info->stride = 1;
      info->flops_per_image = 2.0 * Lh.cout * Lh.cin * so->H * so->W;
      info->bytes_per_image += (double)so->H * so->W * (Lh.cout * 4.0 - so->C * 2.0);
    }
    return 0;
  }
  if (op.kind == OP_LINK) {    // two 1x1 convs; algorithmic traffic: read t and the residual, write out and a
    const Layer& L3 = layers[op.layer];
    const Layer& L1 = layers[op.layer2];
    info->out_h = so->H; info->out_w = so->W; info->cin = L3.cin; info->cout = L3.cout; info->ksize = 1; info->stride = 1;
    const int cin1 = L3.cin + (op.cat >= 0 ? layers[cats[op.cat].lb].cin : 0);
    info->cin = cin1;
    info->flops_per_image = 2.0 * ((double)L3.cout * cin1 + (double)L1.cout * L1.cin) * so->H * so->W;
    info->bytes_per_image = (double)so->H * so->W * 2 * (cin1 + (op.cat >= 0 ? 1.0 : 2.0) * L3.cout + L1.cout);
    info->mb = 1; info->nt = 128; info->ck = 64; info->grid = 148;
    return 0;
  }
  if (op.kind == OP_BLOCK) {   // two 3x3 convs; algorithmic traffic: read x, write y
    const Layer& L1 = layers[op.layer];
    info->out_h = so->H; info->out_w = so->W; info->cin = L1.cin; info->cout = L1.cout; info->ksize = 3; info->stride = 1;
    info->flops_per_image = 2.0 * (2.0 * L1.cout * L1.cin * 9 * so->H * so->W);
    info->bytes_per_image = 2.0 * so->H * so->W * so->C * 2;
    info->mb = 3; info->nt = L1.cout; info->ck = L1.cin; info->grid = 148;
    return 0;
  }
  const Layer& L = layers[op.layer];
  const int ih = si ? si->H : cfg.image_h, iw = si ? si->W : cfg.image_w;
  info->out_h = L.im2col ? ih : ih / L.stride; info->out_w = L.im2col ? iw : iw / L.stride;
  const int cin = L.cin + (op.cat >= 0 ? layers[cats[op.cat].lb].cin : 0);   // both inputs of a concatenated pair
  info->cin = cin; info->cout = L.cout; info->ksize = L.k; info->stride = L.stride;
  info->flops_per_image = 2.0 * L.cout * cin * L.k * L.k * info->out_h * info->out_w;
  const double in_b = (double)ih * iw * (L.im2col ? L.cin_pad : cin) * (si ? 2 : 4);
  const double out_b = (double)info->out_h * info->out_w * L.cout * (op.out_nchw ? 4 : 2);
  info->bytes_per_image = in_b + out_b + (op.res >= 0 ? out_b : 0);
  for (int u = 0; u < op.n_up; ++u) info->bytes_per_image += (double)slots[op.up[u]].H * slots[op.up[u]].W * L.cout * 2;
  if (bound && op.kind == OP_CONV) {
    const Prepared& pr = prepared[i][0];
    info->grid = pr.grid; info->smem = (int)pr.smem; info->mb = pr.params.mb; info->nt = pr.params.nt;
    info->ck = pr.params.ck; info->a_stages = pr.params.a_stages; info->b_stages = pr.params.b_stages;
    info->a_shift = pr.params.a_shift; info->tiles = (int)pr.params.total_tiles;
    info->subs = (int)prepared[i].size();
  }
  return 0;
}

}  // namespace stl

// Plan: static schedule for one HRNet configuration (see plan.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/stlpose_b200.h"
#include "conv.h"

namespace stl {

struct Plan {
  struct Layer {
    std::string conv_key, bn_key;
    int cout, cin, k, stride, cout_pad, cin_pad;
    bool im2col = false;  // stem conv1: executed as a 1x1 conv over the im2col-packed network input (K = cin*k*k -> 32)
    size_t w_off, b_off;
  };
  struct Slot { int C, H, W; };
  enum OpKind { OP_STEM, OP_CONV, OP_FUSE, OP_BLOCK, OP_LINK };   // OP_BLOCK: one BasicBlock (two convs) in one kernel;
                                                                  // OP_LINK: conv3 of a Bottleneck + conv1 of the next
  struct Op {
    OpKind kind = OP_CONV;
    int layer = -1, in = -1, out = -1, res = -1;
    int layer2 = -1;  // OP_BLOCK: the block's second convolution; OP_LINK: conv1 of the next Bottleneck
    int out2 = -1;    // OP_LINK: that convolution's output
    int in2 = -1;     // OP_CONV with cat >= 0: the second input (K = [in | in2])
    int cat = -1;     // index into Plan::cats: this op runs two summed 1x1 convolutions as one (concatenated weights)
    int n_up = 0;
    int up[kMaxUp] = {-1, -1, -1};
    int up_shift[kMaxUp] = {0, 0, 0};
    bool relu = false, out_nchw = false;
    int group = -1;  // ops of one group run sub-batch by sub-batch (L2-resident working set), see Plan::bind
    int par_group = -1;  // >= 0: branch op of HighResolutionModule #par_group; the branches of a module are independent
    int stream = 0;      // ... and run concurrently, branch b on side stream b (0 = the caller's stream)
    int sm_share = 0;    // ... each restricted to its share of the SMs (CTAs of its persistent kernels)
  };
  struct Group { int first = 0, count = 0; };
  // Two 1x1 convolutions whose outputs are added before the activation (Bottleneck conv3 + downsample of layer1.0,
  // HRnet.py:88-101) as ONE convolution over the concatenated inputs: weights [cout_pad][cin_a + cin_b] and the summed
  // bias live in their own arena region, rebuilt by pack_conv whenever either member is packed.
  struct Cat { int la = -1, lb = -1; size_t w_off = 0, b_off = 0; };
  struct Launch { int op, sub; };
  struct Prepared {
    ConvParams params;
    int grid = 0;      // whole machine
    int grid_par = 0;  // when the op runs concurrently with the other branches of its module
    size_t smem = 0;
  };
  struct Builder;

  stl_hrnet_cfg cfg{};
  std::vector<Layer> layers;
  std::vector<Slot> slots;
  std::vector<Op> ops;
  std::vector<Cat> cats;
  int fuse_head = 1;        // STLPOSE_FUSE_HEAD=0: the last fuse row and the heatmap head as two launches
  int use_pdl = 1;          // STLPOSE_PDL=0: no programmatic dependent launch between consecutive kernels
  int fuse_links = 1;       // STLPOSE_FUSE_LINK=0: conv3 of a layer1 Bottleneck and conv1 of the next as two launches
  int fuse_downsample = 1;  // STLPOSE_FUSE_DOWNSAMPLE=0: downsample and conv3 of layer1.0 as two launches
  size_t weight_bytes = 0;
  int tap_reload = 0;  // debugging: force one TMA load per filter tap
  int stem_im2col = 1; // STLPOSE_STEM_IM2COL=0: 16-channel input packing + stride-2 3x3 tensor-core conv instead
  int n_par_groups = 0;
  int branch_streams = 0;  // STLPOSE_BRANCH_STREAMS=1: the branches of a module on concurrent side streams, each on a
                           // share of the SMs.  Measured on B200: 34.8 ms per step against 29.9 ms sequential (the
                           // restricted kernels are 20-40 % more efficient per SM in isolation, but they do not
                           // overlap well enough to pay for running each on a fraction of the machine) - off.
  cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
  int fuse_blocks = 1; // BasicBlocks of 32-channel branches run as one fused kernel (STLPOSE_FUSE_BLOCK=0: two convs)

  // binding state
  bool bound = false;
  int bound_images = 0;
  const void* bound_arena = nullptr;
  void* bound_ws = nullptr;
  std::vector<uint8_t*> slot_ptr;
  std::vector<std::vector<Prepared>> prepared;  // [op][sub-batch]
  std::vector<Group> groups;
  std::vector<Launch> launches;                 // execution order for the current binding

  static Plan* create(const stl_hrnet_cfg& cfg);
  size_t workspace_bytes(int n_images) const;
  int pack_conv(int index, const float* w, const float* gamma, const float* beta, const float* mean,
                const float* var, const float* cbias, float eps, void* arena, cudaStream_t st);
  int bind(int n_images, const void* arena, void* workspace, size_t ws_bytes, cudaStream_t st);
  int forward(const float* x, int B, int flip_pair, float* heat, const void* arena, void* workspace, size_t ws_bytes,
              cudaStream_t st, float* op_ms_host = nullptr);
  int op_info(int i, stl_op_info* info) const;
};

}  // namespace stl

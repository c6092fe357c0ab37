"""Fixed-shape inference pipeline: forward (+ flip-test pass) + fused flip-average/decode, replayed as a CUDA graph.

This is the evaluation hot loop of the reference (03_evaluate.py:124-152: forward_pass(flip=True) followed by
get_final_preds_hrnet) for a fixed batch size, with every launch captured once and replayed.
"""
import ctypes

import torch

from . import _lib
from .hrnet import PoseHighResolutionNet
from .transforms import FLIP_PAIRS, _pairs_array


class KeypointPipeline:
    def __init__(self, model, batch, image_size=(256, 192), flip=True, use_graph=True, pairs=FLIP_PAIRS):
        if not isinstance(model, PoseHighResolutionNet):
            raise TypeError("KeypointPipeline expects stlpose_b200.PoseHighResolutionNet")
        self.model, self.B, self.flip = model, int(batch), bool(flip)
        dev = model.conv1.weight.device
        if dev.type != "cuda":
            raise _lib.StlError("model must live on a CUDA device")
        self.device = dev
        H, W = image_size
        J = model.num_joints
        self.x = torch.zeros((self.B, 3, H, W), dtype=torch.float32, device=dev)
        self.center = torch.zeros((self.B, 2), dtype=torch.float32, device=dev)
        self.scale = torch.ones((self.B, 2), dtype=torch.float32, device=dev)
        self.preds = torch.zeros((self.B, J, 2), dtype=torch.float32, device=dev)
        self.maxvals = torch.zeros((self.B, J, 1), dtype=torch.float32, device=dev)
        self.coords = torch.zeros((self.B, J, 2), dtype=torch.float32, device=dev)
        self.h, self.w = H // 4, W // 4
        self._pairs, self._n_pairs = _pairs_array(pairs)
        # host->device staging: two device buffers filled on a side stream, so the H2D copy of batch k+1 overlaps the
        # network pass of batch k (a 512-crop fp32 batch is 302 MB, ~5 ms over PCIe)
        self._stage = [torch.empty_like(self.x) for _ in range(2)]
        self._stage_c = [torch.zeros_like(self.center) for _ in range(2)]
        self._stage_s = [torch.ones_like(self.scale) for _ in range(2)]
        self._stage_ready = [torch.cuda.Event() for _ in range(2)]
        self._stage_free = [torch.cuda.Event() for _ in range(2)]
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._calls = 0
        self.graph = None
        self._stage_graphs = None                    # one graph per staging buffer: the end-to-end call skips the copy into x
        self._use_graph = bool(use_graph)
        self._capture()

    def _capture(self):
        """(Re)build the captured steps.  The graphs hold raw addresses of the model's weight arena and activation
        workspace; the model bumps ``_buffer_generation`` whenever it reallocates either (a larger eager batch at the
        same resolution, ``.to()`` / ``.cuda()``), and ``_ensure_current`` re-captures before the next replay."""
        model, dev, use_graph = self.model, self.device, self._use_graph
        H, W = self.x.shape[2], self.x.shape[3]
        self.graph, self._stage_graphs = None, None
        with torch.cuda.device(dev):
            self._step_eager()                       # packs weights, binds the workspace, warms everything up
            torch.cuda.synchronize()
            self.launches_per_step = model.launches_per_forward(H, W) + 1   # + the decode kernel
            if use_graph:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._step_eager()
                torch.cuda.current_stream().wait_stream(s)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step_eager()
                self.graph = g
                # the same step reading its crops straight from staging buffer j (same workspace, same outputs)
                self._stage_graphs = []
                for j in range(2):
                    gj = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gj):
                        self._step_eager(self._stage[j])
                    self._stage_graphs.append(gj)
        self._generation = model._buffer_generation

    def _ensure_current(self):
        if self.model._buffer_generation != self._generation:
            if self.model.conv1.weight.device != self.device:
                raise _lib.StlError("the model was moved to another device after the pipeline was built")
            torch.cuda.synchronize(self.device)
            self._capture()

    def _step_eager(self, x=None):
        L = _lib.lib()
        heat = self.model._run(self.x if x is None else x, flip_pair=self.flip)
        self._heat = heat                            # keep the graph's output buffer alive
        hf = heat[self.B:] if self.flip else None
        _lib.check(L.stl_decode(_lib.ptr(heat), _lib.ptr(hf), _lib.ptr(self.center), _lib.ptr(self.scale), self.B,
                                self.model.num_joints, self.h, self.w, self._pairs, self._n_pairs, 1, None,
                                _lib.ptr(self.preds), _lib.ptr(self.maxvals), _lib.ptr(self.coords),
                                _lib.current_stream()))

    def step(self):
        """Run one batch from the static device buffers (x, center, scale) into (preds, maxvals, coords)."""
        self._ensure_current()
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_eager()

    def __call__(self, x_host, center_host, scale_host, preds_host=None, maxvals_host=None):
        """End-to-end call with HOST (ideally pinned) buffers: H2D copies, step, D2H of preds and maxvals.

        Everything is enqueued asynchronously: the crops (and, ahead of them, the boxes) go host -> staging buffer on a
        copy stream (double buffered, so the copy of the next call overlaps this call's network pass); the step is
        replayed from a graph that reads the staging buffer directly (``self.x`` is only the input of ``step()``).
        Synchronise the current stream before reading the returned host tensors."""
        self._ensure_current()
        cur = torch.cuda.current_stream(self.device)
        j = self._calls & 1
        self._calls += 1
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._stage_free[j])      # the pass that last read this buffer is done
            # the boxes travel with the crops, AHEAD of them on the copy stream: issued on the compute stream they would
            # queue on the host->device copy engine behind the 302 MB crop copy of the next call (measured: 2 ms per step)
            self._stage_c[j].copy_(center_host, non_blocking=True)
            self._stage_s[j].copy_(scale_host, non_blocking=True)
            self._stage[j].copy_(x_host, non_blocking=True)
            self._stage_ready[j].record(self._copy_stream)
        cur.wait_event(self._stage_ready[j])
        self.center.copy_(self._stage_c[j], non_blocking=True)
        self.scale.copy_(self._stage_s[j], non_blocking=True)
        if self._stage_graphs is not None:
            self._stage_graphs[j].replay()                          # reads the crops from the staging buffer itself
        else:
            self._step_eager(self._stage[j])
        self._stage_free[j].record(cur)
        if preds_host is None:
            preds_host = torch.empty(tuple(self.preds.shape), dtype=torch.float32).pin_memory()
            maxvals_host = torch.empty(tuple(self.maxvals.shape), dtype=torch.float32).pin_memory()
        preds_host.copy_(self.preds, non_blocking=True)
        maxvals_host.copy_(self.maxvals, non_blocking=True)
        return preds_host, maxvals_host

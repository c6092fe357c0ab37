"""One fine-tuning iteration of the reference (/root/reference/src/02_train.py:203-218) as a single CUDA graph.

    output = forward_pass(model, imgs, "HRNet")      # train-mode BatchNorm
    loss   = PersonMSELoss()(output, target, target_weight)
    optimizer.zero_grad(); loss.backward(); optimizer.step()

A step is ~3 000 kernel launches (293 conv + BatchNorm units, forward and backward); issued eagerly the host is the
bottleneck below ~128 crops per step.  The step is therefore captured once (static input buffers, allocations from the
graph's private pool) and replayed; inputs are copied into the static buffers, the loss is a device scalar.

Data parallel (``reducer=parallel.GradientReducer``, one process per GPU): the reducer is bound to the model, so the
weight-gradient and BatchNorm-backward kernels write straight into flat gradient buckets and every bucket is all-reduced
on a communication stream as soon as backward has filled it - the ``ncclAllReduce`` launches are nodes of the same
captured graph, forked next to the remaining backward kernels and joined before the optimizer step.

The optimizer step is part of the graph when that is safe (``capture_optimizer``): captured kernels bake in the Python
floats of ``param_groups`` (lr, momentum, weight decay ...), so the step keeps a signature of those values and
RE-CAPTURES the graph when a scheduler (lib/model_setup.py: StepLR / ReduceLROnPlateau) or the caller changes them.
Optimizers that cannot be captured (Adam without ``capturable=True``, SGD with dampening, whose first step differs from
the following ones) run eagerly after the replay instead.

Construct the TrainStep before running any eager backward pass of the same parameters on the default stream: autograd
pins each parameter's gradient-accumulation node to the stream of its first backward, and a node pinned to the
default stream makes the capture depend on a non-capturing stream (cudaErrorStreamCaptureInvalidated).  Eager steps
under ``torch.cuda.stream(side_stream)`` are fine.
"""
import torch

import os

from . import optim as _optim
from .inference import forward_pass

FUSED_SGD = os.environ.get("STLPOSE_FUSED_SGD", "1") != "0"

_HYPER_SKIP = ("params",)


def _hyper_signature(optimizer):
    """Everything in param_groups a captured optimizer step would bake into its kernels."""
    sig = []
    for g in optimizer.param_groups:
        items = []
        for k in sorted(g):
            if k in _HYPER_SKIP:
                continue
            v = g[k]
            if torch.is_tensor(v):
                v = ("tensor", v.data_ptr())          # a tensor hyper-parameter is read on the device every replay
            elif isinstance(v, (list, tuple)):
                v = tuple(v)
            items.append((k, v))
        sig.append(tuple(items))
    return tuple(sig)


def _capturable(optimizer):
    for g in optimizer.param_groups:
        if "capturable" in g and not g["capturable"]:
            return False                              # e.g. torch.optim.Adam(capturable=False): host-side step counter
        if g.get("dampening", 0) != 0:
            return False                              # first SGD step (buf = grad) differs from the others
        if g.get("fused", False) or g.get("differentiable", False):
            return False
    return True


def graph_kernel_nodes(graph):
    """Kernel nodes of a captured step (a CUDAGraph created with keep_graph=True), or None when the graph cannot be
    inspected.  Measurement aid for bench.py's ``gpu_launches``."""
    try:
        from cuda.bindings import runtime as cudart
        g = cudart.cudaGraph_t(int(graph.raw_cuda_graph()))
        err, _, n = cudart.cudaGraphGetNodes(g, 0)
        if int(err) != 0 or n == 0:
            return None
        err, nodes, n = cudart.cudaGraphGetNodes(g, n)
        kernels = 0
        for node in nodes[:n]:
            err, t = cudart.cudaGraphNodeGetType(node)
            kernels += int(err) == 0 and t == cudart.cudaGraphNodeType.cudaGraphNodeTypeKernel
        return kernels
    except Exception:
        return None


class TrainStep:
    def __init__(self, model, optimizer, criterion, batch, image_size=(256, 192), joints=17, use_graph=True, warmup=3,
                 reducer=None, capture_optimizer=True, count_kernels=False):
        """reducer: a ``parallel.GradientReducer`` over ALL parameters of ``model`` for data-parallel runs (one process
        per GPU); it is bound to the model here (gradients are written into its buckets and reduced during backward).
        capture_optimizer: put ``optimizer.step()`` inside the graph when the optimizer allows it (see the module
        docstring); False always steps eagerly after the replay."""
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("TrainStep needs the model on a CUDA device (there is no CPU fallback)")
        H, W = image_size
        self.model, self.optimizer, self.criterion, self.reducer = model, optimizer, criterion, reducer
        self.x = torch.zeros((batch, 3, H, W), dtype=torch.float32, device=dev)
        self.target = torch.zeros((batch, joints, H // 4, W // 4), dtype=torch.float32, device=dev)
        self.target_weight = torch.ones((batch, joints, 1), dtype=torch.float32, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.output = None
        self.graph = None
        self.captures = 0
        self.kernel_nodes = None          # kernels of one captured step (count_kernels=True)
        self._count_kernels = bool(count_kernels)
        self.use_graph = bool(use_graph)
        self.warmup = warmup
        self.optimizer_in_graph = bool(use_graph and capture_optimizer and _capturable(optimizer))
        # plain torch.optim.SGD: the whole update as one launch (stlpose_b200.optim) instead of ~47 foreach launches;
        # same arithmetic, the optimizer's own state, so schedulers / state_dict() are unaffected
        self.fused_sgd = bool(FUSED_SGD and _optim.supports(optimizer))
        self._sgd_tables = None
        self._hyper = None
        model.train()
        if reducer is not None:
            reducer.bind(model)
        if use_graph:
            self._capture(first=True)

    # ------------------------------------------------------------------------------------------------- capture
    def _capture(self, first):
        dev = self.x.device
        snapshot = self._snapshot()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        if self.fused_sgd:
            self._sgd_tables = _optim.make_tables(self.optimizer)      # warm-up tables (eager gradient addresses)
        with torch.cuda.stream(side):                      # warm-up: lazy optimizer state, one-time attributes, NCCL setup
            for _ in range(self.warmup if first else 1):
                self._eager(update=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._restore(snapshot)                            # warm-up steps must not count as training
        if self.reducer is None:
            self.optimizer.zero_grad(set_to_none=True)
        if self.fused_sgd:
            # the captured step uploads its pointer table from a pinned buffer on every replay: a fresh set, allocated
            # before the capture begins and never refilled afterwards
            self._sgd_tables = _optim.make_tables(self.optimizer)
        self.graph = torch.cuda.CUDAGraph(keep_graph=True) if self._count_kernels else torch.cuda.CUDAGraph()
        # thread_local: the NCCL watchdog thread polls its events while this thread captures
        mode = "thread_local" if (self.reducer is not None and self.reducer.world > 1) else "global"
        with torch.cuda.graph(self.graph, capture_error_mode=mode):
            self._eager(update=self.optimizer_in_graph)
        if self._count_kernels:
            self.kernel_nodes = graph_kernel_nodes(self.graph)
            self.graph.instantiate()
        self._restore(snapshot)                            # capture does not execute, but keep the contract explicit
        self._hyper = _hyper_signature(self.optimizer)
        self.captures += 1

    def _snapshot(self):
        state = {}
        for p, st in self.optimizer.state.items():
            state[p] = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
        return ({k: v.detach().clone() for k, v in self.model.state_dict().items()}, state)

    def _restore(self, snap):
        """Model parameters / buffers and the optimizer state back to the snapshot, IN PLACE (the captured graph holds
        the addresses).  State the warm-up created lazily (momentum buffers, Adam moments and step counters) is zeroed -
        for the optimizers that may be captured that is exactly the state before their first step - or, when the
        optimizer steps eagerly, removed so that its first real step is torch's own first step."""
        with torch.no_grad():
            sd = self.model.state_dict()
            for k, v in snap[0].items():
                sd[k].copy_(v)
            for p in list(self.optimizer.state.keys()):
                st, old = self.optimizer.state[p], snap[1].get(p)
                if old is None and not self.optimizer_in_graph:
                    del self.optimizer.state[p]
                    continue
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if old is not None and k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()
                    elif old is not None and k in old:
                        st[k] = old[k]

    # ------------------------------------------------------------------------------------------------- one step
    def _eager(self, update):
        out = forward_pass(self.model, self.x, "HRNet", device=self.x.device, flip=False)
        loss = self.criterion(out, self.target, self.target_weight)
        if self.reducer is None:
            self.optimizer.zero_grad()
            loss.backward()
        else:
            # bound reducer: the kernels overwrite the bucket views, nothing accumulates, nothing to zero; the loss
            # gradient carries this rank's share B_r / B of the global mean (lib/loss.py:87)
            loss.backward(gradient=self.reducer.scale_tensor)
            self.reducer.finish_backward()
        if update:
            self._optimizer_step()
        self.output = out.detach()
        self.loss.copy_(loss.detach())

    def _optimizer_step(self):
        if self.fused_sgd:
            if self._sgd_tables is None or (self.graph is not None and self.optimizer_in_graph is False):
                # eager steps (no graph, or an optimizer that steps after the replay): their own tables
                if getattr(self, "_sgd_tables_eager", None) is None:
                    self._sgd_tables_eager = _optim.make_tables(self.optimizer)
                _optim.fused_sgd_step(self.optimizer, self._sgd_tables_eager)
            else:
                _optim.fused_sgd_step(self.optimizer, self._sgd_tables)
        else:
            self.optimizer.step()

    def __call__(self, imgs, target, target_weight):
        """Copies the batch into the static buffers (host or device sources), runs the step, returns the device loss."""
        self.x.copy_(imgs, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        self.target_weight.copy_(target_weight.reshape(self.target_weight.shape), non_blocking=True)
        if self.graph is not None:
            if self.optimizer_in_graph and _hyper_signature(self.optimizer) != self._hyper:
                self._recapture()                          # a scheduler changed lr / momentum / ...: bake the new values
            self.graph.replay()
            if not self.optimizer_in_graph:
                self._optimizer_step()
        else:
            self._eager(update=True)
        self.model.invalidate_packed_weights()
        return self.loss

    def _recapture(self):
        """New graph with the optimizer's current hyper-parameters.  Training state (parameters, running statistics,
        optimizer state) is preserved: _capture snapshots and restores it around its warm-up step."""
        self.graph = None
        self._capture(first=False)

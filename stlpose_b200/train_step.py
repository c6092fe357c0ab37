"""One fine-tuning iteration of the reference (/root/reference/src/02_train.py:203-218) as a single CUDA graph.

    output = forward_pass(model, imgs, "HRNet")      # train-mode BatchNorm
    loss   = PersonMSELoss()(output, target, target_weight)
    optimizer.zero_grad(); loss.backward(); optimizer.step()

A step is ~3 400 kernel launches (293 conv + BatchNorm units, forward and backward); issued eagerly the host is the
bottleneck below ~128 crops per step.  The step is therefore captured once (static input buffers, allocations from the
graph's private pool) and replayed; inputs are copied into the static buffers, the loss is a device scalar.

Construct the TrainStep before running any eager backward pass of the same parameters on the default stream: autograd
pins each parameter's gradient-accumulation node to the stream of its first backward, and a node pinned to the
default stream makes the capture depend on a non-capturing stream (cudaErrorStreamCaptureInvalidated).  Eager steps
under ``torch.cuda.stream(side_stream)`` are fine.
"""
import torch

from .inference import forward_pass


class TrainStep:
    def __init__(self, model, optimizer, criterion, batch, image_size=(256, 192), joints=17, use_graph=True, warmup=3,
                 reducer=None):
        """reducer: a ``parallel.GradientReducer`` for data-parallel runs (one process per GPU).  The graph then holds
        forward + loss + backward of this rank's slice; the gradient all-reduce and the optimizer step follow it."""
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
        model.train()
        if use_graph:
            snapshot = self._snapshot()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                      # warm-up: lazy optimizer state, one-time attributes
                for _ in range(warmup):
                    self._eager()
                    self._exchange_and_update()
            torch.cuda.current_stream(dev).wait_stream(side)
            self._restore(snapshot)                            # warm-up steps must not count as training
            self.optimizer.zero_grad(set_to_none=True)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._eager()
            self._restore(snapshot)                            # capture does not execute, but keep the contract explicit

    def _snapshot(self):
        return ({k: v.detach().clone() for k, v in self.model.state_dict().items()},)

    def _restore(self, snap):
        with torch.no_grad():
            sd = self.model.state_dict()
            for k, v in snap[0].items():
                sd[k].copy_(v)
            for group in self.optimizer.param_groups:          # momentum / Adam moments back to zero, in place
                for p in group["params"]:
                    for v in self.optimizer.state.get(p, {}).values():
                        if torch.is_tensor(v):
                            v.zero_()

    def _eager(self):
        out = forward_pass(self.model, self.x, "HRNet", device=self.x.device, flip=False)
        loss = self.criterion(out, self.target, self.target_weight)
        self.optimizer.zero_grad()
        loss.backward()
        if self.reducer is None:
            self.optimizer.step()
        self.output = out.detach()
        self.loss.copy_(loss.detach())

    def _exchange_and_update(self):
        if self.reducer is not None:
            self.reducer.reduce_all()
            self.optimizer.step()

    def __call__(self, imgs, target, target_weight):
        """Copies the batch into the static buffers (host or device sources), runs the step, returns the device loss."""
        self.x.copy_(imgs, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        self.target_weight.copy_(target_weight.reshape(self.target_weight.shape), non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._eager()
        self._exchange_and_update()
        self.model.invalidate_packed_weights()
        return self.loss

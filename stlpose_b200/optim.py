"""``optimizer.step()`` of the fine-tuning loop (/root/reference/src/02_train.py:218) as ONE kernel launch.

The reference trains with ``torch.optim.SGD(net.parameters(), lr, momentum=0.9, weight_decay=...)``
(lib/model_setup.py:138-139).  torch's foreach implementation walks the 878 parameter tensors of HRNet-W32 in ~47
multi-tensor launches (1.0 ms per step on a B200, 5 % of a 32-crop step); ``fused_sgd_step(optimizer)`` performs the same
update - same operations, same order, the optimizer's own ``param_groups`` and ``state`` (so ``state_dict()`` and
schedulers keep working) - with ``stl_sgd_step_batched``.  ``TrainStep`` uses it automatically for plain SGD.
"""
import numpy as np
import torch

from . import _lib

_ITEM = np.dtype([("p", "<u8"), ("g", "<u8"), ("buf", "<u8"), ("numel", "<i4"), ("pad", "<i4")])


def supports(optimizer):
    """Plain torch.optim.SGD whose update the kernel reproduces exactly (dampening 0, no maximize, fp32 CUDA params)."""
    if type(optimizer) is not torch.optim.SGD:
        return False
    for g in optimizer.param_groups:
        if g.get("dampening", 0) != 0 or g.get("maximize", False) or g.get("differentiable", False):
            return False
        if torch.is_tensor(g["lr"]):
            return False
        for p in g["params"]:
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                return False
    return True


class _Table:
    """Device table of (param, grad, momentum buffer) pointers of one param group, refilled when any pointer changes.
    The pinned host copy and the device buffer are allocated HERE (outside any CUDA-graph capture); inside a capture the
    upload is a captured copy node that is replayed from this very host buffer, so a table that was captured must not
    be refilled afterwards (TrainStep keeps separate tables for its warm-up and for its graph)."""

    def __init__(self, n_max, device):
        nbytes = n_max * _ITEM.itemsize + (n_max + 1) * 4
        self.host = torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
        self.dev = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.n_max, self.key, self.n, self.blocks = n_max, None, 0, 0
        self.items = self.offsets = None

    def update(self, params, grads, bufs):
        key = tuple((p.data_ptr(), g.data_ptr(), 0 if b is None else b.data_ptr()) for p, g, b in zip(params, grads, bufs))
        if key == self.key:
            return
        n = len(params)
        assert n <= self.n_max
        items = np.zeros(n, dtype=_ITEM)
        offsets = np.zeros(n + 1, dtype=np.int32)
        for i, (p, (pp, gp, bp)) in enumerate(zip(params, key)):
            items[i] = (pp, gp, bp, p.numel(), 0)
            offsets[i + 1] = offsets[i] + (p.numel() + 1023) // 1024
        raw = np.concatenate([items.view(np.uint8), offsets.view(np.uint8)])
        self.host.numpy()[: raw.size] = raw
        self.dev.copy_(self.host, non_blocking=True)
        self.items, self.offsets = self.dev[: n * _ITEM.itemsize], self.dev[n * _ITEM.itemsize: raw.size]
        self.key, self.n, self.blocks = key, n, int(offsets[-1])


def make_tables(optimizer):
    """Fresh pointer tables for every param group (allocate them before a CUDA-graph capture begins)."""
    return {gi: _Table(len(g["params"]), g["params"][0].device) for gi, g in enumerate(optimizer.param_groups) if g["params"]}


@torch.no_grad()
def fused_sgd_step(optimizer, tables=None):
    """One ``torch.optim.SGD.step()`` (see ``supports``).  ``tables``: ``make_tables(optimizer)``, kept by the caller
    between steps (pointer tables per param group); without it they are rebuilt on every call."""
    L = _lib.lib()
    tables = make_tables(optimizer) if tables is None else tables
    for gi, group in enumerate(optimizer.param_groups):
        momentum = float(group["momentum"])
        params, grads, bufs = [], [], []
        for p in group["params"]:
            if p.grad is None:
                continue
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous():
                raise _lib.StlError("fused_sgd_step needs contiguous fp32 gradients")
            buf = None
            if momentum != 0:
                st = optimizer.state[p]
                buf = st.get("momentum_buffer")
                if buf is None:                       # torch: buf = clone(grad) on the first step == 0 * momentum + grad
                    buf = st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            params.append(p); grads.append(g); bufs.append(buf)
        if not params:
            continue
        tab = tables[gi] if gi in tables else tables.setdefault(gi, _Table(len(group["params"]), params[0].device))
        tab.update(params, grads, bufs)
        with torch.cuda.device(params[0].device):
            _lib.check(L.stl_sgd_step_batched(_lib.ptr(tab.items), _lib.ptr(tab.offsets), tab.n, tab.blocks,
                                              float(group["lr"]), momentum, float(group["weight_decay"]),
                                              int(bool(group["nesterov"])), _lib.current_stream()))

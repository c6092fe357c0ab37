"""Batch sharding across the GPUs of one box: one process per GPU, no collective on the data path.

Replaces the reference's single-process ``torch.nn.DataParallel`` wrapper (02_train.py:109, 03_evaluate.py:100:
scatter -> replicate -> parallel_apply -> gather every forward).  Person crops are independent in inference, so
each rank runs the whole pipeline on a contiguous slice of the batch with its own resident copy of the weights;
the only exchange is the final all-gather of the keypoints (204 B per crop).
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world_size, rank):
    """Contiguous slice [lo, hi) of a batch of n for `rank`, split like DataParallel.scatter / torch.chunk:
    chunks of ceil(n / world_size); trailing ranks may get a short or empty slice."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    chunk = -(-n // world_size) if n > 0 else 0
    lo = min(n, rank * chunk)
    hi = min(n, lo + chunk)
    return lo, hi


def gather_keypoints(preds_local, maxvals_local, n_total, group=None):
    """All-gather per-rank results ([n_r,J,2], [n_r,J,1]) back into batch order -> ([n_total,J,2], [n_total,J,1]).

    Works on whatever backend the process group uses (NCCL for CUDA tensors, gloo for CPU tensors)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return preds_local, maxvals_local
    J = preds_local.shape[1]
    chunk = -(-n_total // world) if n_total > 0 else 0
    packed = torch.zeros((chunk, J, 3), dtype=preds_local.dtype, device=preds_local.device)
    n_r = preds_local.shape[0]
    packed[:n_r, :, :2] = preds_local
    packed[:n_r, :, 2:] = maxvals_local
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        parts.append(out[r][: hi - lo])
    full = torch.cat(parts, dim=0)
    return full[..., :2].contiguous(), full[..., 2:].contiguous()


class ShardedKeypointInference:
    """forward_pass(flip=True) + get_final_preds_hrnet over a batch sharded across the ranks of the process group.

    Every rank calls ``run`` with the full batch (host or device tensors); each computes only its slice and all
    ranks return the full, ordered result.  The model must already live on this rank's GPU.
    """

    def __init__(self, model, flip=True, group=None, use_graph=True):
        self.model, self.flip, self.group, self.use_graph = model, flip, group, use_graph
        self._pipes = {}           # (local batch, H, W) -> KeypointPipeline: the slice runs as one captured graph replay

    def _pipeline(self, n_local, h, w):
        from .pipeline import KeypointPipeline
        key = (n_local, h, w)
        if key not in self._pipes:
            if len(self._pipes) >= 4:                      # a few batch sizes (full batches + the last, short one)
                self._pipes.pop(next(iter(self._pipes)))
            self._pipes[key] = KeypointPipeline(self.model, n_local, (h, w), flip=self.flip, use_graph=self.use_graph)
        return self._pipes[key]

    def run(self, imgs, center, scale):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        n = imgs.shape[0]
        lo, hi = shard_bounds(n, world, rank)
        dev = self.model.conv1.weight.device
        J = self.model.num_joints
        if hi > lo:
            # forward (+ mirrored forward) + flip-average + decode of this rank's slice: one graph replay
            # (stlpose_b200.pipeline), not ~270 eager launches
            pipe = self._pipeline(hi - lo, imgs.shape[2], imgs.shape[3])
            pipe.x.copy_(imgs[lo:hi], non_blocking=True)
            pipe.center.copy_(torch.as_tensor(center[lo:hi]), non_blocking=True)
            pipe.scale.copy_(torch.as_tensor(scale[lo:hi]), non_blocking=True)
            pipe.step()
            preds, maxvals = pipe.preds.clone(), pipe.maxvals.clone()
        else:
            preds = torch.zeros((0, J, 2), dtype=torch.float32, device=dev)
            maxvals = torch.zeros((0, J, 1), dtype=torch.float32, device=dev)
        return gather_keypoints(preds, maxvals, n, self.group)


class _Sink:
    """Where the training kernels write one parameter gradient (or a BatchNorm's dbeta | dgamma pair): a view into a flat
    bucket of a bound GradientReducer.  ``done()`` is called by the layer's backward once the kernels that fill the view
    are enqueued on the current stream."""
    __slots__ = ("reducer", "bucket", "view", "count")

    def __init__(self, reducer, bucket, view, count):
        self.reducer, self.bucket, self.view, self.count = reducer, bucket, view, count

    def done(self):
        self.reducer._written(self.bucket, self.count)


class GradientReducer:
    """The one exchange step of data-parallel fine-tuning (replaces DataParallel's replicate / gather / reduce of
    02_train.py:109,203-218): every rank runs forward + backward on its own slice of the batch with its own BatchNorm
    batch statistics (as DataParallel's replicas do), then the parameter gradients are summed over the ranks.

    The reference computes ONE loss over the gathered batch (loss.py:87 averages over all B crops), so a rank that
    averaged over its B_r local crops contributes its gradient scaled by B_r / B (``scale``).  Gradients live in flat
    fp32 buckets in reverse parameter order (the order backward produces them).  Two ways to use it:

    * **bound** (``bind(model)``, what ``TrainStep`` does): every ``p.grad`` IS a view into its bucket and the training
      kernels write there directly (training._ConvBN / _Head), so there is no pack or unpack copy; the layer that fills
      the last view of a bucket forks the communication stream, which all-reduces the bucket (``ncclAllReduce`` over
      NVLink) while the rest of backward keeps running - also inside a captured CUDA graph.  The caller passes
      ``loss.backward(gradient=reducer.scale_tensor)`` (the scale, when folding it into the loss gradient is exact) and
      calls ``finish_backward()`` after ``loss.backward()``: the current stream then waits for the last bucket.
    * **unbound** (any autograd graph, CPU/gloo included): ``attach_hooks()`` + ``finish_step()`` launch each bucket from
      post-accumulate-grad hooks, or ``reduce_all()`` after backward; gradients are packed into the buckets, scaled,
      all-reduced and copied back.

    BatchNorm running statistics are not exchanged: like the reference, which keeps the statistics of GPU 0's
    replica, every rank keeps its own and rank 0's are the ones to checkpoint.
    """

    def __init__(self, params, local_batch, group=None, bucket_bytes=32 << 20):
        self.group = group
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = self.params[0].device
        total = torch.tensor([float(local_batch)], device=dev)
        if self.world > 1:
            dist.all_reduce(total, group=group)
        self.scale = float(local_batch) / float(total.item())
        # Bound mode folds the scale into the loss gradient when that is exact - a power of two, i.e. equal slices on
        # 1/2/4/8 GPUs: scaling a bf16 activation gradient by 2^-k changes no mantissa bit - and otherwise multiplies the
        # fp32 buckets on the communication stream just before the all-reduce (uneven slices)
        import math
        self.fold_scale = math.frexp(self.scale)[0] == 0.5
        self.scale_tensor = torch.tensor(self.scale if self.fold_scale else 1.0, dtype=torch.float32, device=dev)
        self.global_batch = int(round(total.item()))
        # buckets in reverse registration order; a bucket only ends after a weight tensor (ndim > 1), so a BatchNorm's
        # (bias, weight) pair - adjacent in this order - is never split and can be written as one dbeta | dgamma block
        self.buckets, cur, cur_bytes = [], [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes and p.dim() > 1:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        # every parameter starts on a 16-byte boundary of its bucket (the kernels that write gradients there use float4
        # stores); a BatchNorm's C-element bias and weight stay adjacent because C is a multiple of 4
        self.offsets, sizes = [], []
        for b in self.buckets:
            offs, off = [], 0
            for p in b:
                offs.append(off)
                off += (p.numel() + 3) // 4 * 4
            self.offsets.append(offs)
            sizes.append(off)
        self.flat = [torch.zeros(n, dtype=torch.float32, device=dev) for n in sizes]
        self.bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._pending = [0] * len(self.buckets)
        self._works = []
        self._hooks = []
        self._events = {}          # id(param) -> event recorded on the stream that accumulated its gradient
        self._bound = None
        self._bucket_events = [[] for _ in self.buckets]
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None

    def views(self, i):
        """Per-parameter views (shaped like the parameters) of bucket i."""
        flat = self.flat[i]
        return [flat[o:o + p.numel()].view_as(p) for p, o in zip(self.buckets[i], self.offsets[i])]

    # ---- bound mode: p.grad is a view of its bucket, kernels write there, buckets are reduced as they fill up
    def bind(self, model):
        """Make every parameter's ``.grad`` a view into its bucket and tell the training path of ``model``
        (stlpose_b200.training) to write parameter gradients there.  Requires all parameters of ``model`` to be in
        this reducer (a frozen parameter would leave its layer half bound)."""
        if self.comm_stream is None:
            raise RuntimeError("GradientReducer.bind needs CUDA parameters")
        mine = {id(p) for p in self.params}
        missing = [n for n, p in model.named_parameters() if id(p) not in mine]
        if missing:
            raise ValueError(f"GradientReducer.bind: parameters outside the reducer, e.g. {missing[:3]}")
        offset = {}
        self._views = [self.views(i) for i in range(len(self.buckets))]
        for i, bucket in enumerate(self.buckets):
            for p, v, off in zip(bucket, self._views[i], self.offsets[i]):
                p.grad = v
                p._stl_sink = _Sink(self, i, v, 1)
                offset[id(p)] = (i, off)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d) and m.weight is not None:
                (ib, ob), (iw, ow) = offset[id(m.bias)], offset[id(m.weight)]
                c = m.bias.numel()
                if ib != iw or ow != ob + c:
                    raise ValueError("GradientReducer.bind: a BatchNorm's bias and weight are not adjacent in a bucket")
                m.bias._stl_sink = _Sink(self, ib, self.flat[ib][ob:ob + 2 * c], 2)       # dbeta | dgamma in one block
                m.weight._stl_sink = None
        self._bound = model
        self._pending = [len(b) for b in self.buckets]
        return self

    def unbind(self):
        if self._bound is not None:
            for p in self.params:
                if hasattr(p, "_stl_sink"):
                    del p._stl_sink
                p.grad = None
            self._bound = None

    def _written(self, i, count):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.flat[i].device))
        self._bucket_events[i].append(ev)
        self._pending[i] -= count
        if self._pending[i] == 0:
            with torch.cuda.stream(self.comm_stream):
                for e in self._bucket_events[i]:       # the views of a bucket are filled on several streams
                    self.comm_stream.wait_event(e)
                if not self.fold_scale:
                    self.flat[i].mul_(self.scale)
                if self.world > 1:
                    dist.all_reduce(self.flat[i], group=self.group)
            self._bucket_events[i] = []
        elif self._pending[i] < 0:
            raise RuntimeError("GradientReducer: a gradient was written twice in one step (finish_backward() not called?)")

    def finish_backward(self):
        """After ``loss.backward()`` of a bound model: every bucket has been launched; the current stream waits for
        the communication stream, so that ``optimizer.step()`` may follow."""
        left = [i for i, n in enumerate(self._pending) if n != 0]
        self._pending = [len(b) for b in self.buckets]
        if left:
            for ev in self._bucket_events:
                ev.clear()
            raise RuntimeError(f"GradientReducer: {len(left)} gradient bucket(s) were not filled by this backward pass")
        torch.cuda.current_stream(self.flat[0].device).wait_stream(self.comm_stream)
        for bucket, views in zip(self.buckets, self._views):      # an optimizer.zero_grad(set_to_none=True) in between
            for p, v in zip(bucket, views):
                if p.grad is None:
                    p.grad = v

    # ---- unbound mode, one bucket: pack (scaled) -> all-reduce -> unpack
    def _launch(self, i):
        bucket, flat = self.buckets[i], self.flat[i]
        views = [v.reshape(-1) for v in self.views(i)]
        grads = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in bucket]
        if self.comm_stream is not None:
            # The gradients of one bucket are accumulated on different streams (module branches and fuse rows run side
            # by side, training._branch_stream): wait for the event each parameter's hook recorded on ITS stream, not
            # only for the stream of the bucket's last gradient.
            for p in bucket:
                ev = self._events.pop(id(p), None)
                if ev is not None:
                    self.comm_stream.wait_event(ev)
            self.comm_stream.wait_stream(torch.cuda.current_stream(flat.device))
            ctx = torch.cuda.stream(self.comm_stream)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            torch._foreach_copy_(views, grads)
            flat.mul_(self.scale)
            work = dist.all_reduce(flat, group=self.group, async_op=True) if self.world > 1 else None
        self._works.append((i, work))

    def _finish(self):
        for i, work in self._works:
            if work is not None:
                work.wait()          # NCCL: makes the current stream wait for the collective
            bucket, flat = self.buckets[i], self.flat[i]
            views = self.views(i)
            if self.comm_stream is not None:
                torch.cuda.current_stream(flat.device).wait_stream(self.comm_stream)
            dst, src = [], []
            for p, v in zip(bucket, views):
                if p.grad is None:
                    p.grad = v.clone()
                else:
                    dst.append(p.grad)
                    src.append(v)
            if dst:
                torch._foreach_copy_(dst, src)       # one multi-tensor launch per bucket instead of one copy per parameter
        self._works = []

    def _check_unbound(self):
        if self._bound is not None:
            raise RuntimeError("GradientReducer is bound to a model (gradients are reduced during backward); "
                               "use finish_backward(), or unbind() first")

    def reduce_all(self):
        """All buckets, after backward has finished (unbound mode)."""
        self._check_unbound()
        for i in range(len(self.buckets)):
            self._launch(i)
        self._finish()

    def attach_hooks(self):
        """Eager steps (unbound mode): all-reduce each bucket as soon as backward has produced its last gradient; call
        ``finish_step()`` after ``loss.backward()`` and before ``optimizer.step()``."""
        self._check_unbound()
        self._pending = [len(b) for b in self.buckets]

        def hook(p):
            i = self.bucket_of[id(p)]
            if self.comm_stream is not None:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(p.device))
                self._events[id(p)] = ev
            self._pending[i] -= 1
            if self._pending[i] == 0:
                self._launch(i)

        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(hook))
        return self

    def finish_step(self):
        for i, n in enumerate(self._pending):          # parameters that received no gradient this step
            if n != 0 and not any(j == i for j, _ in self._works):
                self._launch(i)
        self._finish()
        self._pending = [len(b) for b in self.buckets]

    def detach_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []

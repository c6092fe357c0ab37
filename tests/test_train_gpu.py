"""GPU parity of the training path: train-mode forward (batch-statistics BatchNorm), PersonMSELoss and the backward
pass vs the CPU oracle (== the reference module under model.train(), see test_oracle_vs_reference.py).

Tolerances: the device path keeps activations and activation gradients in bf16, the oracle is fp32.  Train-mode
BatchNorm over a few crops amplifies storage rounding with depth, so the criteria are (a) absolute bounds vs the fp32
oracle on a well-conditioned checkpoint, (b) "no worse than the reference algorithm itself when it stores the same
tensors in bf16" (oracle bf16_storage=True) and (c) near-exact agreement with that oracle on the first layers.  Every
training kernel is separately checked against torch autograd on identical operands in test_train_kernels_gpu.py.
"""
import numpy as np
import pytest
import torch

from oracle import hrnet_oracle, pose_oracle

pytestmark = pytest.mark.gpu


RES_GAIN = 0.25


def _setup(B=2, seed=0, res_gain=RES_GAIN):
    """Synthetic checkpoint for train-mode parity.  The last BatchNorm of every residual block is damped by `res_gain`
    (trained networks have small residual branches; with unit-gain random branches the train-mode network amplifies a
    1e-6 input perturbation 45x and bf16 storage noise to 0.19 relative RMS at the heatmaps for the reference
    algorithm itself, measured in tools/train_parity.py)."""
    import stlpose_b200 as S
    sd0 = hrnet_oracle.synth_state_dict(32, seed=0)
    for k in sd0:
        if (k.endswith("bn2.weight") and "branches" in k) or k.endswith("bn3.weight"):
            sd0[k] = sd0[k] * res_gain
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, 256, 192, generator=g)
    tgt = torch.from_numpy(pose_oracle.blob_heatmaps(B, 17, 64, 48, seed=seed + 1, noise=0.0))
    tw = torch.tensor([0.0, 1.0, 1.2, 1.5])[torch.randint(0, 4, (B, 17, 1), generator=g)]
    m = S.PoseHighResolutionNet(width=32)
    m.load_state_dict(sd0, strict=True)
    return S, sd0, x, tgt, tw, m.cuda()


def _oracle_step(sd0, x, tgt, tw, bf16_storage=False):
    sd = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and "running" not in k else v.clone())
          for k, v in sd0.items()}
    heat = hrnet_oracle.hrnet_forward_train(sd, x, 32, bf16_storage=bf16_storage)
    B, J = heat.shape[:2]
    d = (heat - tgt).reshape(B, J, -1) * tw
    loss = 0.5 * (d * d).mean(dim=(0, 2)).sum() / J          # lib/loss.py:79-92
    loss.backward()
    return sd, heat.detach(), loss.item()


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def _grad_cosines(ga, gb):
    out = []
    for k, b in gb.items():
        if b.norm().item() < 1e-7:
            continue
        out.append((torch.nn.functional.cosine_similarity(ga[k].flatten(), b.flatten(), dim=0).item(), k))
    return sorted(out)


B = 4
# Tolerances (fp32 oracle == reference module; the device path stores activations and their gradients in bf16):
HEAT_REL_RMS = 0.06                  # relative RMS of the train-mode heatmaps vs the fp32 oracle
HEAT_MAX_ABS = 2e-2                  # BASELINE.json north_star tolerance, on heatmaps with abs-max ~0.3
GRAD_COS_MEDIAN, GRAD_COS_P1 = 0.93, 0.80   # per-parameter-tensor cosine similarity vs the fp32 oracle's gradients
GRAD_COS_LAST = 0.99                 # final_layer and the layers next to the loss
VS_BF16_ORACLE = 1.25                # the device may deviate from fp32 at most this much more than the reference
                                     # algorithm does when IT stores the same tensors in bf16 (hrnet_forward_train
                                     # bf16_storage=True), plus EARLY_BN on the first layers where nothing has amplified
EARLY_BN = 2e-5


def test_train_forward_and_running_stats():
    S, sd0, x, tgt, tw, m = _setup(B)
    sd, heat_ref, _ = _oracle_step(sd0, x, tgt, tw)
    sd16, heat16, _ = _oracle_step(sd0, x, tgt, tw, bf16_storage=True)
    m.train()
    heat = m(x.cuda())
    assert heat.requires_grad and heat.shape == (B, 17, 64, 48)
    got = heat.detach().cpu()
    rel, rel16, anchor = _rel(got, heat_ref), _rel(got, heat16), _rel(heat16, heat_ref)
    print(f"train-mode heatmaps: max-abs {(got - heat_ref).abs().max().item():.4f}, relative RMS vs fp32 oracle "
          f"{rel:.4f}, vs bf16-storage oracle {rel16:.4f}; bf16-storage oracle vs fp32 oracle {anchor:.4f}")
    assert rel < HEAT_REL_RMS and (got - heat_ref).abs().max().item() < HEAT_MAX_ABS
    assert rel < VS_BF16_ORACLE * anchor and rel16 < VS_BF16_ORACLE * anchor
    new = {k: v.cpu() for k, v in m.state_dict().items()}
    var_keys = [k for k in sd0 if k.endswith("running_var")]
    for i, k in enumerate(var_keys):                       # batch variance of every BatchNorm, recovered from the update
        vd, v16, v32 = ((t[k] - 0.9) / 0.1 for t in (new, sd16, sd))
        if i < 6:
            assert _rel(vd, v16) < EARLY_BN, k              # same rounding points -> same numbers until noise amplifies
        assert _rel(vd, v32) < max(VS_BF16_ORACLE * _rel(v16, v32), 1e-4) + 2e-3, k
    for k in sd0:
        if k.endswith("running_mean"):
            assert (new[k] - sd[k]).abs().max().item() < 3e-2 * max(1.0, sd[k].abs().max().item()), k
    assert int(new["bn1.num_batches_tracked"]) == 1 and int(new["stage4.2.fuse_layers.0.3.1.num_batches_tracked"]) == 1
    # eval after a train step uses the updated running statistics (weights are re-folded)
    m.eval()
    y_eval = m(x.cuda())
    y_ref = hrnet_oracle.hrnet_forward({k: v.detach() for k, v in sd.items()}, x, 32)
    assert _rel(y_eval.cpu(), y_ref) < HEAT_REL_RMS


def test_backward_matches_oracle_gradients():
    S, sd0, x, tgt, tw, m = _setup(B)
    sd, _, loss_ref = _oracle_step(sd0, x, tgt, tw)
    sd16, _, _ = _oracle_step(sd0, x, tgt, tw, bf16_storage=True)
    m.train()
    heat = m(x.cuda())
    loss = S.PersonMSELoss()(heat, tgt.cuda(), tw.cuda())
    loss.backward()
    assert abs(loss.item() - loss_ref) < 2e-2 * abs(loss_ref) + 1e-6
    gd = {}
    for name, p in m.named_parameters():
        assert p.grad is not None and p.grad.shape == sd[name].grad.shape and p.grad.dtype == torch.float32, name
        gd[name] = p.grad.detach().cpu()
        assert torch.isfinite(gd[name]).all(), name
    g32 = {k: sd[k].grad for k in gd}
    g16 = {k: sd16[k].grad for k in gd}
    dev, anchor = _grad_cosines(gd, g32), _grad_cosines(g16, g32)
    assert len(dev) > 800
    q = lambda v, f: v[int(len(v) * f)][0]
    print(f"gradient cosine vs fp32 oracle over {len(dev)} tensors: device min {dev[0][0]:.4f} / 1% {q(dev, .01):.4f} / "
          f"median {q(dev, .5):.4f};  bf16-storage oracle min {anchor[0][0]:.4f} / 1% {q(anchor, .01):.4f} / "
          f"median {q(anchor, .5):.4f}")
    for c, n in dev[:6]:
        print(f"   {n}: cos {c:.4f}")
    assert q(dev, .5) > GRAD_COS_MEDIAN and q(dev, .01) > GRAD_COS_P1
    # no worse than the reference algorithm under the same storage precision
    assert 1 - q(dev, .5) < VS_BF16_ORACLE * (1 - q(anchor, .5)) and 1 - q(dev, .01) < VS_BF16_ORACLE * (1 - q(anchor, .01))
    cos = dict((n, c) for c, n in dev)
    for k in ("final_layer.weight", "final_layer.bias", "stage4.2.fuse_layers.0.1.1.weight",
              "stage4.2.fuse_layers.0.3.0.weight"):
        assert cos[k] > GRAD_COS_LAST, (k, cos[k])


def test_unit_gain_checkpoint_tracks_bf16_storage_oracle():
    """The stock synthetic checkpoint (unit-gain residual branches) is ill-conditioned in train mode; the device must
    still be as close to the fp32 result as the reference algorithm under bf16 storage is."""
    S, sd0, x, tgt, tw, m = _setup(B, res_gain=1.0)
    sd, heat_ref, _ = _oracle_step(sd0, x, tgt, tw)
    sd16, heat16, _ = _oracle_step(sd0, x, tgt, tw, bf16_storage=True)
    m.train()
    heat = m(x.cuda())
    S.PersonMSELoss()(heat, tgt.cuda(), tw.cuda()).backward()
    got = heat.detach().cpu()
    rel, anchor = _rel(got, heat_ref), _rel(heat16, heat_ref)
    gd = {n: p.grad.detach().cpu() for n, p in m.named_parameters()}
    dev = _grad_cosines(gd, {k: sd[k].grad for k in gd})
    ref16 = _grad_cosines({k: sd16[k].grad for k in gd}, {k: sd[k].grad for k in gd})
    med = lambda v: v[len(v) // 2][0]
    print(f"unit-gain checkpoint: heat rel RMS device {rel:.3f} vs bf16-storage oracle {anchor:.3f}; "
          f"gradient cosine median device {med(dev):.3f} vs bf16-storage oracle {med(ref16):.3f}")
    assert rel < VS_BF16_ORACLE * anchor
    assert 1 - med(dev) < VS_BF16_ORACLE * (1 - med(ref16))


def test_sgd_step_reduces_loss():
    """One fine-tuning iteration as 02_train.py:208-218 runs it: forward, loss, zero_grad, backward, optimizer.step."""
    S, sd0, x, tgt, tw, m = _setup(B=4, seed=3)
    m.train()
    opt = torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9, weight_decay=5e-4)   # model_setup.py:138-139
    crit = S.PersonMSELoss()
    losses = []
    for _ in range(4):
        out = S.forward_pass(m, x.cuda(), "HRNet", device="cuda", flip=False)
        loss = crit(out, tgt.cuda(), tw.cuda())
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses


def test_graph_captured_step_matches_eager_steps():
    """TrainStep (forward + loss + backward + SGD update replayed as one CUDA graph) vs the same steps issued eagerly."""
    S, sd0, x, tgt, tw, m_eager = _setup(B=4, seed=5)
    m_graph = S.PoseHighResolutionNet(width=32)
    m_graph.load_state_dict(sd0, strict=True)
    m_graph = m_graph.cuda()
    mk = lambda m: torch.optim.SGD(m.parameters(), lr=0.02, momentum=0.9, weight_decay=5e-4)
    opt_e, crit = mk(m_eager), S.PersonMSELoss()
    step = S.TrainStep(m_graph, mk(m_graph), crit, batch=4)
    assert step.graph is not None
    for k, v in m_graph.state_dict().items():            # warm-up and capture left the checkpoint untouched
        assert torch.equal(v.cpu(), sd0[k]), k
    m_eager.train()
    xs, tg, tws = x.cuda(), tgt.cuda(), tw.cuda()
    le, lg = [], []
    for _ in range(3):
        out = S.forward_pass(m_eager, xs, "HRNet", device="cuda", flip=False)
        loss = crit(out, tg, tws)
        opt_e.zero_grad(); loss.backward(); opt_e.step()
        le.append(loss.item())
        lg.append(step(x, tgt, tw).item())                # host tensors in, device loss out
    assert all(abs(a - b) < 2e-2 * abs(a) for a, b in zip(le, lg)), (le, lg)
    assert lg[-1] < lg[0]
    sd_e, sd_g = m_eager.state_dict(), m_graph.state_dict()
    assert int(sd_g["bn1.num_batches_tracked"]) == 3
    for k in ("conv1.weight", "stage4.2.branches.0.3.conv2.weight", "final_layer.weight", "bn1.running_var"):
        assert _rel(sd_g[k].cpu(), sd_e[k].cpu()) < 2e-2, k
    m_graph.eval()                                         # folded weights are rebuilt from the trained parameters
    m_eager.eval()
    assert _rel(m_graph(xs).cpu(), m_eager(xs).cpu()) < 5e-2


def test_bound_gradient_buckets_give_the_same_step():
    """TrainStep with a GradientReducer bound to the model (kernels write parameter gradients straight into the flat
    buckets, which are handed to the communication stream as backward fills them; one process: no collective) must leave
    bit-identical parameters to the plain TrainStep, whose gradients go through autograd's accumulation nodes."""
    from stlpose_b200.parallel import GradientReducer
    S, sd0, x, tgt, tw, m_a = _setup(B=4, seed=9)
    m_b = S.PoseHighResolutionNet(width=32)
    m_b.load_state_dict(sd0, strict=True)
    m_b = m_b.cuda()
    mk = lambda m: torch.optim.SGD(m.parameters(), lr=0.02, momentum=0.9, weight_decay=5e-4)
    crit = S.PersonMSELoss()
    red = GradientReducer(m_b.parameters(), local_batch=4, bucket_bytes=4 << 20)
    assert len(red.buckets) > 4 and red.scale == 1.0
    step_a = S.TrainStep(m_a, mk(m_a), crit, batch=4)
    step_b = S.TrainStep(m_b, mk(m_b), crit, batch=4, reducer=red)
    assert step_a.optimizer_in_graph and step_b.optimizer_in_graph
    for _ in range(3):
        la, lb = step_a(x, tgt, tw).item(), step_b(x, tgt, tw).item()
        assert la == lb
    sa, sb = m_a.state_dict(), m_b.state_dict()
    bad = [k for k in sa if not torch.equal(sa[k], sb[k])]
    assert not bad, f"{len(bad)} tensors differ, e.g. {bad[:3]}"
    # every .grad is still a view of its bucket (nothing re-allocated them)
    for bucket, flat in zip(red.buckets, red.flat):
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
        assert all(lo <= p.grad.data_ptr() < hi for p in bucket)


def test_scheduler_changes_reach_the_captured_optimizer():
    """The captured optimizer step bakes lr / momentum / weight decay into its kernels: TrainStep keeps a
    signature of param_groups and re-captures when a scheduler changes them.  lr = 0 must therefore freeze the
    parameters, and restoring lr must move them again; training state survives the re-capture."""
    S, sd0, x, tgt, tw, m = _setup(B=2, seed=4)
    opt = torch.optim.SGD(m.parameters(), lr=0.05, weight_decay=5e-4)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.0)       # lib/model_setup.py: StepLR
    step = S.TrainStep(m, opt, S.PersonMSELoss(), batch=2)
    assert step.optimizer_in_graph and step.captures == 1
    step(x, tgt, tw)
    w1 = m.conv1.weight.detach().clone()
    rv1 = m.bn1.running_var.detach().clone()
    assert not torch.equal(w1.cpu(), sd0["conv1.weight"])
    sched.step()                                                               # lr -> 0
    step(x, tgt, tw)
    assert step.captures == 2
    assert torch.equal(m.conv1.weight.detach(), w1)                            # lr = 0: no update ...
    assert not torch.equal(m.bn1.running_var, rv1)                             # ... but the step itself ran
    assert int(m.bn1.num_batches_tracked) == 2                                 # warm-up of the re-capture did not count
    for g in opt.param_groups:
        g["lr"] = 0.05
    step(x, tgt, tw)
    assert step.captures == 3 and not torch.equal(m.conv1.weight.detach(), w1)


def test_uncapturable_optimizers_step_after_the_replay():
    """torch.optim.Adam(capturable=False) - the reference's other optimizer (lib/model_setup.py:141) - cannot be captured:
    the graph then holds forward + loss + backward and the optimizer steps eagerly after each replay."""
    S, sd0, x, tgt, tw, m = _setup(B=2, seed=6)
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    step = S.TrainStep(m, opt, S.PersonMSELoss(), batch=2)
    assert step.graph is not None and not step.optimizer_in_graph
    assert len(opt.state) == 0                                                 # warm-up state was removed
    losses = [step(x, tgt, tw).item() for _ in range(4)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    assert int(next(iter(opt.state.values()))["step"]) == 4


def test_w48_train_step():
    """HRNet-W48 channel counts (48/96/192/384) run on the tensor-core gradient kernels through the next larger MMA shape
    (test_train_kernels_gpu.py checks those shapes one by one).  A few SGD steps on a small crop size must run, stay
    finite and reduce the loss, and the parameter gradients of one step must agree in direction with torch autograd of
    the same network in fp32."""
    import stlpose_b200 as S
    torch.manual_seed(0)
    m = S.PoseHighResolutionNet(width=48, image_size=(128, 96)).cuda().train()
    sd0 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, 128, 96, generator=g)
    tgt = torch.from_numpy(pose_oracle.blob_heatmaps(2, 17, 32, 24, seed=5, noise=0.0))
    tw = torch.ones(2, 17, 1)
    crit = S.PersonMSELoss()
    loss = crit(m(x.cuda()), tgt.cuda(), tw.cuda())
    loss.backward()
    sd = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and "running" not in k else v.clone())
          for k, v in sd0.items()}
    heat = hrnet_oracle.hrnet_forward_train(sd, x, 48)
    d = (heat - tgt).reshape(2, 17, -1) * tw
    (0.5 * (d * d).mean(dim=(0, 2)).sum() / 17).backward()
    for name in ("final_layer.weight", "stage4.2.fuse_layers.0.1.0.weight", "stage4.2.branches.0.3.conv2.weight"):
        a, b = dict(m.named_parameters())[name].grad.cpu().flatten(), sd[name].grad.flatten()
        assert torch.nn.functional.cosine_similarity(a, b, dim=0).item() > 0.85, name   # batch of 2, undamped random net: see module docstring
    opt = torch.optim.SGD(m.parameters(), lr=2e-3)   # default-init heads are large: small plain steps
    losses = []
    for _ in range(3):
        out = m(x.cuda())
        l = crit(out, tgt.cuda(), tw.cuda())
        opt.zero_grad(); l.backward(); opt.step()
        losses.append(l.item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses


def _three_sgd_steps(monkeypatch, parallel, delay_seed=None):
    """Three eager SGD steps on a side stream; parallel: module branches / fuse rows / weight gradients on their own
    streams.  delay_seed: inject seeded device-side sleeps in front of random units (forward and backward) and on the
    weight-gradient side stream, so that the interleaving of the streams differs from the undisturbed schedule."""
    import random
    from stlpose_b200 import training
    monkeypatch.setattr(training, "BRANCH_STREAMS", parallel)
    monkeypatch.setattr(training, "SIDE_WGRAD", parallel)
    if delay_seed is not None:
        rng = random.Random(delay_seed)

        def nap(p=0.15, hi=400_000):
            if rng.random() < p:
                torch.cuda._sleep(rng.randrange(20_000, hi))        # on the current stream, 10-200 us

        fwd, bwd, wg = training._convbn, training._ConvBN.backward, training._conv_wgrad

        def convbn(*a, **k):
            nap()
            return fwd(*a, **k)

        def backward(ctx, dy):
            nap()
            return bwd(ctx, dy)

        def conv_wgrad(*a, side=None, **k):
            if side is not None and rng.random() < 0.15:
                with torch.cuda.stream(side):
                    torch.cuda._sleep(rng.randrange(20_000, 400_000))
            return wg(*a, side=side, **k)

        monkeypatch.setattr(training, "_convbn", convbn)
        monkeypatch.setattr(training._ConvBN, "backward", staticmethod(backward))
        monkeypatch.setattr(training, "_conv_wgrad", conv_wgrad)
    S, sd0, x, tgt, tw, m = _setup(B=6, seed=3)
    m.train()
    opt = torch.optim.SGD(m.parameters(), lr=1e-2, momentum=0.9, weight_decay=5e-4)
    crit = S.PersonMSELoss()
    xd, td, wd = x.cuda(), tgt.cuda().float(), tw.cuda()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                 # not the default stream (see TrainStep's note on accumulation nodes)
        for _ in range(3):
            loss = crit(S.forward_pass(m, xd, "HRNet", device="cuda", flip=False), td, wd)
            opt.zero_grad()
            loss.backward()
            opt.step()
            m.invalidate_packed_weights()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    monkeypatch.undo()
    return {k: v.detach().clone() for k, v in m.state_dict().items()}, float(loss.detach())


def _assert_same_state(a, b):
    (sa, la), (sb, lb) = a, b
    assert la == lb
    bad = [k for k in sa if not torch.equal(sa[k], sb[k])]
    assert not bad, f"{len(bad)} tensors differ, e.g. {bad[:3]}"


def test_parallel_streams_give_the_same_training_steps(monkeypatch):
    """The stream-parallel schedule (module branches and fuse rows side by side, weight gradients next to input
    gradients) only reorders independent launches: three eager SGD steps must leave bit-identical parameters and
    running statistics to the same steps issued on one stream.  Every training kernel is deterministic (fixed-order
    slab reductions, no floating-point atomics), so any difference is a missing dependency between streams."""
    _assert_same_state(_three_sgd_steps(monkeypatch, True), _three_sgd_steps(monkeypatch, False))


@pytest.mark.parametrize("seed", [1, 2])
def test_parallel_streams_with_injected_delays(monkeypatch, seed):
    """Race detector for the cross-stream hand-offs (compute-sanitizer is closed on the B200 pool, profiles/
    r02_sanitizer.md): seeded device-side sleeps in front of random units and on the weight-gradient side stream change
    which stream runs ahead; a read before its producer finished, or a block reused while another stream still reads it,
    would change bits against the single-stream schedule."""
    _assert_same_state(_three_sgd_steps(monkeypatch, True, delay_seed=seed), _three_sgd_steps(monkeypatch, False))


@pytest.mark.parametrize("momentum,wd,nesterov", [(0.9, 5e-4, False), (0.0, 0.0, False), (0.9, 1e-3, True)])
def test_fused_sgd_step_equals_torch(momentum, wd, nesterov):
    """stlpose_b200.optim.fused_sgd_step (one launch for all parameter tensors) vs torch.optim.SGD.step() on the same
    gradients: same operations in the same order, so the parameters and momentum buffers agree bit for bit over
    several steps (the first one creates the buffers), including tensors whose size is not a multiple of 4."""
    from stlpose_b200 import optim as O
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(64, 3, 3, 3), (17,), (17, 32, 1, 1), (256,), (33, 7), (1,), (128, 128, 3, 3)]
    pa = [torch.randn(s, device="cuda", generator=g).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = torch.optim.SGD(pa, lr=0.05, momentum=momentum, weight_decay=wd, nesterov=nesterov)
    ob = torch.optim.SGD(pb, lr=0.05, momentum=momentum, weight_decay=wd, nesterov=nesterov)
    assert O.supports(ob) and not O.supports(torch.optim.SGD(pb, lr=0.1, momentum=0.9, dampening=0.1))
    tables = O.make_tables(ob)
    for step in range(4):
        for a, b in zip(pa, pb):
            a.grad = torch.randn(a.shape, device="cuda", generator=g)
            b.grad = a.grad.clone()
        if step == 2:
            for o in (oa, ob):
                o.param_groups[0]["lr"] = 0.01                       # a scheduler step
        oa.step()
        O.fused_sgd_step(ob, tables)
        for a, b in zip(pa, pb):
            assert torch.equal(a, b), (step, a.shape)
            if momentum:
                assert torch.equal(oa.state[a]["momentum_buffer"], ob.state[b]["momentum_buffer"])
    assert set(ob.state_dict()["state"].keys()) == set(oa.state_dict()["state"].keys())

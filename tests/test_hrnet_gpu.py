"""GPU parity: whole-network forward, flip test and decode vs the reference-generated goldens / CPU oracle.

Tolerances are BASELINE.json's: heatmaps within 2e-2 max-abs (bf16 tensor-core arithmetic), keypoints within
0.25 px, argmax compared only where the top-2 margin exceeds the heatmap tolerance.
"""
import numpy as np
import pytest
import torch

from oracle import hrnet_oracle, pose_oracle

pytestmark = pytest.mark.gpu
HEAT_TOL = 2e-2


def _model(width, image_size):
    import stlpose_b200 as S
    m = S.PoseHighResolutionNet(width=width, image_size=image_size)
    m.load_state_dict(hrnet_oracle.synth_state_dict(width, seed=0), strict=True)
    return m.cuda().eval()


def test_forward_w32_matches_reference_golden(golden, golden_inputs):
    y_ref = golden("hrnet_w32_fwd.npz")["y"]
    m = _model(32, (256, 192))
    y = m(torch.from_numpy(golden_inputs["x_w32"]).cuda())
    assert y.shape == (2, 17, 64, 48) and y.dtype == torch.float32 and y.is_cuda
    err = np.abs(y.cpu().numpy() - y_ref).max()
    assert err < HEAT_TOL, f"heatmap max-abs error {err}"


def test_forward_w48_matches_reference_golden(golden, golden_inputs):
    y_ref = golden("hrnet_w48_fwd.npz")["y"]
    m = _model(48, (384, 288))
    y = m(torch.from_numpy(golden_inputs["x_w48"]).cuda())
    err = np.abs(y.cpu().numpy() - y_ref).max()
    assert y.shape == (1, 17, 96, 72) and err < HEAT_TOL, f"heatmap max-abs error {err}"


def test_concatenated_downsample_matches_the_two_launch_form(golden_inputs, monkeypatch):
    """layer1.0's conv3 + downsample as ONE 1x1 convolution over [conv2 output | block input] (plan.cu, the default) vs
    the two launches with a bf16 round trip of the downsample branch in between (STLPOSE_FUSE_DOWNSAMPLE=0): same
    heatmaps up to that rounding (measured 7.5e-3 max-abs; both are within HEAT_TOL of the reference), and two kernel
    launches fewer per forward (the fused pair also takes the next block's conv1 with it, link_tc.cu)."""
    x = torch.from_numpy(golden_inputs["x_w32"]).cuda()
    m1 = _model(32, (256, 192))
    y1 = m1(x)
    monkeypatch.setenv("STLPOSE_FUSE_DOWNSAMPLE", "0")
    m0 = _model(32, (256, 192))
    y0 = m0(x)
    assert (y1 - y0).abs().max().item() < 0.5 * HEAT_TOL
    assert m0.launches_per_forward() == m1.launches_per_forward() + 2
    # conv3 (+ downsample) + the next block's conv1 of layer1.0 / .1 / .2 as one kernel each (link_tc.cu) is bit-identical
    # to the two launches (STLPOSE_FUSE_LINK=0)
    monkeypatch.delenv("STLPOSE_FUSE_DOWNSAMPLE")
    monkeypatch.setenv("STLPOSE_FUSE_LINK", "0")
    m2 = _model(32, (256, 192))
    y2 = m2(x)
    assert torch.equal(y1, y2)
    assert m2.launches_per_forward() == m1.launches_per_forward() + 3
    # the last fuse row + the heatmap head in one pass (fuse_head_kernel, fp32 FMA chain on the bf16-rounded fused map) vs
    # fuse_sum + the tensor-core head convolution (STLPOSE_FUSE_HEAD=0): same operands, fp32 summation order differs
    monkeypatch.delenv("STLPOSE_FUSE_LINK")
    monkeypatch.setenv("STLPOSE_FUSE_HEAD", "0")
    m3 = _model(32, (256, 192))
    y3 = m3(x)
    assert (y1 - y3).abs().max().item() < 1e-5 * max(1.0, y3.abs().max().item())
    assert m3.launches_per_forward() == m1.launches_per_forward() + 1


def test_flip_test_and_keypoints_vs_oracle():
    import stlpose_b200 as S
    B = 5
    sd = hrnet_oracle.synth_state_dict(32, seed=0)
    x = torch.randn(B, 3, 256, 192, generator=torch.Generator().manual_seed(11))
    # oracle: lib/inference.py:18-26 then lib/pose_parsing.py:58-92
    h0 = hrnet_oracle.hrnet_forward(sd, x, 32).numpy()
    h1 = hrnet_oracle.hrnet_forward(sd, x.flip(3), 32).numpy()
    heat_ref = pose_oracle.flip_average(h0, h1)
    c, s = pose_oracle.synth_boxes(B, seed=2)
    preds_ref, maxv_ref, coords_ref = pose_oracle.get_final_preds(heat_ref, c, s)

    m = _model(32, (256, 192))
    heat = S.forward_pass(m, x.cuda(), "HRNet", device="cuda", flip=True)
    assert np.abs(heat.cpu().numpy() - heat_ref).max() < HEAT_TOL
    # the batched flip pass equals running the mirrored images on their own
    both = m.forward_flip_pair(x.cuda())
    alone = m(x.flip(3).cuda())
    assert torch.equal(both[B:], alone)
    assert torch.equal(both[:B], m(x.cuda()))

    preds, maxv, coords = S.get_final_preds_hrnet(heat, c, s)
    assert np.abs(maxv - maxv_ref).max() < HEAT_TOL
    # margin gate: compare locations only where the oracle's top-2 gap is larger than what bf16 can move
    flat = np.sort(heat_ref.reshape(B, 17, -1), axis=2)
    margin = flat[:, :, -1] - flat[:, :, -2]
    sure = margin > 2 * HEAT_TOL
    if sure.any():
        k = (s[:, 0] * 200.0 / 48.0)[:, None, None]
        assert (np.abs(preds - preds_ref) / k)[sure].max() <= 0.25 + 1e-3
    # decode of our own heatmaps is self-consistent with the oracle decode bit for bit
    p2, m2, c2 = pose_oracle.get_final_preds(heat.cpu().numpy(), c, s)
    assert np.array_equal(c2, coords) and np.array_equal(m2, maxv)


def test_module_contract():
    import stlpose_b200 as S
    m = S.PoseHighResolutionNet()
    keys = [k for k, _ in hrnet_oracle.hrnet_schema(32)]
    assert list(m.state_dict().keys()) == keys
    with pytest.raises(NotImplementedError):
        S.forward_pass(m, torch.zeros(1, 3, 256, 192), "OpenPose")
    with pytest.raises(S.StlError):
        m.eval()(torch.zeros(1, 3, 256, 192))          # CPU tensor: no fallback
    mc = _model(32, (256, 192))
    assert mc(torch.zeros(0, 3, 256, 192, device="cuda")).shape == (0, 17, 64, 48)
    # re-packing after an in-place parameter update changes the output
    x = torch.randn(1, 3, 256, 192, device="cuda")
    y0 = mc(x).clone()
    with torch.no_grad():
        mc.final_layer.bias.add_(1.0)
    y1 = mc(x)
    assert torch.allclose(y1, y0 + 1.0, atol=1e-5)
    # DataParallel-wrapped module works as the reference's callers pass it (03_evaluate.py:100)
    out = S.forward_pass(torch.nn.DataParallel(mc, device_ids=[0]), x, "HRNet", device="cuda", flip=False)
    assert torch.equal(out, y1)


def test_pipeline_graph_matches_eager_calls():
    """KeypointPipeline (CUDA graph + double-buffered H2D staging) == forward_pass + get_final_preds_hrnet."""
    import stlpose_b200 as S
    from stlpose_b200.pipeline import KeypointPipeline
    B = 6
    m = _model(32, (256, 192))
    pipe = KeypointPipeline(m, B, (256, 192), flip=True, use_graph=True)
    for seed in (1, 2, 3):
        x = torch.randn(B, 3, 256, 192, generator=torch.Generator().manual_seed(seed)).pin_memory()
        c_np, s_np = pose_oracle.synth_boxes(B, seed=seed)
        c = torch.from_numpy(c_np).float().pin_memory()
        s = torch.from_numpy(s_np).float().pin_memory()
        p_host, m_host = pipe(x, c, s)
        torch.cuda.synchronize()
        heat = S.forward_pass(m, x.cuda(), "HRNet", device="cuda", flip=True)
        preds, maxv, _ = S.get_final_preds_hrnet(heat, c_np.astype(np.float32), s_np.astype(np.float32))
        assert np.array_equal(p_host.numpy(), preds) and np.array_equal(m_host.numpy(), maxv)


def test_pipeline_recaptures_when_the_model_reallocates_its_buffers():
    """The captured steps hold raw addresses of the model's workspace and weight arena: an eager call with a
    larger batch at the same resolution makes the model reallocate its workspace; the pipeline must notice (generation
    counter) and re-capture instead of replaying into freed memory.  ShardedKeypointInference runs its slice through
    the same captured pipeline and must agree with the eager path."""
    import stlpose_b200 as S
    from stlpose_b200.parallel import ShardedKeypointInference
    from stlpose_b200.pipeline import KeypointPipeline
    B = 3
    m = _model(32, (256, 192))
    pipe = KeypointPipeline(m, B, (256, 192), flip=True, use_graph=True)
    x = torch.randn(B, 3, 256, 192, generator=torch.Generator().manual_seed(4))
    c_np, s_np = pose_oracle.synth_boxes(B, seed=4)
    c, s = torch.from_numpy(c_np).float(), torch.from_numpy(s_np).float()
    pipe.x.copy_(x); pipe.center.copy_(c); pipe.scale.copy_(s)
    pipe.step()
    torch.cuda.synchronize()
    before = (pipe.preds.clone(), pipe.maxvals.clone())
    gen = pipe._generation
    big = torch.randn(24, 3, 256, 192, generator=torch.Generator().manual_seed(5)).cuda()
    keep = [torch.empty(64 << 20, dtype=torch.uint8, device="cuda") for _ in range(4)]     # churn the allocator
    m(big)                                                    # larger workspace at the same (H, W): reallocation
    del keep
    torch.randn(1 << 24, device="cuda").mul_(3.0)             # scribble over whatever was freed
    assert m._buffer_generation != gen
    pipe.step()
    torch.cuda.synchronize()
    assert pipe._generation == m._buffer_generation
    assert torch.equal(pipe.preds, before[0]) and torch.equal(pipe.maxvals, before[1])
    sh = ShardedKeypointInference(m, flip=True)               # one process: the whole batch is this rank's slice
    preds, maxv = sh.run(x, c_np.astype(np.float32), s_np.astype(np.float32))
    assert torch.equal(preds, before[0]) and torch.equal(maxv, before[1])
    heat = S.forward_pass(m, x.cuda(), "HRNet", device="cuda", flip=True)
    p2, m2, _ = S.get_final_preds_hrnet(heat, c_np.astype(np.float32), s_np.astype(np.float32))
    assert np.array_equal(preds.cpu().numpy(), p2) and np.array_equal(maxv.cpu().numpy(), m2)


def test_batch_independence_at_bench_scale():
    """Crops are independent units: any crop's heatmaps are bit-identical whatever batch (and tile boundaries) it
    travels in.  Run at BASELINE.json config 2's size (512 crops = 1 024 images with the mirrored pass), where every layer
    spans many tiles and CTAs walk several tiles each."""
    B = 512
    m = _model(32, (256, 192))
    x = torch.randn(B, 3, 256, 192, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    big = m.forward_flip_pair(x).clone()
    for lo, hi in ((0, 3), (254, 257), (509, 512)):
        small = m.forward_flip_pair(x[lo:hi])
        n = hi - lo
        assert torch.equal(small[:n], big[lo:hi]) and torch.equal(small[n:], big[B + lo:B + hi])
    assert torch.isfinite(big).all()
    # ... and at that scale the heatmaps still match the reference algorithm: crops from the first, a middle and the last
    # tile region of the batch (plain and mirrored pass) against the fp32 oracle, BASELINE's 2e-2 max-abs
    sd = hrnet_oracle.synth_state_dict(32, seed=0)
    idx = [0, 255, 511]
    xs = x[idx].cpu()
    ref = hrnet_oracle.hrnet_forward(sd, xs, 32)
    ref_f = hrnet_oracle.hrnet_forward(sd, xs.flip(3), 32)
    err = max((big[idx].cpu() - ref).abs().max().item(), (big[[B + i for i in idx]].cpu() - ref_f).abs().max().item())
    assert err < HEAT_TOL, err


def test_boxes_to_keypoints_pipeline_vs_oracle():
    """The evaluation loop either side of the network, as 04_evaluate_vases_qualitatively.py:206-220 / 03_evaluate.py:124-152
    run it: person boxes -> TransformDetection crops -> ToTensor + Normalize -> forward_pass(flip=True) ->
    get_final_preds_hrnet (and the PCK of the loop), device path vs the oracle chain on the same image."""
    import stlpose_b200 as S
    from oracle.make_golden import crop_inputs
    from stlpose_b200 import metrics
    from stlpose_b200.transforms import TransformDetection
    img, boxes = crop_inputs()
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    # oracle chain
    dets, c_ref, s_ref = pose_oracle.transform_detection(img, boxes)
    x_ref = (torch.from_numpy(dets).float().div(255) - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    sd = hrnet_oracle.synth_state_dict(32, seed=0)
    h0 = hrnet_oracle.hrnet_forward(sd, x_ref, 32).numpy()
    h1 = hrnet_oracle.hrnet_forward(sd, x_ref.flip(3), 32).numpy()
    heat_ref = pose_oracle.flip_average(h0, h1)
    preds_ref, maxv_ref, _ = pose_oracle.get_final_preds(heat_ref, c_ref, s_ref)
    # device chain: nothing but the image and the boxes crosses the host boundary
    m = _model(32, (256, 192))
    x, c, s = TransformDetection().extract_normalized(img, boxes, mean, std)
    assert torch.equal(x.cpu(), x_ref) and np.array_equal(c, c_ref) and np.array_equal(s, s_ref)
    heat = S.forward_pass(m, x, "HRNet", device="cuda", flip=True)
    assert np.abs(heat.cpu().numpy() - heat_ref).max() < HEAT_TOL
    preds, maxv, _ = S.get_final_preds_hrnet(heat, c, s)
    flat = np.sort(heat_ref.reshape(len(boxes), 17, -1), axis=2)
    sure = (flat[:, :, -1] - flat[:, :, -2]) > 2 * HEAT_TOL
    k = (s_ref[:, 0] * 200.0 / 48.0)[:, None, None]                  # image pixels per heatmap pixel
    if sure.any():
        assert (np.abs(preds - preds_ref) / k)[sure].max() <= 0.25 + 1e-3
    assert np.abs(maxv - maxv_ref).max() < HEAT_TOL
    # PCK of the loop against a target built from the oracle's own keypoints: same verdicts on both paths
    tgt = np.zeros_like(heat_ref)
    _, _, coords_ref = pose_oracle.get_final_preds(heat_ref, c_ref, s_ref)
    for n in range(len(boxes)):
        for j in range(17):
            tgt[n, j, int(coords_ref[n, j, 1]), int(coords_ref[n, j, 0])] = 1.0
    acc_dev = metrics.accuracy(heat, tgt)
    acc_ref = pose_oracle.accuracy(heat_ref, tgt)
    assert acc_dev[2] == acc_ref[2]
    if sure.all():
        assert abs(acc_dev[1] - acc_ref[1]) < 1e-6

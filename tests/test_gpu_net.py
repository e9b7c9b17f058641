"""GPU parity tests of the conv net (stage 2) and of the whole frame against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import msi_np, net_torch
from matryodshka_b200 import ops, synth
from matryodshka_b200.msi import MSI, MSIConfig
from matryodshka_b200.runtime import MSIPipeline, NetEngine

pytestmark = pytest.mark.gpu
F32 = np.float32
DEV = "cuda"
TOL = 1e-3
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def _oracle_net(x, c_out, wts, ngf):
    with torch.no_grad():
        pred, feats = net_torch.msi_coord_train_net(torch.from_numpy(x), c_out, wts, ngf=ngf, return_feats=True)
    return pred.numpy(), {k: v.numpy() for k, v in feats.items()}


def _check_net(eng, x, wts, c_out, ngf, B):
    pred = eng.forward(_t(x)).cpu().numpy()
    want, feats = _oracle_net(x, c_out, wts, ngf)
    worst = {}
    for scope, f in feats.items():
        got = eng.read_activation(scope, B).cpu().numpy()
        worst[scope] = float(np.abs(got - f).max())
    err = np.abs(pred - want).max()
    assert err < TOL, (err, worst)
    return err, worst


@pytest.mark.parametrize("H,W,P,ngf,B", [(16, 32, 4, 8, 1), (32, 64, 4, 16, 2)])
def test_simt_net_matches_oracle(H, W, P, ngf, B):
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (B, H, W, 6 * P)).astype(F32)
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    eng = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, conv_impl="simt")
    err, worst = _check_net(eng, x, wts, 2 * P, ngf, B)
    assert max(worst.values()) < TOL, worst


@pytest.mark.parametrize("precision,tol", [("fp16x3", TOL), ("fp16", 3e-2), ("fp16_fp8x", TOL)])
def test_tcgen05_net_matches_oracle(precision, tol):
    """The tensor-core path at ngf = 64 on a small frame: every layer's activation and the head.  fp16_fp8x: the layers
    with Cout >= 128 form both cross terms of the split product in ONE e4m3 MMA (2 MMA units per product instead of 3);
    expected ~2e-4 on the prediction (scripts/exp_fp8_cross.py), inside the path's 1e-3."""
    H, W, P, ngf, B = 32, 64, 32, 64, 1
    ref, src = synth.ods_pair(B, H, W)
    d = msi_np.inv_depths(1, 100, P)
    x = msi_np.format_network_input(ref * 2 - 1, src * 2 - 1, synth.identity_poses(1), synth.identity_poses(1), d,
                                    synth.intrinsics(1))
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    eng = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, conv_impl="tcgen05", precision=precision)
    pred = eng.forward(_t(x)).cpu().numpy()
    want, feats = _oracle_net(x, 2 * P, wts, ngf)
    worst = {s: float(np.abs(eng.read_activation(s, B).cpu().numpy() - f).max()) for s, f in feats.items()}
    err = np.abs(pred - want).max()
    print(f"tcgen05 {precision}: max|pred - oracle| = {err:.3e}, worst activation {max(worst.values()):.3e}")
    assert err < tol, (err, worst)
    assert max(worst.values()) < max(tol, 2e-3), worst


def test_tcgen05_matches_simt_batch2_odd_tiles():
    """tcgen05 vs the fp32 SIMT kernel of this library on a shape with partial tiles and B = 2."""
    H, W, P, ngf, B = 48, 80, 32, 64, 2
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, (B, H, W, 6 * P)).astype(F32)
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    a = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, conv_impl="tcgen05")
    b = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, conv_impl="simt")
    pa, pb = a.forward(_t(x)), b.forward(_t(x))
    assert (pa - pb).abs().max().item() < TOL
    c = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, conv_impl="tcgen05", precision="fp16_fp8x")
    assert (c.forward(_t(x)) - pb).abs().max().item() < TOL


def test_golden_fixture_through_c_abi():
    """The committed fixture (tests/golden/msi_small.npz) through the mirror API, SIMT net (ngf = 8)."""
    from tests.golden.make_golden import inputs_small, P, NGF
    z = np.load(os.path.join(GOLDEN, "msi_small.npz"))
    i = inputs_small()
    m = MSI(weights=i["weights"], config=MSIConfig(conv_impl="simt", ngf=NGF))
    planes = list(i["planes"])
    out, net_input = m.infer_msi(_t(i["src"]), _t(i["ref"]), None, None, i["ref_pose"], i["src_pose"],
                                 i["intrinsics"], "blend_psv", P, planes, "blend_weights_alphas_psv", ngf=NGF)
    assert np.abs(net_input.cpu().numpy() - z["psv"]).max() < TOL
    assert np.abs(out["rgba_layers"].cpu().numpy() - z["rgba_layers"]).max() < TOL
    pred = torch.cat([out["blend_weights"], out["alphas"]], -1).cpu().numpy() * 2 - 1
    assert np.abs(pred - z["pred"]).max() < TOL
    res = m.msi_render_equirect(out["rgba_layers"], np.eye(4, dtype=F32)[None], i["tgt_pos"], planes)
    assert np.abs(res["rgb"].cpu().numpy() - z["render"]).max() < TOL
    assert np.abs(res["depth"].cpu().numpy() - z["depth"]).max() < TOL
    assert np.abs(res["rgb_u8"].cpu().numpy().astype(int) - z["render_u8"].astype(int)).max() <= 1
    assert set(out.keys()) == {"rgba_layers", "blend_weights", "alphas", "psv"}


@pytest.mark.parametrize("conv_impl,precision", [("tcgen05", "fp16x3"), ("tcgen05", "fp16_fp8x")])
def test_full_frame_pipeline_matches_oracle(conv_impl, precision):
    """Config C2 (640x320, 32 spheres, B=1) end to end vs the oracle: rendered view within 1e-3."""
    H, W, P, ngf = 320, 640, 32, 64
    ref, src = synth.ods_pair(1, H, W)
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    tp = synth.target_positions(1)
    pipe = MSIPipeline(wts, H, W, P, ngf, batch=1, device=DEV, conv_impl=conv_impl, precision=precision)
    pipe.set_inputs(ref, src, tgt_pos=tp)
    pipe.step()
    pipe.step()  # second call replays the CUDA graph
    torch.cuda.synchronize()
    planes = msi_np.inv_depths(1, 100, P)
    out, _ = msi_np.infer_msi(src, ref, synth.identity_poses(1), synth.identity_poses(1), synth.intrinsics(1), P,
                              planes, wts, ngf=ngf)
    eye = np.eye(4, dtype=F32)[None]
    want = msi_np.msi_render_equirect_view(out["rgba_layers"], eye, tp, planes)
    wdep = msi_np.msi_render_equirect_depth(out["rgba_layers"], eye, tp, planes)
    rgba_err = np.abs(pipe.rgba.cpu().numpy() - out["rgba_layers"]).max()
    err = np.abs(pipe.out["rgb"].cpu().numpy() - want).max()
    derr = np.abs(pipe.out["depth"].cpu().numpy() - wdep).max()
    print(f"full frame {precision}: max-abs rgba {rgba_err:.3e}, view {err:.3e}, depth {derr:.3e}")
    assert rgba_err < TOL and err < TOL and derr < TOL, (rgba_err, err, derr)
    # end-to-end host path returns the same pixels
    rgb8, dep8 = pipe.step_e2e(torch.from_numpy(ref), torch.from_numpy(src))
    assert torch.equal(rgb8, pipe.out["rgb_u8"].cpu())


def test_batched_frames_equal_single_frames():
    """Frames are independent: a batch of 3 must equal three B=1 runs bit for bit."""
    H, W, P, ngf = 32, 64, 32, 64
    ref, src = synth.ods_pair(3, H, W)
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    tp = synth.target_positions(3)
    big = MSIPipeline(wts, H, W, P, ngf, batch=3, device=DEV, use_graph=False)
    big.set_inputs(ref, src, tgt_pos=tp)
    big.step()
    one = MSIPipeline(wts, H, W, P, ngf, batch=1, device=DEV, use_graph=False)
    for b in range(3):
        one.set_inputs(ref[b:b + 1], src[b:b + 1], tgt_pos=tp[b:b + 1])
        one.step()
        assert torch.equal(one.out["rgb"][0], big.out["rgb"][b])
        assert torch.equal(one.out["depth_u8"][0], big.out["depth_u8"][b])


def test_multi_gpu_gather_equals_single_gpu():
    """N ranks x 2 frames + NCCL all-gather == one GPU running all frames (bit for bit)."""
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(root, "tests", "_nccl_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "NCCL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    # the fused gather (render kernel -> every rank's symmetric buffer) either matched NCCL or was skipped loudly
    assert "FUSED_OK" in r.stdout or "FUSED_SKIP" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    print(r.stdout[-600:])


def _pipeline_vs_oracle(H, W, P, B, frames_to_check, ngf=64):
    ref, src = synth.ods_pair(B, H, W)
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    tp = synth.target_positions(B)
    pipe = MSIPipeline(wts, H, W, P, ngf, batch=B, device=DEV)
    pipe.set_inputs(ref, src, tgt_pos=tp)
    pipe.step()
    torch.cuda.synchronize()
    planes = msi_np.inv_depths(1, 100, P)
    eye = np.eye(4, dtype=F32)[None]
    for b in frames_to_check:
        out, _ = msi_np.infer_msi(src[b:b + 1], ref[b:b + 1], synth.identity_poses(1), synth.identity_poses(1),
                                  synth.intrinsics(1), P, planes, wts, ngf=ngf)
        want = msi_np.msi_render_equirect_view(out["rgba_layers"], eye, tp[b:b + 1], planes)
        wdep = msi_np.msi_render_equirect_depth(out["rgba_layers"], eye, tp[b:b + 1], planes)
        e_rgba = np.abs(pipe.rgba[b:b + 1].cpu().numpy() - out["rgba_layers"]).max()
        e_rgb = np.abs(pipe.out["rgb"][b:b + 1].cpu().numpy() - want).max()
        e_dep = np.abs(pipe.out["depth"][b:b + 1].cpu().numpy() - wdep).max()
        assert e_rgba < TOL and e_rgb < TOL and e_dep < TOL, (b, e_rgba, e_rgb, e_dep)
    return pipe


def test_config3_64_spheres_batch8():
    """BASELINE configs[2]: 640x320 ERP, 64-sphere MSI, batch 8 on one GPU (deep-layer composite
    stress): first and last frame against the oracle, all frames finite and distinct."""
    pipe = _pipeline_vs_oracle(320, 640, 64, 8, frames_to_check=(0, 7))
    rgb = pipe.out["rgb"]
    assert torch.isfinite(rgb).all()
    assert not torch.equal(rgb[0], rgb[1])


def test_config4_high_res_1280x640():
    """BASELINE configs[3] shape: 1280x640 ERP, 32 spheres (one of the 4 frames a rank owns)."""
    _pipeline_vs_oracle(640, 1280, 32, 2, frames_to_check=(1,))


def test_tcgen05_cluster_multicast_path(monkeypatch):
    """MSI_CONV_CLUSTER=2: CTA pairs multicast the W tile; odd tile counts exercise the masked
    surplus CTA.  Must equal the un-clustered kernel bit for bit (same MMAs, same order)."""
    H, W, P, ngf, B = 24, 72, 32, 64, 3
    rng = np.random.default_rng(11)
    x = rng.uniform(-1, 1, (B, H, W, 6 * P)).astype(F32)
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    monkeypatch.setenv("MSI_CONV_HALO", "0")  # both runs on the per-tap kernel (the halo kernel does not cluster)
    monkeypatch.setenv("MSI_CONV_CLUSTER", "1")
    a = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, precision="fp16x3").forward(_t(x))
    monkeypatch.setenv("MSI_CONV_CLUSTER", "2")   # (an fp16x3 / fp16 experiment: the per-tap fp8x kernel has no cluster form)
    b = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, precision="fp16x3").forward(_t(x))
    assert torch.equal(a, b)


@pytest.mark.parametrize("precision", ["fp16x3", "fp16_fp8x"])
@pytest.mark.parametrize("H,W,B", [(24, 72, 3), (40, 80, 1), (64, 128, 2), (8, 16, 1), (56, 104, 1), (16, 264, 2)])
def test_halo_kernel_matches_per_tap_kernel(monkeypatch, H, W, B, precision):
    """The halo-reuse kernel (taps read from one smem halo tile through row-shifted descriptors,
    16x8 / 8x16 pixel tiles, K-block-major weights) and the per-tap kernel compute the same
    products in a different order: outputs agree to float32 accumulation noise, and both are within
    the path's 1e-3 bar of the oracle (test_net_* above run on the default = halo kernel).  Ragged
    sizes exercise zero-filled halos and masked tile rows in both tile orientations; 8x16 shrinks the
    deepest layers to 1x2 pixels (a tile that is almost entirely padding), 16x264 has more tile columns
    than rows at every level."""
    P, ngf = 32, 64
    rng = np.random.default_rng(12)
    x = rng.uniform(-1, 1, (B, H, W, 6 * P)).astype(F32)
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    monkeypatch.setenv("MSI_CONV_HALO", "0")
    a = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, precision=precision).forward(_t(x))
    monkeypatch.setenv("MSI_CONV_HALO", "1")
    b = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, precision=precision).forward(_t(x))
    err = float((a - b).abs().max())
    # fp16x3: float32 accumulation order.  fp16_fp8x: a last-bit difference also re-draws the e4m3 rounding of the next
    # layer's cross-term operands, i.e. part of that precision's own 1.6e-4 (see test_stride2_halo_matches_per_tap_kernel)
    assert err < (5e-5 if precision == "fp16x3" else 2e-4), err


def test_streaming_submit_collect_matches_step():
    """MSIPipeline.submit / collect (H2D, compute, D2H on three streams, two batches in flight)
    returns, in order, exactly what the synchronous device-resident step computes."""
    H, W, P, ngf = 32, 64, 32, 64
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    pipe = MSIPipeline(wts, H, W, P, ngf, batch=1, device=DEV)
    frames = [synth.ods_pair(1, H, W, seed=100 + i) for i in range(5)]
    want = []
    for ref, src in frames:
        pipe.set_inputs(ref, src)
        pipe.step()
        torch.cuda.synchronize()
        want.append((pipe.out["rgb_u8"].cpu().clone(), pipe.out["depth_u8"].cpu().clone()))
    got = []
    for i, (ref, src) in enumerate(frames):
        pipe.submit(torch.from_numpy(ref).pin_memory(), torch.from_numpy(src))  # pinned and pageable inputs
        if i >= 1:
            r = pipe.collect()
            got.append((r[0].clone(), r[1].clone()))
    r = pipe.collect()
    got.append((r[0].clone(), r[1].clone()))
    assert len(got) == 5
    for (a, b), (c, d) in zip(want, got):
        assert torch.equal(a, c) and torch.equal(b, d)
    from matryodshka_b200._lib import MsiError
    with pytest.raises(MsiError):
        pipe.collect()


@pytest.mark.parametrize("H,W,B", [(32, 64, 1), (24, 72, 2), (64, 128, 1)])
def test_train_net_wrap_pad_matches_oracle(H, W, B):
    """nets.msi_train_net (nets.py:387-469, MSI_NET_WRAP): no coord channel, circular-x / zero-y
    wrap_pad + VALID convs, stride-2 convs padded (1, 1), deconvs on wrap_pad(x, 2, 2) cropped [5:-5].
    Every layer's activation and the head against the oracle restatement; an input whose left and
    right image borders differ strongly makes a zero-padded (SAME) evaluation fail this test."""
    P, ngf = 32, 64
    rng = np.random.default_rng(31)
    x = rng.uniform(-1, 1, (B, H, W, 6 * P)).astype(F32)
    x[:, :, :2, :] += 1.5   # the wrap columns matter
    x[:, :, -2:, :] -= 1.5
    wts = synth.net_weights(6 * P, 2 * P, ngf, coord=False)
    eng = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, variant="wrap")
    pred = eng.forward(_t(x)).cpu().numpy()
    with torch.no_grad():
        want, feats = net_torch.msi_train_net(torch.from_numpy(x), 2 * P, wts, ngf=ngf, return_feats=True)
    worst = {s: float(np.abs(eng.read_activation(s, B).cpu().numpy() - f.numpy()).max()) for s, f in feats.items()}
    err = float(np.abs(pred - want.numpy()).max())
    assert err < TOL, (err, worst)
    assert max(worst.values()) < TOL, worst


def test_infer_msi_with_train_net(monkeypatch):
    """MSI.infer_msi with FLAGS.coord_net = False (msi.py:120-127 picks nets.msi_train_net) against the oracle."""
    H, W, P, ngf = 32, 64, 32, 64
    ref, src = synth.ods_pair(1, H, W)
    wts = synth.net_weights(6 * P, 2 * P, ngf, coord=False)
    m = MSI(weights=wts, config=MSIConfig(coord_net=False), device=DEV)
    planes = m.inv_depths(1, 100, P)
    out, net_input = m.infer_msi(_t(src), _t(ref), None, None, _t(synth.identity_poses(1)), _t(synth.identity_poses(1)),
                                 _t(synth.intrinsics(1)), 'blend_psv', P, planes, ngf=ngf)
    x = msi_np.format_network_input(ref * 2 - 1, src * 2 - 1, synth.identity_poses(1), synth.identity_poses(1), planes,
                                    synth.intrinsics(1))
    with torch.no_grad():
        pred = net_torch.msi_train_net(torch.from_numpy(x), 2 * P, wts, ngf=ngf).numpy()
    want = msi_np.rgba_layers_blend_psv(pred, x, P) if hasattr(msi_np, "rgba_layers_blend_psv") else None
    assert float(np.abs(net_input.cpu().numpy() - x).max()) < TOL
    if want is not None:
        assert float(np.abs(out['rgba_layers'].cpu().numpy() - want).max()) < TOL
    else:
        bw = (pred[..., :P] + 1) / 2
        al = (pred[..., P:] + 1) / 2
        fg = x[..., :3 * P].reshape(1, H, W, P, 3)
        bg = x[..., 3 * P:].reshape(1, H, W, P, 3)
        rgb = bw[..., None] * fg + (1 - bw[..., None]) * bg
        want = np.concatenate([rgb, al[..., None]], axis=-1)
        assert float(np.abs(out['rgba_layers'].cpu().numpy() - want).max()) < TOL


@pytest.mark.parametrize("which", ["blend_bg", "blend_bg_psv", "alpha_only"])
def test_infer_msi_other_color_schemes(which):
    """which_color_pred != blend_psv (msi.py:166-268): net heads with 2L+3 / 3L+3 / L outputs (padded to a
    multiple of 64 for the tensor-core kernel) and the matching RGBA assembly, against the oracle."""
    H, W, P, ngf = 32, 64, 32, 64
    ref, src = synth.ods_pair(1, H, W, seed=77)
    n_out = ops.color_pred_channels(which, P)
    wts = synth.net_weights(6 * P, n_out, ngf)
    m = MSI(weights=wts, device=DEV)
    planes = m.inv_depths(1, 100, P)
    eye = synth.identity_poses(1)
    out, net_input = m.infer_msi(_t(src), _t(ref), None, None, _t(eye), _t(eye), _t(synth.intrinsics(1)), which, P,
                                 planes, extra_outputs='blend_weights_alphas', ngf=ngf)
    want, _ = msi_np.infer_msi(src, ref, eye, eye, synth.intrinsics(1), P, planes, wts, extra_outputs='blend_weights_alphas',
                               ngf=ngf, which_color_pred=which)
    assert float(np.abs(out['rgba_layers'].cpu().numpy() - want['rgba_layers']).max()) < TOL
    assert float(np.abs(out['alphas'].cpu().numpy() - want['alphas']).max()) < TOL
    if 'blend' in which:
        assert float(np.abs(out['blend_weights'].cpu().numpy() - want['blend_weights']).max()) < TOL
    else:
        assert 'blend_weights' not in out
    assert ('bg_blend_weights' in out) == (which == 'blend_bg_psv')
    # the assembly itself is bit-exact given the same prediction (separate mul / add, no FMA)
    pred = torch.rand(1, H, W, n_out, device=DEV) * 2 - 1
    rgba, bw, al, bgw = ops.rgba_assemble_ex(pred, net_input, which, P, want_weights=True)
    w_rgba, w_bw, w_al, w_bgw = msi_np.assemble_rgba_ex(pred.cpu().numpy(), net_input.cpu().numpy(), P, which)
    assert np.array_equal(rgba.cpu().numpy(), w_rgba)
    assert np.array_equal(al.cpu().numpy(), w_al)


def test_train_net_encoder_circular_shift_equivariance_on_gpu():
    """Size-independent property of the wrap-pad net (no oracle needed): rolling the panorama by 8 columns
    rolls every encoder activation by 8 / 4 / 2 / 1 columns (circular padding along the width), at a size
    where the image border falls inside tiles, between tiles and at tile edges."""
    H, W, P, ngf, B = 64, 136, 32, 64, 1
    rng = np.random.default_rng(41)
    x = rng.uniform(-1, 1, (B, H, W, 6 * P)).astype(F32)
    wts = synth.net_weights(6 * P, 2 * P, ngf, coord=False)
    eng = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, variant="wrap")
    eng.forward(_t(x))
    a = {s: eng.read_activation(s, B).clone() for s in ("conv1_1", "conv1_2", "conv2_2", "conv3_3", "conv4_3")}
    eng.forward(_t(np.roll(x, 8, axis=2)))
    for s, shift in [("conv1_1", 8), ("conv1_2", 4), ("conv2_2", 2), ("conv3_3", 1), ("conv4_3", 1)]:
        b = eng.read_activation(s, B)
        err = float((torch.roll(a[s], shift, dims=2) - b).abs().max())
        assert err < 1e-4, (s, err)


@pytest.mark.parametrize("coord_net,which,P", [(True, "blend_psv", 16), (False, "blend_psv", 32), (True, "blend_bg", 32),
                                               (False, "blend_bg_psv", 32), (True, "alpha_only", 32)])
def test_pipeline_variants_equal_the_mirror_api(coord_net, which, P):
    """The graph-captured MSIPipeline with msi_train_net (coord_net = False), the other colour schemes and a head
    that is padded for the tensor-core kernel (2 * 16 = 32 -> 64 channels) agrees with MSI.infer_msi +
    msi_render_equirect to rounding (same kernels; the pipeline reads the padded prediction in place and assembles
    from the fp16 hi + lo PSV operand, ~22 bits, where the mirror API assembles from the float32 PSV)."""
    H, W, ngf = 32, 64, 64
    ref, src = synth.ods_pair(1, H, W, seed=31)
    tp = synth.target_positions(1, 31)
    wts = synth.net_weights(6 * P, ops.color_pred_channels(which, P), ngf, coord=coord_net)
    pipe = MSIPipeline(wts, H, W, P, ngf, batch=1, device=DEV, coord_net=coord_net, which_color_pred=which)
    pipe.set_inputs(ref, src, tgt_pos=tp)
    pipe.step()
    pipe.step()   # second step = graph replay
    torch.cuda.synchronize()
    m = MSI(weights=wts, config=MSIConfig(coord_net=coord_net), device=DEV)
    planes = m.inv_depths(1, 100, P)
    eye = synth.identity_poses(1)
    out, _ = m.infer_msi(_t(src), _t(ref), None, None, eye, eye, synth.intrinsics(1), which, P, planes, ngf=ngf)
    assert float((pipe.rgba - out["rgba_layers"]).abs().max()) < 1e-5
    res = m.msi_render_equirect(out["rgba_layers"], np.eye(4, dtype=F32)[None], tp, planes)
    assert float((pipe.out["rgb"] - res["rgb"]).abs().max()) < 1e-5
    assert int((pipe.out["depth_u8"].int() - res["depth_u8"].int()).abs().max()) <= 1
    want, _ = msi_np.infer_msi(src, ref, eye, eye, synth.intrinsics(1), P, planes, wts, ngf=ngf, coord_net=coord_net,
                               which_color_pred=which)
    assert float(np.abs(pipe.rgba.cpu().numpy() - want["rgba_layers"]).max()) < TOL


@pytest.mark.parametrize("precision", ["fp16x3", "fp16_fp8x"])
@pytest.mark.parametrize("H,W,B", [(64, 128, 1), (24, 72, 3), (8, 16, 1)])
def test_cta_pair_kernel_matches_single_cta_kernel(monkeypatch, H, W, B, precision):
    """Default (MSI_CONV_PAIR unset / 1) vs MSI_CONV_PAIR=0: the halo kernel as CTA pairs (tcgen05 cta_group::2, each CTA holds half of the weight
    rows) computes the same products as the single-CTA halo kernel; per-layer raw outputs agree to float32
    accumulation noise.  (24, 72, 3) has odd tile counts: the surplus CTA of the last pair is masked; at (8, 16, 1)
    the deepest layers have a single tile and fall back to one CTA per unit."""
    P, ngf = 32, 64
    rng = np.random.default_rng(13)
    x = rng.uniform(-1, 1, (B, H, W, 6 * P)).astype(F32)
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    monkeypatch.setenv("MSI_CONV_PAIR", "0")
    ea = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, precision=precision)
    a = ea.forward(_t(x))
    monkeypatch.setenv("MSI_CONV_PAIR", "1")
    eb = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, precision=precision)
    b = eb.forward(_t(x))
    torch.cuda.synchronize()
    errs = {}
    for scope in ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2",
                  "conv4_3", "conv6_1", "conv6_2", "conv6_3", "conv7_1", "conv7_2", "conv8_1", "conv8_2"]:
        errs[scope] = float((ea.read_raw(scope, B) - eb.read_raw(scope, B)).abs().max())
    errs["pred"] = float((a - b).abs().max())
    # conv1_1 sees identical inputs: summation order alone.  Downstream of it fp16_fp8x also re-draws the e4m3 rounding of
    # the cross-term operands (part of that precision's own 1.6e-4 noise), fp16x3 stays at float32 accumulation noise.
    assert errs["conv1_1"] < 2e-5, errs
    assert max(errs.values()) < (5e-5 if precision == "fp16x3" else 2e-4), errs
    if H * W >= 4 * 128:   # conv1_1 has several pixel tiles per frame, so it runs as pairs (different summation order)
        assert errs["conv1_1"] > 0.0, "the pair kernel was not selected (outputs are bit-identical)"


@pytest.mark.parametrize("H,W,B,variant,precision,pair", [
    (64, 128, 1, "coord", "fp16_fp8x", "1"), (24, 72, 3, "coord", "fp16_fp8x", "1"), (8, 16, 1, "coord", "fp16_fp8x", "1"),
    (56, 104, 1, "coord", "fp16x3", "1"), (64, 128, 2, "coord", "fp16_fp8x", "0"), (64, 128, 1, "wrap", "fp16_fp8x", "1"),
    (24, 72, 2, "wrap", "fp16x3", "1")])
def test_stride2_halo_matches_per_tap_kernel(monkeypatch, H, W, B, variant, precision, pair):
    """The stride-2 convs (conv1_2, conv2_2, conv3_3) on the halo kernel: the input splits into four parity planes
    (TMA element stride 2), each plane's halo tile serves its taps (4 + 2 + 2 + 1) through row-shifted descriptors.
    Opt-in (MSI_CONV_HALO_S2=1; measured no faster than the per-tap kernel, which stays the default for these three
    layers): same products, taps summed in a different order, so every layer's raw output agrees to float32
    accumulation noise.  Covers SAME padding (coord net: offsets 0..2) and the
    wrap net's (1, 1) padding (offsets -1..1: negative plane offsets), ragged tiles, pairs and single CTAs."""
    P, ngf = 32, 64
    rng = np.random.default_rng(17)
    x = rng.uniform(-1, 1, (B, H, W, 6 * P)).astype(F32)
    wts = synth.net_weights(6 * P, 2 * P, ngf, coord=(variant == "coord"))
    monkeypatch.setenv("MSI_CONV_PAIR", pair)
    monkeypatch.setenv("MSI_CONV_HALO_S2", "0")
    ea = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, variant=variant, precision=precision)
    a = ea.forward(_t(x))
    monkeypatch.setenv("MSI_CONV_HALO_S2", "1")
    eb = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, variant=variant, precision=precision)
    b = eb.forward(_t(x))
    torch.cuda.synchronize()
    errs = {s: float((ea.read_raw(s, B) - eb.read_raw(s, B)).abs().max()) for s in ("conv1_2", "conv2_2", "conv3_3")}
    errs["pred"] = float((a - b).abs().max())
    print("stride-2 halo vs per-tap:", errs)
    # conv1_2 sees bit-identical inputs in both runs (conv1_1 is untouched): its difference is the summation order alone.
    # Downstream, fp16_fp8x re-rounds the activations' e4m3 copies, so a last-bit difference re-draws part of that
    # precision's own quantisation noise (1.6e-4 against the oracle): the bound there is a fraction of it, not float noise.
    assert errs["conv1_2"] < 2e-5, errs
    assert max(errs.values()) < (5e-5 if precision == "fp16x3" else 2e-4), errs
    assert not bool(torch.isnan(b).any())
    if H * W >= 4 * 128:
        assert errs["conv1_2"] > 0.0, "the stride-2 halo form was not selected (outputs are bit-identical)"


@pytest.mark.parametrize("H,W,P,B,coord", [(64, 128, 32, 1, True), (24, 72, 32, 2, True), (32, 64, 64, 1, True),
                                           (32, 64, 32, 1, False)])
def test_head_fused_rgba_equals_forward_plus_assemble_bit_for_bit(H, W, P, B, coord):
    """msi_net_forward_rgba (the head's epilogue assembles the RGBA layers, msi.py:130-147) against msi_net_forward +
    msi_rgba_assemble on the same fp16 hi / lo PSV operand: same arithmetic in the same order, so the SAME BITS.
    Cases: L = 32 (two pixels per warp iteration), a batch with ragged tiles (72 = 64 + 8 columns), L = 64 (N tile
    128), and the wrap-padded input of msi_train_net (the PSV rows are stored 2 pixels wider on both sides)."""
    ngf = 64
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, (B, H, W, 6 * P)).astype(F32)
    wts = synth.net_weights(6 * P, 2 * P, ngf, coord=coord)
    eng = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV, max_batch=B, variant="coord" if coord else "wrap")
    assert eng.can_fuse_rgba
    pred = eng.forward(_t(x)).clone()
    cs = eng.in_c_stride
    # the operand the net itself consumed: hi / lo split of the scaled input (dense copy; for the wrap variant the
    # engine's own buffer is wrap-padded, so split the input again the same way)
    xs = _t(x) * 16.0
    hi = torch.zeros((B, H, W, cs), dtype=torch.float16, device=DEV)
    lo = torch.zeros_like(hi)
    hi[..., :6 * P] = xs.to(torch.float16)
    lo[..., :6 * P] = (xs - hi[..., :6 * P].float()).to(torch.float16)
    want, _, _ = ops.rgba_assemble(pred, None, hi_lo=(hi, lo), c_stride=cs)
    got = eng.forward_rgba(_t(x))
    torch.cuda.synchronize()
    assert torch.equal(got, want), float((got - want).abs().max())
    # and against the oracle's assembly of the oracle's prediction
    with torch.no_grad():
        fn = net_torch.msi_coord_train_net if coord else net_torch.msi_train_net
        opred = fn(torch.from_numpy(x), 2 * P, wts, ngf=ngf).numpy()
    orgba, _, _ = msi_np.assemble_rgba(opred, x, P)
    assert float(np.abs(got.cpu().numpy() - orgba).max()) < TOL


def test_pipeline_fused_head_equals_unfused_pipeline():
    """MSIPipeline with the fused head (default) and with the separate K4 launch: identical frames."""
    H, W, P, ngf = 64, 128, 32, 64
    ref, src = synth.ods_pair(1, H, W, seed=77)
    tp = synth.target_positions(1, 77)
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    outs = []
    for fuse in (True, False):
        pipe = MSIPipeline(wts, H, W, P, ngf, batch=1, device=DEV, fuse_rgba=fuse)
        assert pipe.fused_rgba == fuse
        assert [n for n, _ in pipe._stages()] == (["psv_build", "net", "render_composite"] if fuse else
                                                  ["psv_build", "net", "rgba_assemble", "render_composite"])
        pipe.set_inputs(ref, src, tgt_pos=tp)
        pipe.step()
        pipe.step()
        torch.cuda.synchronize()
        outs.append((pipe.rgba.clone(), pipe.out["rgb"].clone(), pipe.out["rgb_u8"].clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_pipeline_static_rig_table_equals_per_frame_chain_and_follows_the_rig():
    """static_rig=True (cached coordinate table, default) and False (the chain evaluated every frame) render the same
    bits; changing the baseline through set_inputs rebuilds the table (and re-captures the graph)."""
    H, W, P, ngf = 32, 64, 32, 64
    ref, src = synth.ods_pair(1, H, W, seed=78)
    tp = synth.target_positions(1, 78)
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    a = MSIPipeline(wts, H, W, P, ngf, batch=1, device=DEV, static_rig=True)
    b = MSIPipeline(wts, H, W, P, ngf, batch=1, device=DEV, static_rig=False)
    for base in (0.032, 0.05):
        for pipe in (a, b):
            pipe.set_inputs(ref, src, tgt_pos=tp, baselines=[base])
            pipe.step()
            pipe.step()
        torch.cuda.synchronize()
        assert torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo)
        assert torch.equal(a.out["rgb"], b.out["rgb"]) and torch.equal(a.out["depth_u8"], b.out["depth_u8"])


def test_submit_pageable_batches_back_to_back():
    """`depth` pageable (un-pinned) batches submitted back to back, no collect in between: every batch is staged
    through its own pinned buffer pair, so the asynchronous H2D copy of batch i never reads pixels of batch i + 1.
    The pattern s0 s1 c0 c1 s2 s3 c2 c3 is repeated at a frame size whose copy takes long enough to expose a race."""
    H, W, P, ngf = 64, 128, 32, 64
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    pipe = MSIPipeline(wts, H, W, P, ngf, batch=1, device=DEV)
    frames = [synth.ods_pair(1, H, W, seed=300 + i) for i in range(6)]
    want = []
    for ref, src in frames:
        pipe.set_inputs(ref, src)
        pipe.step()
        torch.cuda.synchronize()
        want.append(pipe.out["rgb_u8"].cpu().clone())
    got = []
    for i in range(0, 6, 2):
        pipe.submit(frames[i][0].copy(), frames[i][1].copy())
        pipe.submit(frames[i + 1][0].copy(), frames[i + 1][1].copy())
        got.append(pipe.collect()[0].clone())
        got.append(pipe.collect()[0].clone())
    for a, b in zip(want, got):
        assert torch.equal(a, b)
    with pytest.raises(Exception):
        u8 = MSIPipeline(wts, H, W, P, ngf, batch=1, device=DEV, img_dtype=torch.uint8)
        u8.set_inputs(frames[0][0], frames[0][1])   # float images into a uint8 pipeline

"""High-res plane-streamed re-render (reference test.py:284-394) against its oracle restatement."""
import numpy as np
import pytest
import torch

from oracle import highres_np, msi_np
from matryodshka_b200 import synth
from matryodshka_b200.highres import deprocess_high_res, high_res_rerender

pytestmark = pytest.mark.gpu
F32 = np.float32
DEV = "cuda"


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


@pytest.mark.parametrize("lh,lw,Hh,Wh,P", [(8, 16, 32, 64, 4), (16, 32, 100, 200, 3)])
def test_high_res_rerender_matches_oracle(lh, lw, Hh, Wh, P):
    ref, src = synth.ods_pair(1, Hh, Wh)
    rng = np.random.default_rng(2)
    # smooth low-res weights / alphas, as a net would produce
    bw = synth.band_limited_images(P, lh, lw, 3)[..., 0].transpose(1, 2, 0)[None].astype(F32)
    al = synth.band_limited_images(P, lh, lw, 4)[..., 1].transpose(1, 2, 0)[None].astype(F32)
    planes = msi_np.inv_depths(1, 100, P)
    tp = np.array([[0.03, -0.02, 0.04]], F32)
    eye = synth.identity_poses(1)
    want_rgb, want_dep = highres_np.high_res_rerender(ref, src, bw, al, eye, eye, synth.intrinsics(1), tp, planes)
    got_rgb, got_dep = high_res_rerender(_t(ref), _t(src), _t(bw), _t(al), eye, eye, synth.intrinsics(1), tp, planes)
    assert np.abs(got_rgb.cpu().numpy() - want_rgb).max() < 1e-3
    assert np.abs(got_dep.cpu().numpy() - want_dep).max() < 1e-3
    u8, d8 = deprocess_high_res(got_rgb, got_dep)
    w8, wd8 = highres_np.deprocess_high_res(want_rgb, want_dep)
    assert np.abs(u8.cpu().numpy().astype(int) - w8.astype(int)).max() <= 1
    assert np.abs(d8.cpu().numpy().astype(int) - wd8.astype(int)).max() <= 1


def test_upsample_oracle_matches_torch_align_corners():
    """The TF-1.14 align_corners bilinear restated in the oracle equals torch's align_corners=True."""
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1, (1, 7, 9, 2)).astype(F32)
    a = highres_np.resize_bilinear_align_corners(x, 20, 31)
    b = torch.nn.functional.interpolate(torch.from_numpy(x).permute(0, 3, 1, 2), size=(20, 31), mode="bilinear",
                                        align_corners=True).permute(0, 2, 3, 1).numpy()
    assert np.abs(a - b).max() < 1e-5


def test_full_size_high_res_runs():
    """4096x2048, 32 planes from 640x320 weights: finite output in range (no oracle at this size)."""
    Hh, Wh, lh, lw, P = 2048, 4096, 320, 640, 32
    g = torch.Generator(device="cpu").manual_seed(0)
    ref = torch.rand((1, Hh, Wh, 3), generator=g).to(DEV)
    src = torch.roll(ref, 5, dims=2)
    bw = torch.rand((1, lh, lw, P), generator=g).to(DEV)
    al = torch.rand((1, lh, lw, P), generator=g).to(DEV)
    planes = msi_np.inv_depths(1, 100, P)
    eye = synth.identity_poses(1)
    rgb, dep = high_res_rerender(ref, src, bw, al, eye, eye, synth.intrinsics(1), np.array([[0.02, 0.0, -0.03]], F32),
                                 planes)
    torch.cuda.synchronize()
    assert torch.isfinite(rgb).all() and rgb.abs().max() <= 1.0001 and dep.min() >= 0 and dep.max() < 1

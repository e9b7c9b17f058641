"""Generates tests/golden/msi_small.npz with the CPU oracle (run from the repo
root: ``python -m tests.golden.make_golden``).

These vectors come from the oracle restatement: they guard against drift of the
oracle and give the GPU tests a fixture that travels to the GPU box.  The vectors
that pin the oracle to the reference are tests/golden/reference_run.npz (the
reference's own source executed by oracle/refrun/run_reference.py); on the same
inputs the two fixtures agree bit for bit on the PSV and to 1.3e-6 on the RGBA layers.
"""
import os

import numpy as np

from oracle import msi_np
from matryodshka_b200 import synth

H, W, P, NGF, SEED = 16, 32, 4, 8, 8964


def inputs_small():
    ref, src = synth.ods_pair(1, H, W, SEED)
    return dict(
        ref=ref, src=src,
        ref_pose=synth.identity_poses(1), src_pose=synth.identity_poses(1),
        intrinsics=synth.intrinsics(1),
        tgt_pos=np.array([[0.03, -0.02, 0.04]], np.float32),
        planes=np.array(msi_np.inv_depths(1, 100, P), np.float64),
        weights=synth.net_weights(6 * P, 2 * P, NGF, SEED),
    )


def run_small():
    i = inputs_small()
    planes = list(i["planes"])
    out, net_input = msi_np.infer_msi(i["src"], i["ref"], i["ref_pose"], i["src_pose"], i["intrinsics"],
                                      P, planes, i["weights"], "blend_weights_alphas_psv", ngf=NGF)
    pred = np.concatenate([out["blend_weights"], out["alphas"]], -1) * 2 - 1
    eye = np.eye(4, dtype=np.float32)[None]
    render = msi_np.msi_render_equirect_view(out["rgba_layers"], eye, i["tgt_pos"], planes)
    depth = msi_np.msi_render_equirect_depth(out["rgba_layers"], eye, i["tgt_pos"], planes)
    return dict(psv=net_input, pred=pred.astype(np.float32), rgba_layers=out["rgba_layers"],
                render=render, depth=depth, render_u8=msi_np.deprocess_image(render),
                depth_u8=msi_np.deprocess_depth_image(depth))


if __name__ == "__main__":
    out = run_small()
    path = os.path.join(os.path.dirname(__file__), "msi_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})

"""CPU numerics experiment, per-layer product forms (follows scripts/exp_fp8_cross.py).

Which of the three partial products  x_hi*w_hi + x_hi*w_lo + x_lo*w_hi  does each layer need, and in which format?
Forms per layer:  'x3' fp16x3;  'f8x' both cross terms e4m3 (the shipped MSI_PREC_FP16_FP8X form of the N = 128
layers);  'lo8' x_hi*[w_hi | w_lo] in fp16 and x_lo*w_hi in e4m3 (Cout = 64 candidate);  'nolo' x_hi*[w_hi | w_lo]
only (the activation residual dropped);  '1' one fp16 pass.
    python scripts/exp_layer_modes.py [H W]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from matryodshka_b200 import synth  # noqa: E402
from oracle import msi_np, net_torch as nt  # noqa: E402
from exp_fp8_cross import f16, f8, ACT_SCALE, W_SCALE  # noqa: E402

SA, SW, TA, TW = 2, 3, 9, 3   # the shipped scales (net_internal.cuh)


def product(conv, x, w, form):
    if form == "f32":
        return conv(x, w)
    xs, ws = x * ACT_SCALE, w * W_SCALE
    x_hi, w_hi = f16(xs), f16(ws)
    x_lo, w_lo = xs - x_hi, ws - w_hi
    main = conv(x_hi, w_hi)
    s = 1.0 / (ACT_SCALE * W_SCALE)
    if form == "1":
        return main * s
    if form == "x3":
        return (main + conv(x_hi, f16(w_lo)) + conv(f16(x_lo), w_hi)) * s
    if form == "nolo":
        return (main + conv(x_hi, f16(w_lo))) * s
    a_lo8 = f8(x_lo * 2.0 ** TA, "e4m3")
    w_hi8 = f8(w_hi * 2.0 ** (-TA + TW), "e4m3") * 2.0 ** (-TW)
    if form == "lo8":
        return (main + conv(x_hi, f16(w_lo)) + conv(a_lo8, w_hi8)) * s
    if form == "f8x":
        a_hi8 = f8(x_hi * 2.0 ** (-SA), "e4m3")
        w_lo8 = f8(w_lo * 2.0 ** (SA + SW), "e4m3") * 2.0 ** (-SW)
        return (main + conv(a_hi8, w_lo8) + conv(a_lo8, w_hi8)) * s
    raise ValueError(form)


def run_net(inputs, weights, forms):
    T = lambda n: torch.from_numpy(np.asarray(weights[n])).float()  # noqa: E731

    def cconv(x, scope, stride=1, rate=1):
        w = T(f"net/{scope}/weights")
        xin = nt.add_sph_coords(x)
        cin = x.shape[3]
        y = product(lambda a, b: nt.conv2d_same(a, b, stride=stride, rate=rate), x, w[:, :, :cin, :], forms[scope])
        y = y + nt.conv2d_same(xin[..., cin:], w[:, :, cin:, :], stride=stride, rate=rate)
        return nt.layer_norm_relu(y, T(f"net/{scope}/LayerNorm/gamma"), T(f"net/{scope}/LayerNorm/beta"))

    def deconv(x, scope):
        y = product(nt.conv2d_transpose_same, x, T(f"net/{scope}/weights"), forms[scope])
        return nt.layer_norm_relu(y, T(f"net/{scope}/LayerNorm/gamma"), T(f"net/{scope}/LayerNorm/beta"))

    c11 = cconv(inputs, "conv1_1")
    c12 = cconv(c11, "conv1_2", stride=2)
    c21 = cconv(c12, "conv2_1")
    c22 = cconv(c21, "conv2_2", stride=2)
    c31 = cconv(c22, "conv3_1")
    c32 = cconv(c31, "conv3_2")
    c33 = cconv(c32, "conv3_3", stride=2)
    c41 = cconv(c33, "conv4_1", rate=2)
    c42 = cconv(c41, "conv4_2", rate=2)
    c43 = cconv(c42, "conv4_3", rate=2)
    c61 = deconv(torch.cat([c43, c33], dim=3), "conv6_1")
    c62 = cconv(c61, "conv6_2")
    c63 = cconv(c62, "conv6_3")
    c71 = deconv(torch.cat([c63, c22], dim=3), "conv7_1")
    c72 = cconv(c71, "conv7_2")
    c81 = deconv(torch.cat([c72, c12], dim=3), "conv8_1")
    c82 = cconv(c81, "conv8_2")
    pred = torch.tanh(product(lambda a, bb: nt.conv2d_same(a, bb), c82, T("net/color_pred/weights"), forms["color_pred"])
                      + T("net/color_pred/biases").view(1, 1, 1, -1))
    return pred


SCOPES = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2", "conv4_3",
          "conv6_1", "conv6_2", "conv6_3", "conv7_1", "conv7_2", "conv8_1", "conv8_2", "color_pred"]
NARROW = ("conv1_1", "conv8_1", "conv8_2", "color_pred")   # Cout = 64 layers + the head: fp16x3 in the shipped default


def forms(default="f8x", **over):
    f = {s: ("x3" if s in NARROW else default) for s in SCOPES}
    f.update(over)
    return f


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 128)
    P, ngf = 32, 64
    torch.set_num_threads(os.cpu_count() or 1)
    cases = [
        ("f32", {s: "f32" for s in SCOPES}),
        ("shipped default (f8x on N=128 layers, x3 on Cout=64 + head)", forms()),
        ("+ conv1_1 nolo", forms(conv1_1="nolo")),
        ("+ conv1_1 lo8", forms(conv1_1="lo8")),
        ("+ conv8_1, conv8_2 lo8", forms(conv8_1="lo8", conv8_2="lo8")),
        ("+ conv1_1, conv8_1, conv8_2 lo8", forms(conv1_1="lo8", conv8_1="lo8", conv8_2="lo8")),
        ("+ conv1_1 nolo, conv8_x lo8", forms(conv1_1="nolo", conv8_1="lo8", conv8_2="lo8")),
        ("+ conv8_1 nolo", forms(conv8_1="nolo")),
        ("+ conv8_2 nolo", forms(conv8_2="nolo")),
        ("+ all three + head lo8", forms(conv1_1="lo8", conv8_1="lo8", conv8_2="lo8", color_pred="lo8")),
        ("+ all three f8x", forms(conv1_1="f8x", conv8_1="f8x", conv8_2="f8x")),
    ]
    for seed in (8964, 1234):
        ref, src = synth.ods_pair(1, H, W, seed)
        wts = synth.net_weights(6 * P, 2 * P, ngf, seed)
        planes = msi_np.inv_depths(1, 100, P)
        eye = synth.identity_poses(1)
        x = torch.from_numpy(msi_np.format_network_input(msi_np.preprocess_image(ref), msi_np.preprocess_image(src), eye,
                                                         eye, planes, synth.intrinsics(1)))
        base = None
        with torch.no_grad():
            for name, f in cases:
                pred = run_net(x, wts, f)
                if base is None:
                    base = pred
                    continue
                print(f"seed {seed} {H}x{W}  {name:62s} max|pred - f32| = {float((pred - base).abs().max()):.3e}", flush=True)


if __name__ == "__main__":
    main()

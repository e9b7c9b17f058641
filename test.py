#!/usr/bin/env python
"""Inference driver with the reference's flags and output tree (reference test.py:36-83, :87-281).

    python test.py --cameras_glob 'glob/test/reg/*.txt' --image_dir /data/360TestData \
                   --input_type ODS --experiment_name ods-wotemp-elpips-coord --coord_net

For every camera line ``scene img_ref img_src img_tgt baseline tx ty tz``
(datasets.py:413-424) it loads ``<image_dir>/<scene>_pos<img>.jpeg`` (area-resized to
--height x --width, datasets.py:507-518), infers the MSI, renders the target view + depth and writes
``<output_root>/<experiment_name>/<scene>_<ids>/{tgt_image_*,output_tgt_*,output_depth_*,src_image_*,
ref_image_*,msi_alpha_%02d,msi_rgb_%02d,blend_weight_%03d}.png``, ``blend_weights.npy``,
``alphas.npy`` and ``step.txt`` -- the files the reference's eval.py reads (eval.py:132-136).

Differences from the reference (it needs a TF-1.14 session, this needs a B200):
* weights come from the TensorFlow checkpoint ``tf.train.latest_checkpoint(<checkpoint_dir>/<experiment_name>)``
  names (test.py:192-202; parsed by ``matryodshka_b200.tf_checkpoint``, no TensorFlow needed), or from
  ``<checkpoint_dir>/<experiment_name>/weights.npz`` (arrays keyed by the TF variable names).  ``--random_init`` uses seeded
  random weights, ``--synthetic N`` fabricates N synthetic ODS triples (no dataset needed);
* ``--test_type high_res`` / ``high_res_only`` (test.py:284-394) re-render at --hres_height x
  --hres_width from the saved blend_weights.npy / alphas.npy, plane by plane on the GPU, and write
  output_hrestgt_*.png / output_hresdepth_*.png; with --synthetic the high-res images are fabricated too;
* ``on_video`` only changes the output directory names (test.py:209-217), as in the reference; the
  psp / ODS re-renders of ``--dry_run_inference`` and the GCN path are not built into this driver.
"""
from __future__ import annotations

import argparse
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def str2bool(v):
    return str(v).lower() in ("1", "true", "t", "yes", "y", "")


def parse_flags(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    a = ap.add_argument
    # i/o (test.py:39-47)
    a("--cameras_glob", default="glob/test/regular/*.txt")
    a("--image_dir", default="/path/to/test_640x320")
    a("--hres_image_dir", default="/path/to/test_4096x2048")
    a("--output_root", default="./test")
    a("--checkpoint_dir", default="checkpoints")
    a("--experiment_name", default="")
    a("--shuffle_seq_length", type=int, default=3)
    # model (test.py:49-62)
    a("--operation", default="train")
    a("--input_type", default="ODS")
    a("--coord_net", nargs="?", const=True, default=False, type=str2bool)
    a("--transform_inverse_reg", nargs="?", const=True, default=False, type=str2bool)
    a("--jitter", nargs="?", const=True, default=False, type=str2bool)
    a("--which_color_pred", default="blend_psv")
    a("--ngf", type=int, default=64)
    a("--min_depth", type=float, default=1)
    a("--max_depth", type=float, default=100)
    a("--num_psv_planes", type=int, default=32)
    a("--num_msi_planes", type=int, default=32)
    a("--random_seed", type=int, default=8964)
    a("--net_only", nargs="?", const=True, default=False, type=str2bool)
    a("--smoothed", nargs="?", const=True, default=False, type=str2bool)
    a("--supervision", default="tgt")
    a("--rot_factor", type=float, default=1.0)
    a("--tr_factor", type=float, default=1.0)
    # test script specific (test.py:75-83)
    a("--test_type", default="")
    a("--prefix", default="")
    a("--test_outputs", default="rgba_layers_src_image_ref_image_tgt_image_blend_weights_alphas")
    a("--num_runs", type=int, default=-1)
    a("--gcn", nargs="?", const=True, default=False, type=str2bool)
    a("--subdiv", type=int, default=7)
    # loader.py:30-42
    a("--height", type=int, default=320)
    a("--width", type=int, default=640)
    a("--batch_size", type=int, default=1)
    a("--hres_height", type=int, default=2048)
    a("--hres_width", type=int, default=4096)
    # ours
    a("--random_init", action="store_true", help="seeded random weights instead of weights.npz")
    a("--synthetic", type=int, default=0, help="fabricate N synthetic ODS triples instead of reading a dataset")
    a("--conv_impl", default="tcgen05", choices=["tcgen05", "simt"])
    a("--device", default="cuda:0")
    return ap.parse_args(argv)


def write_image(filename, image):
    """utils.py:76-81: clip to [0, 255], cast to uint8, save."""
    from PIL import Image
    byte_image = np.clip(image, 0, 255).astype("uint8")
    if byte_image.ndim == 3 and byte_image.shape[2] == 1:
        byte_image = byte_image[..., 0]
    Image.fromarray(byte_image).save(filename)


def load_image(path, height, width):
    """JPEG -> float32 [H, W, 3] in [0, 1], area-resized (datasets.py:507-518)."""
    from PIL import Image
    img = Image.open(path).convert("RGB")
    if img.size != (width, height):
        img = img.resize((width, height), Image.BOX)
    return np.asarray(img, dtype=np.float32) / 255.0


def read_camera_lines(pattern):
    """Camera files: ``scene id0 id1 id2 baseline tx ty tz`` per line (datasets.py:413-424)."""
    seqs = []
    for f in sorted(glob.glob(pattern)):
        for line in open(f).read().split("\n"):
            t = line.split()
            if len(t) < 8:
                continue
            seqs.append(dict(scene_id=t[0], image_id=t[1:4], baseline=float(t[4]),
                             tgt_pos=np.array([float(x) for x in t[5:8]], np.float32)))
    return seqs


def make_synthetic_dataset(root, n, height, width, seed, sub="images"):
    """Writes n synthetic (ref, src, tgt) JPEG triples + one camera file; returns (glob, image_dir)."""
    from PIL import Image
    from matryodshka_b200 import synth
    img_dir = os.path.join(root, sub)
    cam_dir = os.path.join(root, "glob")
    os.makedirs(img_dir, exist_ok=True)
    os.makedirs(cam_dir, exist_ok=True)
    lines = []
    rng = np.random.default_rng(seed)
    for i in range(n):
        ref, src = synth.ods_pair(1, height, width, seed + 10 * i)
        tgt = synth.band_limited_images(1, height, width, seed + 10 * i + 5)
        for k, im in enumerate((ref[0], src[0], tgt[0])):
            Image.fromarray((im * 255).astype(np.uint8)).save(os.path.join(img_dir, f"synth_pos{3 * i + k:03d}.jpeg"),
                                                               quality=95)
        tp = rng.uniform(-0.05, 0.05, 3)
        lines.append(f"synth {3 * i:03d} {3 * i + 1:03d} {3 * i + 2:03d} 0.032 {tp[0]:.5f} {tp[1]:.5f} {tp[2]:.5f}")
    with open(os.path.join(cam_dir, "synth.txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return os.path.join(cam_dir, "*.txt"), img_dir


def scene_dirname(flags, s):
    """test.py:208-217 / :350-360: [video_[<prefix>_]]<scene>_<src id><ref id><tgt id>."""
    name = s["scene_id"]
    if "on_video" in flags.test_type:
        name = "video_" + (flags.prefix + "_" if flags.prefix != "" else "") + name
    return name + "_%s%s%s" % tuple(s["image_id"])


def load_weights(flags):
    from matryodshka_b200 import synth
    from matryodshka_b200 import tf_checkpoint
    from matryodshka_b200.ops import color_pred_channels
    if flags.random_init:
        n_out = color_pred_channels(flags.which_color_pred, flags.num_msi_planes)
        return synth.net_weights(6 * flags.num_psv_planes, n_out, flags.ngf, flags.random_seed,
                                 coord=bool(flags.coord_net)), 0
    ckpt_dir = os.path.join(flags.checkpoint_dir, flags.experiment_name)
    npz = os.path.join(ckpt_dir, "weights.npz")
    if tf_checkpoint.latest_checkpoint(ckpt_dir) is not None:      # test.py:193-202
        w = tf_checkpoint.load_weights(ckpt_dir)
    elif os.path.exists(npz):
        w = tf_checkpoint.load_weights(npz)
    else:
        raise SystemExit(f"no TensorFlow checkpoint ('checkpoint' file) and no weights.npz under {ckpt_dir}; "
                         "pass --random_init for seeded random weights")
    step = int(np.asarray(w["global_step"]).reshape(-1)[0]) if "global_step" in w else 0
    return {k: v for k, v in w.items() if k.startswith("net/")}, step


def main(argv=None):
    flags = parse_flags(argv)
    assert flags.batch_size == 1, "Currently, batch_size must be 1 when testing."  # test.py:89
    rest = flags.test_type
    for tok in ("high_res_only", "high_res", "on_video"):   # flags are concatenated with '_' (test.py:74-75)
        rest = rest.replace(tok, "")
    kinds = set() if rest.strip("_") == "" else {rest}
    if flags.gcn or flags.input_type != "ODS" or kinds:
        raise SystemExit("only the ODS inference path is built (test_type: on_video, high_res, high_res_only concatenated "
                         "with _; no gcn / PP)")
    # flags the reference's test.py accepts that change its outputs and that this driver does not implement: refuse them
    # rather than exit 0 with files missing (test.py:141-146,160-165: jitter_output_*, jitter_msi_*; export-only flags)
    unbuilt = [n for n in ("transform_inverse_reg", "jitter", "smoothed", "net_only") if getattr(flags, n)]
    if unbuilt:
        raise SystemExit("not built in this driver: --" + ", --".join(unbuilt) + " (the jittered sweep itself is available "
                         "through MSI.infer_msi(jitter_pose_inv=...))")
    if flags.rot_factor != 1.0 or flags.tr_factor != 1.0:
        raise SystemExit("--rot_factor / --tr_factor only scale the training-time jitter; not built in this driver")
    import torch
    from matryodshka_b200.msi import MSI, MSIConfig

    if flags.synthetic > 0:
        flags.cameras_glob, flags.image_dir = make_synthetic_dataset(
            os.path.join(flags.output_root, "_synthetic_input"), flags.synthetic, flags.height, flags.width,
            flags.random_seed)
        if "high_res" in flags.test_type:
            _, flags.hres_image_dir = make_synthetic_dataset(
                os.path.join(flags.output_root, "_synthetic_input"), flags.synthetic, flags.hres_height,
                flags.hres_width, flags.random_seed, sub="hres_images")
    seqs = read_camera_lines(flags.cameras_glob)
    if flags.num_runs >= 0:
        seqs = seqs[:flags.num_runs]
    if not seqs:
        raise SystemExit(f"no camera lines under {flags.cameras_glob}")

    weights, step = load_weights(flags)
    cfg = MSIConfig(height=flags.height, width=flags.width, num_psv_planes=flags.num_psv_planes,
                    num_msi_planes=flags.num_msi_planes, min_depth=flags.min_depth, max_depth=flags.max_depth,
                    ngf=flags.ngf, which_color_pred=flags.which_color_pred, coord_net=bool(flags.coord_net),
                    input_type=flags.input_type, operation=flags.operation, conv_impl=flags.conv_impl)
    model = MSI(weights=weights, config=cfg, device=flags.device)
    psv_planes = model.inv_depths(flags.min_depth, flags.max_depth, flags.num_psv_planes)
    msi_planes = model.inv_depths(flags.min_depth, flags.max_depth, flags.num_msi_planes)
    dev = torch.device(flags.device)
    eye = np.eye(4, dtype=np.float32)[None]
    out_root = os.path.join(flags.output_root, flags.experiment_name)
    os.makedirs(out_root, exist_ok=True)

    for run, s in enumerate(seqs):
        if "high_res_only" in flags.test_type:
            break
        imgs = [load_image(os.path.join(flags.image_dir, f"{s['scene_id']}_pos{i}.jpeg"), flags.height, flags.width)
                for i in s["image_id"]]
        ref, src, tgt = (torch.from_numpy(im[None]).to(dev) for im in imgs)  # data_loader.py:133-135
        intrinsics = np.array([[[s["baseline"], 0, 0], [0, 1, 0], [0, 0, 1]]], np.float32)  # data_loader.py:160
        outs, _ = model.infer_msi(src, ref, None, None, eye, eye, intrinsics, flags.which_color_pred,
                                  flags.num_msi_planes, psv_planes, flags.test_outputs, ngf=flags.ngf)
        r = model.msi_render_equirect(outs["rgba_layers"], eye, s["tgt_pos"][None], msi_planes)
        torch.cuda.synchronize(dev)

        dirname = scene_dirname(flags, s)  # test.py:208-217
        output_dir = os.path.join(out_root, dirname)
        os.makedirs(output_dir, exist_ok=True)
        print("Saving to %s" % output_dir)
        if run == 0:
            with open(os.path.join(out_root, "step.txt"), "w") as fh:
                fh.write("%d" % step)
        to = flags.test_outputs
        if "tgt_image" in to:
            write_image(output_dir + "/tgt_image_%s.png" % dirname, imgs[2] * 255.0)
            write_image(output_dir + "/output_tgt_%s.png" % dirname, r["rgb_u8"][0].cpu().numpy())
            write_image(output_dir + "/output_depth_%s.png" % dirname, r["depth_u8"][0].cpu().numpy())
        rgba_layers, tgt_pos1 = outs["rgba_layers"], s["tgt_pos"][None]
        if "ref_output_image" in to:    # test.py:179-187,240-241: the MSI seen from the reference (left) ODS eye
            v = model.deprocess_image(model.msi_render_ods_view(rgba_layers, 1, eye, tgt_pos1, msi_planes, intrinsics))
            write_image(output_dir + "/output_ref_%s.png" % dirname, v[0].cpu().numpy())
        if "src_output_image" in to:    # test.py:171-178,242-243
            v = model.deprocess_image(model.msi_render_ods_view(rgba_layers, -1, eye, tgt_pos1, msi_planes, intrinsics))
            write_image(output_dir + "/output_src_%s.png" % dirname, v[0].cpu().numpy())
        if "psp" in to:                 # test.py:161-170,245-249: four perspective crops
            for vw in range(4):
                v = model.deprocess_image(model.msi_render_perspective_view(rgba_layers, eye, tgt_pos1, msi_planes, intrinsics,
                                                                            viewing_window=vw))
                write_image(output_dir + "/output_ptgt%d_%s.png" % (vw, dirname), v[0].cpu().numpy())
        if "src_image" in to:
            write_image(output_dir + "/src_image_%s.png" % dirname, imgs[1] * 255.0)
        if "ref_image" in to:
            write_image(output_dir + "/ref_image_%s.png" % dirname, imgs[0] * 255.0)
        if "psv" in to:
            psv = outs["psv"].cpu().numpy()
            for j in range(flags.num_psv_planes):
                write_image(output_dir + "/psv_plane_%.3d.png" % j, (psv[0, :, :, j * 3:(j + 1) * 3] + 1.0) / 2.0 * 255)
        if "blend" in flags.which_color_pred and "blend_weights" in to:
            bw = outs["blend_weights"].cpu().numpy()
            np.save(output_dir + "/blend_weights.npy", bw)
            for i in range(flags.num_msi_planes):
                write_image(output_dir + "/blend_weight_%.3d.png" % i, bw[0, :, :, i] * 255.0)
        if "alphas" in to:
            np.save(output_dir + "/alphas.npy", outs["alphas"].cpu().numpy())
        if "rgba_layers" in to:
            rgba = outs["rgba_layers"].cpu().numpy()
            for i in range(flags.num_msi_planes):
                write_image(output_dir + "/msi_alpha_%.2d.png" % i, rgba[0, :, :, i, 3] * 255.0)
                write_image(output_dir + "/msi_rgb_%.2d.png" % i, (rgba[0, :, :, i, :3] + 1.0) / 2.0 * 255)

    if "high_res" in flags.test_type:  # test.py:284-394
        from matryodshka_b200.highres import deprocess_high_res, high_res_rerender
        for s in seqs:
            dirname = scene_dirname(flags, s)
            output_dir = os.path.join(out_root, dirname)
            bw = torch.from_numpy(np.load(output_dir + "/blend_weights.npy")).to(dev)
            al = torch.from_numpy(np.load(output_dir + "/alphas.npy")).to(dev)
            hres = [load_image(os.path.join(flags.hres_image_dir, f"{s['scene_id']}_pos{i}.jpeg"), flags.hres_height,
                               flags.hres_width) for i in s["image_id"][:2]]
            href, hsrc = (torch.from_numpy(im[None]).to(dev) for im in hres)
            intrinsics = np.array([[[s["baseline"], 0, 0], [0, 1, 0], [0, 0, 1]]], np.float32)
            print("Rendering %s at %dx%d, %d planes." % (output_dir, flags.hres_width, flags.hres_height,
                                                          flags.num_psv_planes))
            rgb, dep = high_res_rerender(href, hsrc, bw, al, eye, eye, intrinsics, s["tgt_pos"][None], psv_planes)
            u8, d8 = deprocess_high_res(rgb, dep)
            write_image(output_dir + "/output_hrestgt_%s.png" % dirname, u8.cpu().numpy())
            write_image(output_dir + "/output_hresdepth_%s.png" % dirname, d8.cpu().numpy())
    return 0


if __name__ == "__main__":
    sys.exit(main())

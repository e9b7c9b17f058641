"""The test.py-equivalent driver writes the reference's output tree (test.py:219-281)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_driver_writes_reference_output_tree(tmp_path):
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("msi_test_driver", os.path.join(root, "test.py"))
    drv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(drv)
    out = str(tmp_path / "out")
    rc = drv.main(["--synthetic", "2", "--random_init", "--coord_net", "--experiment_name", "exp", "--output_root", out,
                   "--height", "32", "--width", "64"])
    assert rc == 0
    exp = os.path.join(out, "exp")
    assert open(os.path.join(exp, "step.txt")).read() == "0"
    dirs = sorted(d for d in os.listdir(exp) if os.path.isdir(os.path.join(exp, d)))
    assert dirs == ["synth_000001002", "synth_003004005"]
    d = os.path.join(exp, dirs[0])
    files = set(os.listdir(d))
    for f in ["tgt_image_%s.png", "output_tgt_%s.png", "output_depth_%s.png", "src_image_%s.png", "ref_image_%s.png"]:
        assert f % dirs[0] in files
    for i in range(32):
        assert "msi_alpha_%.2d.png" % i in files and "msi_rgb_%.2d.png" % i in files and "blend_weight_%.3d.png" % i in files
    bw, al = np.load(os.path.join(d, "blend_weights.npy")), np.load(os.path.join(d, "alphas.npy"))
    assert bw.shape == (1, 32, 64, 32) and al.shape == (1, 32, 64, 32) and bw.min() >= 0 and al.max() <= 1
    from PIL import Image
    im = np.asarray(Image.open(os.path.join(d, "output_tgt_%s.png" % dirs[0])))
    assert im.shape == (32, 64, 3) and im.std() > 1


def test_driver_high_res_mode(tmp_path):
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("msi_test_driver", os.path.join(root, "test.py"))
    drv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(drv)
    out = str(tmp_path / "out")
    rc = drv.main(["--synthetic", "1", "--random_init", "--coord_net", "--experiment_name", "exp", "--output_root", out,
                   "--height", "32", "--width", "64", "--hres_height", "96", "--hres_width", "192",
                   "--test_type", "high_res"])
    assert rc == 0
    d = os.path.join(out, "exp", "synth_000001002")
    from PIL import Image
    im = np.asarray(Image.open(os.path.join(d, "output_hrestgt_synth_000001002.png")))
    dep = np.asarray(Image.open(os.path.join(d, "output_hresdepth_synth_000001002.png")))
    assert im.shape == (96, 192, 3) and dep.shape == (96, 192, 3) and im.std() > 1


def _driver():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("msi_test_driver", os.path.join(root, "test.py"))
    drv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(drv)
    return drv


def test_driver_ods_and_perspective_outputs_and_unbuilt_flags(tmp_path):
    """--test_outputs psp / src_output_image / ref_output_image write output_ptgt0-3 / output_src / output_ref
    (test.py:161-187,240-249); flags the driver does not implement are refused, not silently ignored."""
    drv = _driver()
    out = str(tmp_path / "out")
    rc = drv.main(["--synthetic", "1", "--random_init", "--coord_net", "--experiment_name", "exp", "--output_root", out,
                   "--height", "32", "--width", "64", "--test_outputs", "tgt_image_psp_src_output_image_ref_output_image"])
    assert rc == 0
    d = os.path.join(out, "exp", "synth_000001002")
    files = set(os.listdir(d))
    from PIL import Image
    for f in ["output_ptgt0_%s.png", "output_ptgt1_%s.png", "output_ptgt2_%s.png", "output_ptgt3_%s.png", "output_src_%s.png",
              "output_ref_%s.png", "output_tgt_%s.png"]:
        assert f % "synth_000001002" in files
    assert np.asarray(Image.open(os.path.join(d, "output_ptgt0_synth_000001002.png"))).shape == (270, 480, 3)
    assert np.asarray(Image.open(os.path.join(d, "output_src_synth_000001002.png"))).shape == (32, 64, 3)
    assert "msi_rgb_00.png" not in files
    for flag in ("--transform_inverse_reg", "--jitter", "--smoothed", "--net_only"):
        with pytest.raises(SystemExit):
            drv.main(["--synthetic", "1", "--random_init", "--coord_net", "--experiment_name", "exp", "--output_root", out,
                      "--height", "32", "--width", "64", flag])

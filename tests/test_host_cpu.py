"""CPU-side tests: the C-ABI library loads and exports every declared symbol, host logic of the
mirror API, frame sharding, and the gloo world_size-2 all-gather path.  No GPU compute."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from matryodshka_b200 import _lib, nets, ops, synth
from matryodshka_b200.msi import MSI, MSIConfig
from matryodshka_b200.runtime import shard_frames

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "msi_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(msi_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header drifted apart"
    assert lib.msi_b200_abi_version() == 1


def test_net_plan_sizes_without_gpu():
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.msi_net_create(ctypes.byref(h), 320, 640, 192, 64, 64, 1, _lib.CONV_TCGEN05, _lib.PREC_FP16X3) == 0
    assert lib.msi_net_input_c_stride(h) == 192
    assert lib.msi_net_workspace_bytes(h) > 700e6 and lib.msi_net_arena_bytes(h) > 2 * 17.0e6 * 4
    assert lib.msi_net_num_launches_per_forward(h) == 17 * 2 + 1  # conv (LN statistics fused) + normalise, + head
    lib.msi_net_destroy(h)
    # invalid arguments are reported, not crashed on
    assert lib.msi_net_create(ctypes.byref(h), 321, 640, 192, 64, 64, 1, 0, 0) == -1
    assert b"multiples of 8" in lib.msi_last_error()
    assert lib.msi_net_create(ctypes.byref(h), 32, 64, 24, 8, 8, 1, _lib.CONV_TCGEN05, 0) == -3  # needs ngf % 64


def test_compute_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MsiError):
        ops.psv_build(torch.zeros(1, 8, 8, 3), torch.zeros(1, 8, 8, 3), np.eye(4).reshape(1, 1, 16).repeat(2, 1),
                      [0.032], [2.0, 1.0])
    with pytest.raises(_lib.MsiError):
        MSI(weights={}).msi_render_equirect_view(torch.zeros(1, 8, 8, 2, 4), np.eye(4)[None], np.zeros((1, 3)),
                                                 [2.0, 1.0])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "matryodshka_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_arch_table_and_flops():
    shp = nets.layer_shapes(192, 64, 64)
    assert shp["net/conv1_1/weights"] == (3, 3, 193, 64)
    assert shp["net/conv6_1/weights"] == (4, 4, 256, 1024)
    assert shp["net/color_pred/weights"] == (1, 1, 64, 64) and shp["net/color_pred/biases"] == (64,)
    n_params = sum(int(np.prod(s)) for s in shp.values())
    assert abs(n_params - 17.0e6) < 0.1e6
    assert abs(nets.net_flops(320, 640, 192, 64) / 1e9 - 302.4) < 0.05
    hw = nets.layer_geometry(320, 640)
    assert hw["conv4_3"] == (40, 80) and hw["conv8_2"] == (320, 640)


def test_inv_depths_and_tables_match_oracle_restatement():
    from oracle import geometry_np as g, msi_np
    assert MSI().inv_depths(1, 100, 32) == msi_np.inv_depths(1, 100, 32)
    s, t = ops.lat_long_axes(320, 640)
    so, to = g.lat_long_axes((320, 640))
    assert np.array_equal(s, so) and np.array_equal(t, to)  # bit-identical ERP tables


def test_sweep_pose_composition():
    m = MSI()
    a = 0.3
    ref = np.array([[np.cos(a), -np.sin(a), 0, 1], [np.sin(a), np.cos(a), 0, 2], [0, 0, 1, 3], [0, 0, 0, 1]],
                   np.float32)[None]
    poses = m._sweep_poses(ref, ref)
    assert np.allclose(poses[0, 0], np.eye(4), atol=1e-6) and poses.shape == (1, 2, 4, 4)
    poses = m._sweep_poses(np.eye(4)[None], np.eye(4)[None])
    assert np.array_equal(poses[0, 1], np.eye(4, dtype=np.float32))


def test_infer_msi_rejects_unbuilt_modes():
    m = MSI(weights={}, config=MSIConfig(input_type='PP'))  # perspective input: not built (coord_net=False is)
    with pytest.raises(NotImplementedError):
        m.infer_msi(torch.zeros(1, 8, 8, 3), torch.zeros(1, 8, 8, 3), None, None, np.eye(4)[None], np.eye(4)[None],
                    synth.intrinsics(1), "blend_psv", 2, [2.0, 1.0])
    with pytest.raises(NotImplementedError):
        MSI(weights={}).infer_msi(torch.zeros(1, 8, 8, 3), torch.zeros(1, 8, 8, 3), None, None, np.eye(4)[None],
                                  np.eye(4)[None], synth.intrinsics(1), "no_such_scheme", 2, [2.0, 1.0])


def test_shard_frames_partition():
    for n, ws in [(64, 8), (16, 4), (1, 1), (7, 2), (3, 4)]:
        spans = [shard_frames(n, r, ws) for r in range(ws)]
        covered = [i for lo, hi in spans for i in range(lo, hi)]
        assert covered == list(range(n))


def test_synth_is_deterministic():
    a, b = synth.ods_pair(1, 16, 32)
    a2, b2 = synth.ods_pair(1, 16, 32)
    assert np.array_equal(a, a2) and np.array_equal(b, b2) and a.min() >= 0 and a.max() <= 1
    w = synth.net_weights(24, 8, 8)
    assert set(w) == set(nets.layer_shapes(24, 8, 8))


def test_gloo_world2_all_gather_of_frames():
    """The path's only collective on the CPU backend: 2 ranks each own 2 frames."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", script],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GLOO_OK" in r.stdout


def test_driver_scene_dirname_on_video():
    """test.py:209-217: on_video prefixes the output directory with video_[<prefix>_]."""
    import importlib.util
    import os
    import types
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("msi_test_driver_cpu", os.path.join(root, "test.py"))
    drv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(drv)
    s = {"scene_id": "apartment_0", "image_id": ["0001", "0002", "0003"]}
    f = types.SimpleNamespace(test_type="", prefix="")
    assert drv.scene_dirname(f, s) == "apartment_0_000100020003"
    f = types.SimpleNamespace(test_type="on_video_high_res", prefix="run7")
    assert drv.scene_dirname(f, s) == "video_run7_apartment_0_000100020003"
    f = types.SimpleNamespace(test_type="on_video", prefix="")
    assert drv.scene_dirname(f, s) == "video_apartment_0_000100020003"


def _header_prototypes():
    """{name: [parameter type strings]} for every function include/msi_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "msi_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"^\s*#.*$", "", hdr, flags=re.M)
    protos = {}
    for m in re.finditer(r"\b(msi_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
        if params == ["void"] or params == [""]:
            params = []
        protos[m.group(1)] = [re.sub(r"\s+[A-Za-z_][A-Za-z0-9_]*(\[\d*\])*$", "", p).strip() if "*" not in p.split()[-1]
                              else p.rsplit("*", 1)[0].strip() + "*" for p in params]
    return protos


def _type_class(c_type):
    if "*" in c_type:
        return "ptr"
    if c_type in ("float",):
        return "float"
    if c_type in ("double",):
        return "double"
    if c_type in ("size_t", "long long", "unsigned long long", "uint64_t", "int64_t"):
        return "int64"
    return "int32"


def _ctypes_class(t):
    if t in (ctypes.c_void_p, ctypes.c_char_p) or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
        return "ptr"
    if t is ctypes.c_float:
        return "float"
    if t is ctypes.c_double:
        return "double"
    if t in (ctypes.c_size_t, ctypes.c_longlong, ctypes.c_uint64, ctypes.c_int64):
        return "int64"
    return "int32"


def test_ctypes_signatures_match_the_header_prototypes():
    """A drifted argtypes list corrupts the call silently (wrong registers): every entry of _lib.SIGNATURES must have
    the header's parameter count and, per parameter, the same class (pointer / 32-bit int / 64-bit int / float)."""
    protos = _header_prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES)
    for name, (restype, argtypes) in _lib.SIGNATURES.items():
        want = [_type_class(p) for p in protos[name]]
        got = [_ctypes_class(t) for t in argtypes]
        assert got == want, (name, protos[name], got)


def test_threaded_cpu_arm_equals_the_oracle_bit_for_bit():
    """bench.py's CPU arm spreads the oracle's planes / layers over the host cores; the reassembled frame must be
    the same bits as the plain oracle calls (same functions on runs of planes)."""
    from concurrent.futures import ThreadPoolExecutor
    import bench
    from oracle import msi_np
    H, W, P, ngf, seed = 16, 32, 8, 8, 3
    ref, src = synth.ods_pair(1, H, W, seed)
    wts = synth.net_weights(6 * P, 2 * P, ngf, seed)
    tp = synth.target_positions(1, seed)
    planes = msi_np.inv_depths(1, 100, P)
    with ThreadPoolExecutor(max_workers=3) as pool:
        got, _ = bench.oracle_frame_threaded(ref, src, wts, tp, planes, P, ngf, pool, 3)
    eye = synth.identity_poses(1)
    out, net_input = msi_np.infer_msi(src, ref, eye, eye, synth.intrinsics(1), P, planes, wts, ngf=ngf)
    assert np.array_equal(got["net_input"], net_input)
    assert np.array_equal(got["rgba"], out["rgba_layers"])
    e4 = np.eye(4, dtype=np.float32)[None]
    assert np.array_equal(got["view"], msi_np.msi_render_equirect_view(out["rgba_layers"], e4, tp, planes))
    assert np.array_equal(got["depth"], msi_np.msi_render_equirect_depth(out["rgba_layers"], e4, tp, planes))
    assert bench._chunks(32, 16) == [(2 * i, 2 * i + 2) for i in range(16)] and bench._chunks(3, 8) == [(0, 1), (1, 2), (2, 3)]


def test_torch_library_ops_are_registered_with_fake_impls():
    """torch.ops.msi.* (SURVEY.md 8b): every op is registered, infers shapes on meta tensors without touching the CUDA
    library, and has NO CPU kernel (the dispatcher refuses CPU tensors: no fallback)."""
    import torch
    from matryodshka_b200 import torch_ops
    for name in torch_ops.OP_NAMES:
        assert hasattr(torch.ops.msi, name), name
    m = lambda *s, dt=torch.float32: torch.empty(s, device="meta", dtype=dt)  # noqa: E731
    B, H, W, P = 2, 8, 16, 4
    psv = torch.ops.msi.psv_build(m(B, H, W, 3), m(B, H, W, 3), m(B, 2, 16), m(B), m(P), True)
    assert psv.shape == (B, H, W, 6 * P) and psv.device.type == "meta"
    table = torch.ops.msi.sweep_table(m(1, 2, 16), m(1), m(P), H, W)
    assert table.shape == (1, H, W, P, 4)
    assert torch.ops.msi.psv_gather(m(B, H, W, 3, dt=torch.uint8), m(B, H, W, 3, dt=torch.uint8), table, True).shape == psv.shape
    rgba, bw, al, bgw = torch.ops.msi.rgba_assemble(m(B, H, W, 3 * P + 3), psv, 2, P)
    assert rgba.shape == (B, H, W, P, 4) and bw.shape == al.shape == bgw.shape == (B, H, W, P)
    rgb, depth, rgb8, dep8 = torch.ops.msi.render_composite(rgba, m(B, 16), m(B, 3), m(P))
    assert rgb.shape == (B, H, W, 3) and rgb8.dtype == torch.uint8 and depth.dtype == torch.float32
    assert torch.ops.msi.project_layers(rgba, m(B, 16), m(B, 3), m(P)).shape == (P, B, H, W, 4)
    assert torch.ops.msi.intersect_sphere_coords(m(B, 16), m(B, 3), m(P), H, W, True).shape == (B, P, H, W, 2)
    assert torch.ops.msi.resample(m(3, H, W, 5), m(3, 2, 7, 2)).shape == (3, 2, 7, 5)
    assert torch.ops.msi.over_composite(m(P, B, H, W, 4), True).shape == (B, H, W, 3)
    with pytest.raises(NotImplementedError):
        torch.ops.msi.over_composite(torch.zeros(2, 1, 4, 4, 4), False)


def test_reference_module_names_resolve():
    """`from matryodshka.msi import MSI` / `import geometry.projector as pj` (reference test.py:27, msi.py:26-31)."""
    from matryodshka.msi import MSI
    from matryodshka import nets
    import geometry.projector as pj
    import geometry.sampling as sampling
    import geometry.spherical as spherical
    from matryodshka_b200.msi import MSI as Ours
    assert MSI is Ours and MSI().inv_depths(1, 100, 32)[1] == pytest.approx(23.846153846153847)
    for mod, names in ((pj, ["ods_sphere_sweep", "sweep_one", "projective_forward_sphere", "over_composite",
                             "over_composite_depth", "apply_pose"]),
                       (sampling, ["bilinear_wrapper2", "resample"]),
                       (spherical, ["lat_long_grid", "backproject_spherical", "project_ods", "project_spherical",
                                    "intersect_sphere", "theta_phi_to_pixels"]),
                       (nets, ["msi_coord_train_net", "msi_train_net"])):
        for n in names:
            assert callable(getattr(mod, n)), (mod.__name__, n)


def test_bench_stall_guard_reports_what_was_measured(capfd, monkeypatch):
    """bench.py's stall guard: no progress for BENCH_STALL_S seconds -> rank 0 prints the line assembled so far, marked
    partial with the phase it stalled in, and the process ends (exit code 0 when the headline was measured, else 3);
    progress (phase()) re-arms the clock."""
    import json
    import time
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    monkeypatch.delenv("RANK", raising=False)
    exits = []
    bench._STALL.update({"armed": True, "t": time.time(), "phase": "start", "fallback": None})
    bench.phase("device-resident regions done")
    assert bench.stall_guard_tick(limit=150, exit_fn=exits.append) is False and exits == []
    bench._STALL["fallback"] = {"metric": bench.METRIC, "value": 1000.0, "unit": bench.UNIT, "e2e": None}
    assert bench.stall_guard_tick(now=time.time() + 200, limit=150, exit_fn=exits.append) is True
    assert exits == [0]   # the headline was measured: the line is marked partial, the process ends cleanly
    out = capfd.readouterr().out.strip().splitlines()[-1]
    line = json.loads(out)
    assert line["value"] == 1000.0 and line["e2e"] is None
    assert "stalled" in line["partial"] and "device-resident regions done" in line["partial"]
    # disarmed after firing: a second tick does nothing
    assert bench.stall_guard_tick(now=time.time() + 400, limit=150, exit_fn=exits.append) is False
    # nothing measured yet: still one well-formed line
    bench._STALL.update({"armed": True, "t": time.time() - 500, "phase": "pipelines built", "fallback": None})
    assert bench.stall_guard_tick(limit=150, exit_fn=exits.append) is True
    assert exits[-1] == 3
    line = json.loads(capfd.readouterr().out.strip().splitlines()[-1])
    assert line["value"] is None and line["metric"] == bench.METRIC
    # other ranks end silently
    monkeypatch.setenv("RANK", "3")
    bench._STALL.update({"armed": True, "t": time.time() - 500, "phase": "x", "fallback": None})
    assert bench.stall_guard_tick(limit=150, exit_fn=exits.append) is True
    assert capfd.readouterr().out.strip() == ""
    bench._STALL["armed"] = False


def test_bench_run_ours_dry_run_with_the_gpu_layer_stubbed(monkeypatch, capfd):
    """bench.run_ours end to end with the GPU layer replaced by stand-ins (no CUDA here): every statement of the
    measurement flow executes -- regions, reductions, stall-guard fallbacks, roofline assembly, the energy block -- and
    ONE JSON line with the contract's keys comes out.  Guards the bench against a typo that only a GPU box would find."""
    import contextlib
    import json
    import types
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from matryodshka_b200 import runtime as rt

    class FakeEvent:
        def __init__(self, enable_timing=False):
            pass

        def record(self, *a):
            pass

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return 20.0

    class FakeNet:
        ws_bytes = 900 << 20

    class FakePipe:
        def __init__(self):
            self.use_graph, self.fused_rgba, self.static_rig = True, True, True
            self.net, self.hi, self.lo, self.pred_buf = FakeNet(), None, None, None
            self.rgba = torch.zeros(4)
            self.out = {"rgb_u8": torch.arange(12, dtype=torch.uint8)}
            self.h2d_bytes_per_step, self.d2h_bytes_per_step = 4915200, 1228800

        def step(self):
            pass

        def stage_times(self, reps=5):
            return {"psv_build": 0.074, "net": 0.9, "render_composite": 0.066}

    class FakeLanes:
        def __init__(self, *a, lanes=2, **k):
            self.lanes = [FakePipe() for _ in range(lanes)]
            self.streams = [None] * lanes
            self._n = 0

        def __len__(self):
            return len(self.lanes)

        def set_inputs(self, *a, **k):
            pass

        def fork(self):
            pass

        def join(self):
            pass

        def step(self, after_compute=None):
            return self.lanes[0]

        def submit(self, *a, **k):
            self._n += 1

        def collect(self):
            self._n -= 1
            return (self.lanes[0].out["rgb_u8"].clone(), None)

    n_layers = 18
    scopes = [f"l{i}" for i in range(n_layers - 1)] + ["color_pred"]

    def fake_profile(net, hi_lo, out, reps=3, rgba=None, cold_l2=False):
        return scopes, np.full(n_layers, 0.045), np.full(n_layers, 0.009), np.ones(n_layers)

    monkeypatch.setattr(rt, "MSIFrameLanes", FakeLanes)
    monkeypatch.setattr(rt, "profile_net_layers", fake_profile)
    monkeypatch.setattr(_lib, "launch_count", lambda: 0)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(bench, "sustained_energy", lambda *a, **k: {"mJ_per_frame": 987.0, "watts": 950.0})
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    args = types.SimpleNamespace(gpus=1, steps=4, warmup=3, impl="ours", height=32, width=64, planes=32, ngf=64, batch=1,
                                 conv_impl="tcgen05", precision="fp16_fp8x", lanes=2, gather="auto", no_graph=False,
                                 no_cpu_baseline=True, no_layer_profile=False)
    # ---- first the N = 2 flow (rank 0 of a stubbed process group, NCCL form of the gather): reductions, per-rank
    # medians, no energy block
    import torch.distributed as dist
    real_tensor = torch.tensor

    def cpu_tensor(data, *a, **k):
        k.pop("device", None)
        return real_tensor(data, *a, **k)

    def fake_all_gather(out_list, t, *a, **k):
        for o in out_list:
            o.copy_(t)

    monkeypatch.setattr(torch, "tensor", cpu_tensor)
    monkeypatch.setattr(dist, "init_process_group", lambda *a, **k: None)
    monkeypatch.setattr(dist, "all_reduce", lambda *a, **k: None)
    monkeypatch.setattr(dist, "all_gather", fake_all_gather)
    monkeypatch.setattr(dist, "barrier", lambda *a, **k: None)
    monkeypatch.setattr(dist, "destroy_process_group", lambda *a, **k: None)
    monkeypatch.setattr(rt, "all_gather_frames", lambda t, world, group=None: t)
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("LOCAL_RANK", "0")
    args2 = types.SimpleNamespace(**{**vars(args), "gpus": 2, "gather": "nccl"})
    assert bench.run_ours(args2) == 0
    out = [ln for ln in capfd.readouterr().out.strip().splitlines() if ln.startswith("{")]
    assert len(out) == 1
    line2 = json.loads(out[0])
    assert line2["n_gpus"] == 2 and abs(line2["value"] - 2 * 4 / 20.0e-3) < 1e-6 and line2["energy"] is None
    assert len(line2["config"]["timed_regions"]["per_rank_median"]["device_ms_per_step"]) == 2
    assert "NCCL all_gather" in line2["config"]["collective"] and line2["scaling"] == "weak"
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    # ---- then N = 1
    assert bench.run_ours(args) == 0
    assert bench._STALL["armed"] is False   # disarmed before the line is printed
    out = [ln for ln in capfd.readouterr().out.strip().splitlines() if ln.startswith("{")]
    assert len(out) == 1
    line = json.loads(out[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "energy"):
        assert key in line, key
    assert line["metric"] == bench.METRIC and line["n_gpus"] == 1 and line["steps"] == 4 and line["warmup"] == 3
    assert abs(line["value"] - 4 / 20.0e-3) < 1e-6          # 4 frames in the fake 20 ms region
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 4915200
    assert line["energy"]["mJ_per_frame"] == 987.0
    assert line["roofline"]["hbm_kernels"]["psv_build"]["frac"] > 0 and "partial" not in line
    assert line["config"]["frames_in_flight"] == 2 and line["config"]["timed_regions"]["repeats"] == bench.REPEATS

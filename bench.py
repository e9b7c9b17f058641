#!/usr/bin/env python
"""bench.py -- ERP frames/sec of the MSI inference hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun)
    python bench.py --impl reference ...                   (CPU arm: the oracle port of the reference)

Workload (configs[1] of BASELINE.json): 640x320 ERP, 32-sphere MSI, batch = 1 frame per GPU per
step, synthetic ODS pairs + random-init weights of the reference architecture (ngf 64).  A step is
one pass of the hot path over one batch: PSV build -> conv net -> RGBA assembly -> reprojection +
over-composite (+ the all-gather of the rendered frames when N > 1: fused into the render kernel
as multimem / peer stores into every rank's symmetric buffer, `--gather nccl` for the NCCL form).
`--lanes` frames are in flight per GPU (independent pipelines on their own streams, round-robin);
`config.one_frame_at_a_time` reports the same steps on one lane.  Weak scaling: per-GPU work is
fixed as N grows; `value` = frames all ranks processed / max-over-ranks device time.

One JSON line on stdout from rank 0 (everything else goes to stderr).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "erp_frames_per_sec_640x320_32spheres"
UNIT = "frames/s"
MIN_WARMUP = 3   # timing rule: at least 3 warm-up steps (both arms report the warm-up they actually ran)
REPEATS = 10     # the K-step timed region is repeated this many times inside one run; `value` is the median region


def workload_string(W, H, P, batch, ngf, world=1):
    """config.workload, the same string in both arms (the driver compares them).  Names the BASELINE.json config the
    shape belongs to: [1] 640x320 / 32 spheres / batch 1 per GPU (the metric's config, any N); [2] 64 spheres, batch 8;
    [3] 1280x640, batch 16 over 4 GPUs; [4] 640x320 video, batch 64 over 8 GPUs."""
    if (W, H, P) == (640, 320, 64) and batch == 8:
        cfg = "configs[2]"
    elif (W, H, P) == (1280, 640, 32) and batch * world == 16:
        cfg = "configs[3]"
    elif (W, H, P) == (640, 320, 32) and batch * world == 64:
        cfg = "configs[4]"
    elif (W, H, P, batch) == (640, 320, 32, 1):
        cfg = "configs[1]"
    else:
        cfg = "a shape outside configs"
    return f"{W}x{H} ERP, {P}-sphere MSI, batch={batch}/GPU, ngf={ngf} (BASELINE.json {cfg})"


_T0 = time.perf_counter()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Stall guard.  A measurement that stops making progress (a stuck GPU kernel, a rank that never reaches a collective) must
# not cost the whole run: phase() stamps the time of the last progress; a watchdog thread (start_stall_guard) that sees no
# progress for BENCH_STALL_S seconds has rank 0 print the line assembled from what WAS measured (marked "partial", with the
# phase it stalled in) and ends the process (exit code 0 if the headline value is in that line, else 3).
_STALL = {"t": time.time(), "phase": "start", "fallback": None, "armed": False}
_CHILDREN = []   # subprocesses to kill when the stall guard ends the process


def phase(msg):
    """Progress line on stderr (rank-tagged, seconds since start): a run that stalls is attributable from its log."""
    _STALL["t"], _STALL["phase"] = time.time(), msg
    log(f"[bench +{time.perf_counter() - _T0:6.1f}s rank {os.environ.get('RANK', '0')}] {msg}")


def stall_guard_tick(now=None, limit=None, exit_fn=os._exit):
    """One check of the stall guard (separate from the thread so that it can be tested without a GPU).  Returns True when
    it fired."""
    now = time.time() if now is None else now
    limit = float(os.environ.get("BENCH_STALL_S", "120")) if limit is None else limit
    if not _STALL["armed"] or now - _STALL["t"] <= limit:
        return False
    why = f"no progress for {now - _STALL['t']:.0f} s after: {_STALL['phase']}"
    log(f"[bench] STALLED ({why}); ending the run")
    have = _STALL["fallback"] is not None and _STALL["fallback"].get("value") is not None
    if int(os.environ.get("RANK", "0")) == 0:
        line = dict(_STALL["fallback"] or {"metric": METRIC, "value": None, "unit": UNIT, "higher_is_better": True})
        line["partial"] = f"stalled: {why}; only what was measured before the stall is reported"
        emit(line)
    _STALL["armed"] = False
    for proc in list(_CHILDREN):   # (os._exit skips clean-up: do not leave the nvidia-smi sampler running on the box)
        try:
            proc.kill()
        except Exception:
            pass
    # exit code 0 when the headline was measured (the line says `partial`), 3 when there is nothing to report
    exit_fn(0 if have else 3)
    return True


def start_stall_guard():
    _STALL["t"], _STALL["armed"] = time.time(), True

    def loop():
        while True:
            time.sleep(2.0)
            stall_guard_tick()
    threading.Thread(target=loop, daemon=True).start()


# Libraries (NCCL's version banner, for one) write to fd 1.  main() keeps the real stdout for the ONE JSON
# line and sends everything else that lands on fd 1 to stderr (not done at import: tests import this module).
_JSON_FD = None


def reserve_stdout_for_json():
    global _JSON_FD
    if _JSON_FD is None:
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    # B200_PROFILING.md fallback
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def conv_traffic(H, W, P, B, args):
    """DRAM bytes (read + write) of the conv kernel family per step, from the committed
    `ncu --set full` capture (profiles/conv_traffic.json), scaled by the batch; None when the capture
    was taken on a different workload."""
    p = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if not os.path.exists(p) or args.conv_impl != "tcgen05":
        return None
    try:
        t = json.load(open(p))
        t = t.get("by_precision", {}).get(args.precision, t if "workload" in t else None)
        if t is None:
            return None
        w = t["workload"]
        if (w["H"], w["W"], w["P"], w["precision"]) != (H, W, P, args.precision):
            return None
        return {"dram_bytes_per_step": t["dram_bytes_per_forward"] * B / w["B"], "unit": "B",
                "algorithmic_bytes_note": "operands + outputs of the 18 conv launches, once each = 0.73 GB per frame",
                "source": t.get("source", "profiles/conv_traffic.json")}
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            _CHILDREN.append(self.proc)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no_samples"]}
        # "under load": the upper half of the samples (idle samples before/after the region drop out)
        hi = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


def sustained_energy(gpu_index, lanes, one_step, frames_per_step, ms_per_step_hint, seconds=1.5):
    """~`seconds` of back-to-back steps on this rank: frames/s, watts (NVML total-energy counter), mJ per frame and the
    SM clock they settle at.  None when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        limit_w = pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1e3
        chunk = max(1, int(50.0 / max(ms_per_step_hint, 1e-3)))   # ~50 ms of steps between host synchronisations
        torch.cuda.synchronize()
        e0, t0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h), time.perf_counter()
        steps, clocks = 0, []
        while time.perf_counter() - t0 < seconds:
            lanes.fork()
            for _ in range(chunk):
                one_step()
            lanes.join()
            clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))   # sampled while the chunk runs
            torch.cuda.synchronize()
            steps += chunk
        e1, t1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h), time.perf_counter()
        frames = steps * frames_per_step
        return {"seconds": t1 - t0, "frames_per_s_per_gpu": frames / (t1 - t0), "watts": (e1 - e0) / 1e3 / (t1 - t0),
                "power_limit_w": limit_w, "mJ_per_frame": (e1 - e0) / frames, "sm_mhz_median": statistics.median(clocks),
                "note": "continuous steps on this GPU (no host synchronisation for ~50 ms at a time): at the power limit the "
                        "SM clock drops and frames/s = power limit / joules per frame; the headline regions are K-step bursts"}
    except Exception as e:  # pragma: no cover
        log("energy block unavailable:", e)
        return None


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (the reference itself needs Python 2.7 + TF 1.14)
# ------------------------------------------------------------------------------------------------
def _chunks(n, k):
    """n items in at most k contiguous runs: [(start, stop), ...]."""
    k = max(1, min(k, n))
    base, extra = divmod(n, k)
    out, a = [], 0
    for i in range(k):
        b = a + base + (1 if i < extra else 0)
        out.append((a, b))
        a = b
    return out


def oracle_frame_threaded(ref, src, wts, tp, planes, P, ngf, pool, n_threads):
    """One frame of the oracle port with the geometry spread over the host cores: the sweep planes and the
    reprojected layers are independent, so each thread runs the oracle's own functions on a run of planes (NumPy
    releases the GIL inside its loops) and the pieces are reassembled in the reference's channel / layer order.
    Same functions, same arithmetic, same bits as oracle.msi_np.infer_msi + msi_render_equirect_view / _depth
    (checked in tests/test_host_cpu.py); the reference's TF-CPU graph would likewise use all cores."""
    from oracle import geometry_np as g, msi_np, net_torch
    eye_p, intr = np.eye(4, dtype=np.float32)[None], np.array([[[0.032, 0, 0], [0, 1, 0], [0, 0, 1]]], np.float32)
    t0 = time.perf_counter()
    ref_p, src_p = msi_np.preprocess_image(ref), msi_np.preprocess_image(src)
    runs = _chunks(P, n_threads)
    parts = list(pool.map(lambda ab: msi_np.format_network_input(ref_p, src_p, eye_p, eye_p, planes[ab[0]:ab[1]], intr),
                          runs))
    # each part is [1,H,W,6k] = [ref-eye planes | src-eye planes]; the full tensor is [all ref | all src]
    net_input = np.concatenate([q[..., :q.shape[-1] // 2] for q in parts] + [q[..., q.shape[-1] // 2:] for q in parts], axis=3)
    with torch.no_grad():
        pred = net_torch.msi_coord_train_net(torch.from_numpy(net_input), 2 * P, wts, ngf=ngf).numpy()
    rgba, _, _ = msi_np.assemble_rgba(pred, net_input, P)
    t1 = time.perf_counter()
    layers = np.transpose(rgba, (3, 0, 1, 2, 4))                       # [L,B,H,W,4]
    depths = np.asarray(planes, np.float32).reshape(P, 1)

    def project(ab):
        return g.projective_forward_sphere(layers[ab[0]:ab[1]], None, eye_p, tp, depths[ab[0]:ab[1]])

    # the reference reprojects twice, once for the view and once for the depth (msi.py:384-429)
    proj_v = np.concatenate(list(pool.map(project, runs)), axis=0)
    view = g.over_composite([proj_v[i] for i in range(P)])
    proj_d = np.concatenate(list(pool.map(project, runs)), axis=0)
    depth = g.over_composite_depth([proj_d[i] for i in range(P)])
    u8 = (msi_np.deprocess_image(view), msi_np.deprocess_depth_image(depth))
    t2 = time.perf_counter()
    return dict(net_input=net_input, rgba=rgba, view=view, depth=depth, u8=u8), (t1 - t0, t2 - t1)


def oracle_frame_seconds(H, W, P, ngf, n_frames, seed=8964, want_outputs=False):
    """Times the CPU oracle on n_frames full frames (PSV + net + assemble + view + depth render) with all host
    cores.  Returns (seconds per frame list, per-stage seconds of the last frame[, the last frame's outputs])."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import msi_np
    from matryodshka_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref, src = synth.ods_pair(1, H, W, seed)
    wts = synth.net_weights(6 * P, 2 * P, ngf, seed)
    tp = synth.target_positions(1, seed)
    planes = msi_np.inv_depths(1, 100, P)
    times, stages, out = [], {}, None
    with ThreadPoolExecutor(max_workers=cores) as pool:
        for _ in range(n_frames):
            out, (ta, tb) = oracle_frame_threaded(ref, src, wts, tp, planes, P, ngf, pool, cores)
            times.append(ta + tb)
            stages = {"infer_msi_s": ta, "render_view_and_depth_s": tb}
    if want_outputs:
        out.update(tgt_pos=tp, planes=planes)
        return times, stages, out
    return times, stages


def parity_block(pipe, ora, H, W, P, dev):
    """Parity of the bench frame against the oracle's frame (the cpu_baseline run), once per run, so that a kernel
    change that widens the relaxed integer contract shows up in BENCH_r*.json and not only in a test threshold:
    floor() flips of the sample-index grids (sweep table, render coordinates as the fused kernel evaluates them),
    how many of them lie away from a knife edge (> 1e-3 px from an integer: must be 0), validity-mask mismatches,
    uint8 mismatches of the rendered view / depth, and max-abs of the float RGBA layers and the float view."""
    from oracle import geometry_np as g
    from matryodshka_b200 import ops, synth
    F = np.float32
    planes = ora["planes"]
    eye = np.eye(4, dtype=F)[None]

    def flips(c_dev, c_ref):
        bad = np.floor(c_dev) != np.floor(c_ref)
        far = bad & (np.abs(c_ref - np.round(c_ref)) > 1e-3)
        return int(bad.sum()), int(far.sum())

    pipe.step()
    torch.cuda.synchronize(dev)
    out = {"frame": "the bench frame (lane 0, rank 0) vs the oracle port's frame of the cpu_baseline run"}
    # K1: the cached table against project_ods of the oracle
    S, T = g.lat_long_grid((H, W))
    pts = g.backproject_spherical(S, T, F(planes))
    tbl = pipe.sweep_table().table[0].cpu().numpy() if pipe.static_rig else None   # [H,W,P,4]
    if tbl is not None:
        n_bad = n_far = n_mask = 0
        for e, order in enumerate((1, -1)):
            ref_uv, aux = g.project_ods(pts, order, None, synth.intrinsics(1), W, H, return_aux=True)   # [P,H,W,2]
            dev_uv = np.transpose(tbl[..., 2 * e:2 * e + 2], (2, 0, 1, 3))
            ok = aux["valid"]
            n_mask += int((np.all(dev_uv == 1.0, axis=-1) & ok).sum() + (~np.all(dev_uv[~ok] == 1.0, axis=-1)).sum())
            for k in range(2):
                a, b = flips(dev_uv[..., k][ok], ref_uv[..., k][ok])
                n_bad, n_far = n_bad + a, n_far + b
        out["sweep"] = {"samples": int(tbl.size // 2), "index_flips": n_bad, "index_flips_off_edge": n_far,
                        "validity_mask_mismatches": n_mask}
    # K5: the fused kernel's own (fast-chain) coordinates against intersect_sphere of the oracle
    tp = np.asarray(ora["tgt_pos"], F)
    uvf = ops.intersect_sphere_coords(eye, tp, planes, 1, H, W, dev, fast=True).cpu().numpy()[0]
    ref_uv = g.intersect_sphere(eye[0], tp[0], F(planes), P, 1, W, H)
    n_bad = n_far = 0
    for k in range(2):
        a, b = flips(uvf[..., k], ref_uv[..., k])
        n_bad, n_far = n_bad + a, n_far + b
    out["render"] = {"samples": int(uvf.size // 2), "index_flips": n_bad, "index_flips_off_edge": n_far,
                     "max_abs_coord_px": float(np.abs(uvf - ref_uv).max())}
    out["index_flips"] = n_bad + (out["sweep"]["index_flips"] if "sweep" in out else 0)
    out["index_flips_off_edge"] = n_far + (out["sweep"]["index_flips_off_edge"] if "sweep" in out else 0)
    v8, d8 = pipe.out["rgb_u8"].cpu().numpy(), pipe.out["depth_u8"].cpu().numpy()
    out["u8_mismatches"] = int((v8 != ora["u8"][0]).sum() + (d8 != ora["u8"][1]).sum())
    out["u8_values"] = int(v8.size + d8.size)
    out["u8_max_abs"] = int(max(np.abs(v8.astype(int) - ora["u8"][0].astype(int)).max(),
                                np.abs(d8.astype(int) - ora["u8"][1].astype(int)).max()))
    out["max_abs_rgba"] = float(np.abs(pipe.rgba.cpu().numpy() - ora["rgba"]).max())
    out["max_abs_view"] = float(np.abs(pipe.out["rgb"].cpu().numpy() - ora["view"]).max())
    out["tolerance_float"] = 1e-3
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    H, W, P, ngf = args.height, args.width, args.planes, args.ngf
    budget_s = 150.0
    t_start = time.perf_counter()
    # a CPU frame takes seconds: ONE warm-up frame (thread pools, allocator) stands for the requested warm-up and is
    # what the line reports when it differs from --warmup
    warm_frames = 1 if args.warmup > 0 else 0
    for _ in range(warm_frames):
        oracle_frame_seconds(H, W, P, ngf, 1)
    times = []
    stages = {}
    for _ in range(args.steps):
        t, stages = oracle_frame_seconds(H, W, P, ngf, 1)
        times += t
        if time.perf_counter() - t_start > budget_s:
            break
    sec = sum(times) / len(times)
    fps = 1.0 / sec
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(W, H, P, args.batch, ngf, args.gpus),
                   "frames_per_cpu_step": 1,
                   "warmup_frames_run": warm_frames, "steps_requested": args.steps,
                   "note": "oracle port of the reference path on host cores; the reference itself needs "
                           "Python 2.7 + TensorFlow 1.14 and cannot run here.  A step = one full frame (seconds on a "
                           "CPU): the run stops after the requested steps or ~150 s, whichever comes first, and one "
                           "warm-up frame stands for the requested warm-up"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{len(times)} full frame(s): NumPy float32 geometry (planes / layers over {cores} threads) + "
                                   f"torch-CPU float32 conv net ({torch.get_num_threads()} threads); stages {stages}"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from matryodshka_b200 import _lib, synth
    from matryodshka_b200.nets import net_flops
    from matryodshka_b200.runtime import FrameGather, MSIFrameLanes, all_gather_frames, profile_net_layers

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"note: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (matryodshka_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    H, W, P, ngf, Bp = args.height, args.width, args.planes, args.ngf, args.batch
    K, Wm = args.steps, max(args.warmup, MIN_WARMUP)
    start_stall_guard()

    def reduce_max(xs):
        """Max over ranks of every repeat (a region ends when the slowest rank ends)."""
        if world == 1:
            return list(xs)
        t = torch.tensor(xs, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def fallback_line(dev_red, e2e_red=None):
        """The line as far as it is known (stall guard): the headline from the device-resident regions, e2e when measured."""
        frames = world * Bp * K
        d = statistics.median(dev_red)
        return {"metric": METRIC, "value": frames / (d * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": d / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic",
                "config": {"workload": workload_string(W, H, P, Bp, ngf, world), "frames_per_step": world * Bp,
                           "precision": args.precision, "frames_in_flight": max(1, args.lanes)},
                "e2e": None if e2e_red is None else {"value": frames / (statistics.median(e2e_red) * 1e-3), "unit": UNIT,
                                                      "ms_per_step": statistics.median(e2e_red) / K}}
    seed = 8964 + rank
    ref, src = synth.ods_pair(Bp, H, W, seed)
    wts = synth.net_weights(6 * P, 2 * P, ngf, 8964)
    tp = synth.target_positions(Bp, seed)
    # `lanes` frames in flight: independent pipelines (own workspace / CUDA graph / stream) fed round-robin,
    # so that one frame's kernels fill the tail and ramp bubbles of the other's (runtime.MSIFrameLanes)
    n_lanes = max(1, args.lanes)
    lanes = MSIFrameLanes(wts, H, W, P, ngf, lanes=n_lanes, batch=Bp, device=dev, conv_impl=args.conv_impl,
                          precision=args.precision, use_graph=not args.no_graph)
    lanes.set_inputs(ref, src, tgt_pos=tp)
    pipe = lanes.lanes[0]
    phase(f"pipelines built: {n_lanes} lane(s), batch {Bp}, world {world}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def check_gathered():
        """After a barrier: every rank's slot of every gathered buffer holds that rank's last frame."""
        for lane, g in zip(lanes.lanes, gathers):
            torch.cuda.synchronize(dev)
            g.barrier()
            mine = g.frames[rank * Bp:(rank + 1) * Bp]   # (first_frame counts from the start of the whole allocation)
            assert torch.equal(mine, lane.out["rgb_u8"]), "fused gather: own slot differs from the local frame"
            per_rank = g.frames.view(world, -1).float().mean(dim=1)
            assert bool((per_rank > 0).all()), "fused gather: a rank's slot is empty"
            g.barrier()

    # The path's only collective, the all-gather of the rendered uint8 frames.  Default:
    # FUSED into the render kernel -- it stores its frame into every rank's gathered buffer (symmetric
    # memory; multimem.st through the NVSwitch multicast address, else peer stores), no collective kernel
    # and no per-step rendezvous.  --gather nccl (or no symmetric memory): all_gather_into_tensor per step.
    gather, gathers, collective = None, [], "none"
    if world > 1:
        if args.gather != "nccl":
            # ONE symmetric allocation for all lanes (one rendezvous, one multicast object), set up under a time limit:
            # if it cannot be had within 75 s on EVERY rank, all ranks use the NCCL all-gather instead (the stall guard allows 120 s)
            box = {}

            def setup():
                try:
                    box["g"] = FrameGather(Bp, H, W, dev, mode=args.gather, slots=n_lanes)
                except Exception as e:  # noqa: BLE001
                    box["err"] = f"{type(e).__name__}: {e}"

            phase("setting up the symmetric gathered buffer")
            th = threading.Thread(target=setup, daemon=True)
            th.start()
            th.join(75.0)
            ok = torch.tensor([1 if "g" in box else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                gathers = [box["g"].slot(k) for k in range(n_lanes)]
                for lane, g in zip(lanes.lanes, gathers):
                    lane.attach_gather(g)
                collective = ("fused into the render kernel: uint8 frames stored into every rank's gathered buffer, "
                              + gathers[0].mode + " (symmetric memory); barrier at the end of the timed region")
            else:
                log(f"fused gather unavailable on some rank ({box.get('err', 'timed out' if 'g' not in box else 'ok here')}); "
                    "using NCCL all_gather")
                gathers = []
        if not gathers:
            gather = lambda lane: all_gather_frames(lane.out["rgb_u8"], world)  # noqa: E731
            collective = "NCCL all_gather_into_tensor of the rendered uint8 frames, every step"

    def one_step():
        lanes.step(after_compute=gather)

    def timed(n, step_fn):
        """Device time of n steps: CUDA events on the current stream around fork / join of the lane streams."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        lanes.fork()
        for _ in range(n):
            step_fn()
        lanes.join()
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    # launches per step, counted on one eager (un-graphed) step
    was_graph = pipe.use_graph
    pipe.use_graph = False
    c0 = _lib.launch_count()
    pipe.step()
    torch.cuda.synchronize(dev)
    launches_per_step = _lib.launch_count() - c0
    pipe.use_graph = was_graph

    phase(f"collective: {collective}")
    lanes.fork()
    for _ in range(Wm * n_lanes):
        one_step()
    lanes.join()
    barrier()
    phase("warm-up done")

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- device-resident timed region: inputs already in HBM ------------------------------------
    # EXACTLY K steps between the two events, barrier + synchronize on both sides; the region is repeated REPEATS
    # times (a 20-step region is ~20 ms: one scheduling hiccup on one rank moves a single sample by percents) and
    # the line reports the median region, with min / max beside it
    dev_all = [timed(K, one_step)]
    _STALL["fallback"] = fallback_line(reduce_max(dev_all))   # (the first region alone: what a stall in the others would leave)
    phase("first device-resident region done")
    dev_all += [timed(K, one_step) for _ in range(REPEATS - 1)]
    phase(f"device-resident regions done: median {statistics.median(dev_all) / K:.4f} ms/step")
    dev_own = list(dev_all)
    dev_red = reduce_max(dev_all)
    _STALL["fallback"] = fallback_line(dev_red)
    if gathers:
        check_gathered()
        phase("gathered buffers checked")
    # the same K steps on ONE lane (one frame at a time), reported beside the headline for reference
    single_all = dev_all
    if n_lanes > 1:
        def lane0_step():
            with torch.cuda.stream(lanes.streams[0]):
                pipe.step()
                if gather is not None:
                    gather(pipe)
        single_all = [timed(K, lane0_step) for _ in range(max(3, REPEATS // 3))]

    # ---- end-to-end region: host (pinned) images in, host uint8 view + depth out ----------------
    # Every step copies its inputs host -> device from pinned memory and its results (uint8 view +
    # depth) device -> host; the public streaming API (submit / collect) keeps two batches in flight so
    # that those copies overlap the neighbouring batches' compute.  Wall clock, all K steps collected.
    h_ref, h_src = torch.from_numpy(ref).pin_memory(), torch.from_numpy(src).pin_memory()
    in_flight = 2 * n_lanes  # every lane double-buffers its copies

    def e2e_loop(n):
        last = None
        for i in range(n):
            lanes.submit(h_ref, h_src, after_compute=gather)
            if i >= in_flight - 1:
                last = lanes.collect()
        for _ in range(min(n, in_flight - 1)):
            last = lanes.collect()
        return last

    phase("one-lane regions done")
    e2e_loop(2 * in_flight)
    e2e_all = []
    for _ in range(REPEATS):
        barrier()
        t0 = time.perf_counter()
        last = e2e_loop(K)
        barrier()
        e2e_all.append((time.perf_counter() - t0) * 1e3)
    assert torch.equal(last[0], pipe.out["rgb_u8"].cpu()), "e2e result differs from the device-resident result"
    phase(f"end-to-end regions done: median {statistics.median(e2e_all) / K:.4f} ms/step")
    e2e_own = list(e2e_all)
    e2e_red = reduce_max(e2e_all)
    _STALL["fallback"] = fallback_line(dev_red, e2e_red)

    # ---- per-kernel timing for the roofline (CUDA events on the launching stream) ---------------
    if args.no_layer_profile:  # e.g. under `ncu` for the launch list: only the real steps' kernels
        n_layers = 18
        scopes, conv_ms, ln_ms = ["-"] * n_layers, np.full(n_layers, np.nan), np.full(n_layers, np.nan)
    else:
        # the head as the pipeline runs it (fused RGBA assembly); warm = back-to-back repeats (the tensor-bound layers),
        # cold = a 256 MB flush before every launch (reported beside it; the HBM-bound head is quoted cold)
        fused_out = pipe.rgba if pipe.fused_rgba else None
        scopes, conv_ms, ln_ms, _ = profile_net_layers(pipe.net, (pipe.hi, pipe.lo), pipe.pred_buf, reps=max(3, min(K, 10)),
                                                       rgba=fused_out)
        _, conv_cold, ln_cold, _ = profile_net_layers(pipe.net, (pipe.hi, pipe.lo), pipe.pred_buf, reps=3, rgba=fused_out,
                                                      cold_l2=True)
    stage_ms = None if args.no_layer_profile else pipe.stage_times(reps=5)
    clocks = sampler.stop() if rank == 0 else None
    phase("per-kernel profile done")

    # max over ranks of every repeat (a region ends when the slowest rank ends), then the median repeat
    per_rank = None
    if world > 1:
        mine = torch.tensor([statistics.median(dev_own), statistics.median(e2e_own)], device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"device_ms_per_step": [round(float(a[0]) / K, 5) for a in allr],
                    "e2e_ms_per_step": [round(float(a[1]) / K, 5) for a in allr]}
    dev_all, e2e_all, single_all = dev_red, e2e_red, reduce_max(single_all)
    dev_ms, e2e_ms, single_ms = statistics.median(dev_all), statistics.median(e2e_all), statistics.median(single_all)

    if rank == 0:
        peaks, peak_src = load_peaks()
        frames = world * Bp * K
        value = frames / (dev_ms * 1e-3)
        e2e = frames / (e2e_ms * 1e-3)
        conv_total_ms = float(conv_ms.sum())
        algo_flops = net_flops(H, W, 6 * P, 2 * P, ngf) * Bp
        achieved = algo_flops / (conv_total_ms * 1e-3) / 1e12
        # MMA work units per algorithmic product: fp16x3 = 3 fp16 MMAs; fp16_fp8x = 1 fp16 MMA + 1 e4m3 MMA of twice the
        # K (= one more unit of pipe time) on the Cout >= 128 layers (70 % of the FLOPs), 3 on the rest
        mma_mult = {"fp16x3": 3.0, "fp16_fp8x": 2.30}.get(args.precision, 1.0) if args.conv_impl == "tcgen05" else 1.0
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        roofline = {
            "kernel": "conv_halo_tcgen05 as cta_group::2 CTA pairs (13 launches; conv8_2 single-CTA) + conv_igemm_tcgen05 "
                      "(3 stride-2 convs + 1x1 head)" if args.conv_impl == "tcgen05"
                      else "conv_simt (fp32 CUDA-core bring-up back end)",
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": conv_traffic(H, W, P, Bp, args),
            "peak_source": f"{peak_src}: cuBLAS bf16 sustained (kernel timed inside a long step)",
            "algorithmic_gflop_per_step": algo_flops / 1e9,
            "kernel_ms_per_step": conv_total_ms,
            "executed_mma_tflops": achieved * mma_mult,
            "note": "achieved = algorithmic FLOPs (SURVEY 8a a10 table, coord channels counted) / sum of conv "
                    "launch durations; fp16x3 issues 3 MMAs per algorithmic product (hi*hi + lo*hi + hi*lo), "
                    "executed_mma_tflops counts those (fp16_fp8x: pipe-time units, an e4m3 MMA of 2K counted as one)",
            "per_layer_ms": {s: round(float(c), 4) for s, c in zip(scopes, conv_ms)},
            "layernorm_ms_per_step": float(ln_ms.sum()),
            "cold_l2": None if args.no_layer_profile else {
                "per_layer_ms": {s: round(float(c), 4) for s, c in zip(scopes, conv_cold)},
                "kernel_ms_per_step": float(conv_cold.sum()), "layernorm_ms_per_step": float(ln_cold.sum()),
                "note": "every launch timed alone after a 256 MB L2 flush (inputs from HBM); the figures above time "
                        "back-to-back repeats of each launch"},
        }
        if not args.no_layer_profile:
            # the metric's second half: achieved HBM GB/s of the warp (K1) and composite (K5) kernels, and of
            # the RGBA assembly (K4), on ALGORITHMIC bytes, each kernel timed alone
            s_out = 2  # the PSV leaves K1 as fp16 hi + fp16 lo = 4 bytes per element, like float32
            npx = Bp * H * W
            algo = {"psv_build": 2 * npx * 3 * 4 + npx * 6 * P * 2 * s_out,
                    "render_composite": npx * 4 * P * 4 + 2 * npx * 3 * 4,
                    "rgba_assemble": npx * 2 * P * 4 + npx * 6 * P * 2 * s_out + npx * 4 * P * 4}
            hbm_peak = float(peaks["hbm_gbs"])
            times_ms = dict(stage_ms)
            if pipe.fused_rgba:
                # K4 lives in the head's epilogue: the head launch reads conv8_2 (hi + lo, as many bytes as `pred`) and
                # the PSV and writes the RGBA layers -- the bytes of SURVEY 8d's K4 -- timed as the color_pred launch
                times_ms["rgba_assemble"] = float(conv_cold[list(scopes).index("color_pred")])
            roofline["hbm_kernels"] = {
                k: {"ms": times_ms[k], "algorithmic_bytes": algo[k], "achieved_gbs": algo[k] / (times_ms[k] * 1e-3) / 1e9,
                    "frac": algo[k] / (times_ms[k] * 1e-3) / 1e9 / hbm_peak} for k in algo}
            hk = roofline["hbm_kernels"]
            if pipe.fused_rgba:
                hk["rgba_assemble"]["kernel"] = "fused into the color_pred head epilogue (msi_net_forward_rgba): no separate launch"
            if pipe.static_rig:
                # what the gather really moves: the cached coordinate table is read once per frame on top of the
                # algorithmic bytes (16 B per (pixel, plane)); reported so that `frac` is not mistaken for DRAM idleness
                tbl = npx * P * 16
                hk["psv_build"]["kernel"] = "prep_images + psv_gather_pair (coordinates from the per-rig table, ops.sweep_table)"
                hk["psv_build"]["bytes_moved_incl_table"] = algo["psv_build"] + tbl
                hk["psv_build"]["dram_frac_incl_table"] = (algo["psv_build"] + tbl) / (times_ms["psv_build"] * 1e-3) / 1e9 / hbm_peak
            wc = (algo["psv_build"] + algo["render_composite"]) / ((times_ms["psv_build"] + times_ms["render_composite"]) * 1e-3) / 1e9
            roofline["warp_plus_composite"] = {"bound": "hbm", "achieved": wc, "peak": hbm_peak, "unit": "GB/s",
                                               "frac": wc / hbm_peak,
                                               "note": "K1 gathers from a cached per-rig coordinate table (bits identical to the "
                                                       "per-frame chain) and is bound by load latency / L1 throughput; K5 evaluates "
                                                       "its per-frame coordinates with a fast chain and is instruction-issue bound "
                                                       "(profiles/r2_geom_ncu_full_summary.csv)"}
            roofline["net_ms_per_step"] = stage_ms["net"]
        if args.no_layer_profile:
            roofline = None
        cpu, parity = None, None
        if world == 1 and not args.no_cpu_baseline:
            times, stages, ora = oracle_frame_seconds(H, W, P, ngf, 1, want_outputs=True)
            cpu = {"value": 1.0 / times[0], "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"1 full frame of the same workload through the oracle port: NumPy float32 geometry "
                             f"(planes / layers over {os.cpu_count() or 1} threads) + torch-CPU float32 net "
                             f"({torch.get_num_threads()} threads); {stages}"}
            if Bp == 1:   # (the oracle frame is frame 0 of a batch-1 run of the same seed)
                parity = parity_block(pipe, ora, H, W, P, dev)
        ws_gb = (pipe.net.ws_bytes + pipe.rgba.numel() * 4) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16x3": "f16x3-split operands, f32 accumulate",
                      "fp16_fp8x": "f16 main product + e4m3 cross terms (f16x3 on the Cout = 64 layers), f32 accumulate",
                      "fp16": "f16 operands, f32 accumulate"}[args.precision] if args.conv_impl == "tcgen05" else "f32",
            "data": "synthetic",
            "config": {"workload": workload_string(W, H, P, Bp, ngf, world),
                       "frames_per_step": world * Bp, "conv_impl": args.conv_impl, "precision": args.precision,
                       "timed_regions": {"repeats": REPEATS, "steps_each": K, "statistic": "median (max over ranks per repeat)",
                                         "device_ms_per_step_min_med_max": [min(dev_all) / K, dev_ms / K, max(dev_all) / K],
                                         "e2e_ms_per_step_min_med_max": [min(e2e_all) / K, e2e_ms / K, max(e2e_all) / K],
                                         "per_rank_median": per_rank},
                       "sweep_coordinates": "cached per-rig table (static rig)" if pipe.static_rig else "evaluated per frame",
                       "rgba_assembly": "fused into the head epilogue" if pipe.fused_rgba else "separate kernel",
                       "cuda_graph": not args.no_graph, "frames_in_flight": n_lanes,
                       "one_frame_at_a_time": {"value": frames / (single_ms * 1e-3), "unit": UNIT,
                                               "ms_per_step": single_ms / K},
                       "collective": collective,
                       "l2": f"per-step working set {ws_gb:.2f} GB exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes_per_step,
                    "d2h_bytes_per_step": pipe.d2h_bytes_per_step, "ms_per_step": e2e_ms / K,
                    "input": "float32 host images (pinned) -> uint8 view + depth on host; MSIFrameLanes.submit/collect, "
                             f"{n_lanes} lane(s) x 2 batches in flight (H2D / compute / D2H on three streams per lane)"},
            "gpu_launches": int(launches_per_step * K),
            "clocks": clocks,
            "energy": None,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": parity,
        }
        # ---- sustained region + energy: the frame rate sits on the board's power limit ------------------------------------
        # The K-step regions above are short bursts (tens of ms between host synchronisations); run continuously, the step
        # draws the board's power limit and the SM clock settles below its maximum, so frames/s is bounded by JOULES PER
        # FRAME.  This block reports that regime beside the headline: ~1 s of back-to-back steps, the NVML total-energy
        # counter around them, SM clock sampled through NVML (scripts/exp_energy.py attributes the joules).  It runs last
        # and on one GPU only: the complete line is already the stall guard's fallback, so a stall here costs only this key.
        _STALL["fallback"] = line
        phase("line assembled")
        if world == 1 and not args.no_layer_profile:
            line["energy"] = sustained_energy(local_rank, lanes, one_step, Bp, dev_ms / K, seconds=1.0)
            phase("sustained / energy region done")
        _STALL["armed"] = False
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)  # 200 x 1.5 ms: long enough for the clock sampler
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--height", type=int, default=320)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--planes", type=int, default=32)
    ap.add_argument("--ngf", type=int, default=64)
    ap.add_argument("--batch", type=int, default=1, help="frames per GPU per step")
    ap.add_argument("--conv-impl", default="tcgen05", choices=["tcgen05", "simt"])
    ap.add_argument("--precision", default="fp16_fp8x", choices=["fp16_fp8x", "fp16x3", "fp16"])
    ap.add_argument("--lanes", type=int, default=4, help="frames in flight per GPU (independent pipelines on own streams)")
    ap.add_argument("--gather", default="auto", choices=["auto", "multicast", "peer", "nccl"],
                    help="N > 1: how the rendered frames reach every rank (default: fused into the render kernel)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-layer-profile", action="store_true",
                    help="skip the per-kernel timing pass (roofline numbers become NaN); for ncu launch lists")
    args = ap.parse_args()
    reserve_stdout_for_json()
    # a stalled run dumps every thread's stack to stderr (and again every 120 s) instead of dying silently at the
    # driver's timeout
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("BENCH_WATCHDOG_S", "180")), repeat=True, file=sys.stderr)
    if args.impl == "reference":
        return run_reference(args)
    if os.environ.get("MSI_BEACON") == "1":
        # debugging: a run that is still going after BENCH_BEACON_S seconds prints which conv launches have CTAs that
        # never exited and the phase each of their roles reached (msi_debug_beacon_dump reads host-mapped memory)
        def beacon_dump():
            time.sleep(float(os.environ.get("BENCH_BEACON_S", "30")))
            from matryodshka_b200 import _lib
            _lib.load().msi_debug_beacon_dump()
            sys.stderr.flush()
        threading.Thread(target=beacon_dump, daemon=True).start()
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

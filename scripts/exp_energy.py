"""Energy per stage (NVML total-energy counter): each stage of the frame looped alone for ~2.5 s on one GPU.

The frame rate sits on the board's power cap (bench.py: reasons = sw_power_cap, SM clock below max), so the
quantity that bounds frames/s is JOULES PER FRAME, not kernel time.  This script attributes them:
K1 (gather), the conv net in the three precisions (fp16 = 1 MMA unit per product, fp16_fp8x = 2 / 3, fp16x3 = 3:
the differences are the energy of one MMA unit over the whole net), K5 (render + composite), and the whole step.
    gpurun --timeout 300 -- 'python scripts/exp_energy.py > gpurun_out/exp_energy.json'
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pynvml
from matryodshka_b200 import synth
from matryodshka_b200.runtime import MSIPipeline, MSIFrameLanes

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
H, W, P, ngf = 320, 640, 32, 64
SECS = float(os.environ.get("SECS", "2.5"))
wts = synth.net_weights(6 * P, 2 * P, ngf)
ref, src = synth.ods_pair(1, H, W)
tp = synth.target_positions(1)


def measure(fn, name, per=1):
    fn(); torch.cuda.synchronize()
    # calibrate
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    n_chunk = max(1, int(0.05 / dt))
    time.sleep(0.5)
    torch.cuda.synchronize()
    e0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    t0 = time.perf_counter()
    n = 0
    clocks = []
    while time.perf_counter() - t0 < SECS:
        for _ in range(n_chunk):
            fn()
        n += n_chunk
        torch.cuda.synchronize()
        clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    t1 = time.perf_counter()
    e1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    clocks.sort()
    r = {"stage": name, "launches": n, "ms_per_launch": (t1 - t0) / n * 1e3 / per, "watts": (e1 - e0) / 1e3 / (t1 - t0),
         "mJ_per_launch": (e1 - e0) / n / per, "sm_mhz_median": clocks[len(clocks) // 2]}
    print(json.dumps(r), flush=True)
    return r


out = []
# idle power
time.sleep(1.0)
e0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h); t0 = time.perf_counter(); time.sleep(2.0)
idle_w = (pynvml.nvmlDeviceGetTotalEnergyConsumption(h) - e0) / 1e3 / (time.perf_counter() - t0)
print(json.dumps({"stage": "idle (context up)", "watts": idle_w}), flush=True)
for prec in ["fp16_fp8x", "fp16x3", "fp16"]:
    pipe = MSIPipeline(wts, H, W, P, ngf, batch=1, device="cuda", use_graph=False, precision=prec)
    pipe.set_inputs(ref, src, tgt_pos=tp)
    pipe.step(); torch.cuda.synchronize()
    st = dict(pipe._stages())
    if prec == "fp16_fp8x":
        out.append(measure(st["psv_build"], "K1 psv_gather"))
        out.append(measure(st["render_composite"], "K5 render_composite"))
    out.append(measure(st["net"], f"net ({prec}: convs + LayerNorm + fused head)"))
    del pipe
    torch.cuda.empty_cache()
lanes = MSIFrameLanes(wts, H, W, P, ngf, lanes=4, batch=1, device="cuda")
lanes.set_inputs(ref, src, tgt_pos=tp)


def four():
    lanes.fork()
    for _ in range(4):
        lanes.step()
    lanes.join()


out.append(measure(four, "whole frame, 4 lanes (per frame)", per=4))

"""MSIFrameLanes: several frames in flight on one GPU (independent pipelines on their own streams)
must return, in order, exactly what the one-frame-at-a-time pipeline computes."""
import numpy as np
import pytest
import torch

from matryodshka_b200 import synth
from matryodshka_b200.runtime import MSIFrameLanes, MSIPipeline

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_lanes_match_single_pipeline_bit_for_bit():
    H, W, P, ngf = 32, 64, 32, 64
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    frames = [synth.ods_pair(1, H, W, seed=200 + i) for i in range(7)]
    single = MSIPipeline(wts, H, W, P, ngf, batch=1, device=DEV)
    want = []
    for ref, src in frames:
        single.set_inputs(ref, src)
        single.step()
        torch.cuda.synchronize()
        want.append((single.out["rgb_u8"].cpu().clone(), single.out["depth_u8"].cpu().clone()))

    lanes = MSIFrameLanes(wts, H, W, P, ngf, lanes=3, batch=1, device=DEV)
    # device-resident round-robin steps: lane k holds frame k
    for k, (ref, src) in enumerate(frames[:3]):
        lanes.lanes[k].set_inputs(ref, src)
    torch.cuda.synchronize()
    lanes.fork()
    used = [lanes.step() for _ in range(3)]
    lanes.join()
    torch.cuda.synchronize()
    for k, lane in enumerate(used):
        assert lane is lanes.lanes[k]
        assert torch.equal(lane.out["rgb_u8"].cpu(), want[k][0])
        assert torch.equal(lane.out["depth_u8"].cpu(), want[k][1])

    # end-to-end submit / collect, more frames than lanes x depth, results come back in submission order
    got = []
    cap = 2 * len(lanes)
    for i, (ref, src) in enumerate(frames):
        lanes.submit(torch.from_numpy(ref).pin_memory(), torch.from_numpy(src).pin_memory())
        if i >= cap - 1:
            r = lanes.collect()
            got.append((r[0].clone(), r[1].clone()))
    while len(got) < len(frames):
        r = lanes.collect()
        got.append((r[0].clone(), r[1].clone()))
    for (a, b), (c, d) in zip(want, got):
        assert torch.equal(a, c) and torch.equal(b, d)

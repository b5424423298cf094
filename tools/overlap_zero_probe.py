"""Probe: can the 1.15 GB zero fill of the dense dL_dsh tensor hide under the (issue-bound) blend backward?

Times the cfg3 step (bench.py's Runner, inputs resident) three ways: as is; with a side-stream memset of a P x 144
float buffer started when the backward starts; with a side-stream device-to-device copy (copy engine) of the same size
from a zero source.  If the step barely moves, sh_bwd can stop writing the 793 MB of zero rows (profiles/README.md).
Usage (GPU box):  python tools/overlap_zero_probe.py
"""
import os
import sys
import json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench


def main():
    device = "cuda:0"
    torch.cuda.set_device(0)
    wl = bench.Workload("cfg3", device, [0])
    r = bench.Runner(wl, "ours", 1, 1)
    P = wl.P
    buf = torch.empty(P, 144, device=device)
    zsrc = torch.zeros(P // 8, 144, device=device)
    side = torch.cuda.Stream(device=device)
    mode = {"m": "none"}

    def step():
        cur = torch.cuda.current_stream()
        wl_params = wl.params
        for v in wl_params.values():
            v.grad = None
        (color, radii, depth, alpha, flow, covs), means2D = r._raster(wl.settings[0])
        loss = (color * r.G_dev[0]).sum()
        ev = torch.cuda.Event()
        ev.record(cur)
        if mode["m"] != "none":
            with torch.cuda.stream(side):
                side.wait_event(ev)
                if mode["m"] == "memset":
                    buf.zero_()
                elif mode["m"] == "ce_copy":
                    for k in range(8):
                        buf[k * (P // 8):(k + 1) * (P // 8)].copy_(zsrc, non_blocking=True)
                done = torch.cuda.Event()
                done.record(side)
        loss.backward()
        if mode["m"] != "none":
            cur.wait_event(done)

    out = {}
    for m in ("none", "memset", "ce_copy", "none", "memset", "ce_copy"):
        mode["m"] = m
        ms, _ = bench.timed(step, 20, 5, device, 1)
        out.setdefault(m, []).append(round(ms, 4))
    # the fills alone
    for m in ("memset", "ce_copy"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            if m == "memset":
                buf.zero_()
            else:
                for k in range(8):
                    buf[k * (P // 8):(k + 1) * (P // 8)].copy_(zsrc, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        out[m + "_alone_ms"] = round(e0.elapsed_time(e1) / 10, 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

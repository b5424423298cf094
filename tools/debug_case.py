#!/usr/bin/env python3
"""Run one synthetic configuration through the extension with debug=True (every stage synchronised;
FDGS_TRACE=1 prints the stage names).  Usage: debug_case.py P W H [seed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("4d-gaussian-splatting_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch  # noqa: E402
import fdgs  # noqa: E402
import helpers  # noqa: E402

P, W, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 4242 + P
cfg = dict(P=P, W=W, H=H, seed=seed)
cfg, cam, sc, st = helpers.build(cfg, device="cuda")
st["debug"] = True
C = fdgs.ext()
fw = C.rasterize_gaussians(*helpers.fwd_args(st, sc, cfg))
torch.cuda.synchronize()
print("forward ok, num_rendered", int(fw[0]), flush=True)
bw = C.rasterize_gaussians_backward(*helpers.bwd_args(st, sc, cfg, fw, helpers.pixel_grads(cfg, device="cuda")))
torch.cuda.synchronize()
print("backward ok", flush=True)

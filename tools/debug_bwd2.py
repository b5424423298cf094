#!/usr/bin/env python3
"""Compare the colour-only backward (blend_bwd v2) with the compiled reference on one configuration and
print the Gaussians with the largest dL_dmeans2D error.  Usage: debug_bwd2.py <config name>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("4d-gaussian-splatting_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch  # noqa: E402
import fdgs  # noqa: E402
import helpers  # noqa: E402
import oracle_py  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
DEV = "cuda"
C = fdgs.ext()
ref = oracle_py.ref_module()
cfg, cam, sc, st = helpers.build(name, device=DEV)
gc, gd, ga, gf = helpers.pixel_grads(cfg, device=DEV)
e = torch.empty(0, device=DEV)
fw = C.rasterize_gaussians(*helpers.fwd_args(st, sc, cfg))
ours = C.rasterize_gaussians_backward(*helpers.bwd_args(st, sc, cfg, fw, (gc, e, e, e)))
rf = ref.rasterize_gaussians(*helpers.fwd_args(st, sc, cfg))
rb = ref.rasterize_gaussians_backward(*helpers.bwd_args(st, sc, cfg, rf, (gc, 0 * gd, 0 * ga, 0 * gf)))
torch.cuda.synchronize()
geom = C.debug_export_geom(fw[6], cfg["P"])
conic = geom[2]
m2o, m2r = ours[0], rb[0]
err = (m2o - m2r).abs().max(dim=1).values
scale = m2r.abs().max().item()
print("max |ref| =", scale, " max err =", err.max().item(), " rel", err.max().item() / scale)
top = torch.topk(err, 8).indices
for i in top.tolist():
    A, B, Cc, o = conic[i].tolist()
    det = A * Cc - B * B
    tr = A + Cc
    import math
    lam1 = 0.5 * (tr + math.sqrt(max(tr * tr - 4 * det, 0.0)))
    lam2 = 0.5 * (tr - math.sqrt(max(tr * tr - 4 * det, 0.0)))
    print("gid %8d err %.3e  ours %s  ref %s  conic (%.4g %.4g %.4g) o %.3f  aniso %.1f radius %d tiles %d xy %s" % (
        i, err[i].item(), [round(x, 5) for x in m2o[i].tolist()], [round(x, 5) for x in m2r[i].tolist()], A, B, Cc, o,
        lam1 / max(lam2, 1e-30), int(fw[5][i]), int(geom[5][i]), [round(x, 2) for x in geom[1][i].tolist()]))
for k, nm in enumerate(helpers.GRAD_NAMES):
    a, b = ours[k], rb[k]
    if b.numel() == 0 or b.abs().max() == 0:
        continue
    print("%-14s l2 %.3e  max %.3e" % (nm, ((a - b).double().norm() / b.double().norm()).item(),
                                       (a - b).abs().max().item() / b.abs().max().item()))

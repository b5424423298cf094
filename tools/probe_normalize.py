#!/usr/bin/env python3
"""Which summation order does ATen's F.normalize use for a [P,4] row norm?  (run under gpurun, once per
FDGS_NORMALIZE_MODE = 0 / 1 / 2).  Renders the same raw parameters through the getter path (torch activations) and
through the raw-parameter entry (in-kernel activations) and counts differing words of the conditional covariance --
zero means the in-kernel exp / normalize are bit-identical to torch's."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("4d-gaussian-splatting_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import helpers  # noqa: E402
from raw_model import RawModel, Pipe  # noqa: E402
from gaussian_renderer import render  # noqa: E402
from gaussian_renderer.diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402

dev = "cuda:0"
cfg, cam, sc, st = helpers.build("mid", device=dev)
m = RawModel(sc, seed=3)
r = GaussianRasterizer(GaussianRasterizationSettings(**st))
with torch.no_grad():
    z = torch.zeros_like(m._xyz)
    a = r(means3D=m.get_xyz, means2D=z, opacities=m.get_opacity, shs=m.get_features, flow_2d=z[:, :2].contiguous(), ts=m.get_t,
          scales=m.get_scaling, scales_t=m.get_scaling_t, rotations=m.get_rotation, rotations_r=m.get_rotation_r)
    b = r.forward_raw(means3D=m._xyz, means2D=z, opacity_logits=m._opacity, features_dc=m._features_dc,
                      features_rest=m._features_rest, log_scales=m._scaling, rotations_raw=m._rotation,
                      flow_2d=z[:, :2].contiguous(), ts=m._t, log_scales_t=m._scaling_t, rotations_r_raw=m._rotation_r)
vis = a[1] > 0
diff = lambda x, y: int((x.contiguous().view(torch.int32) != y.contiguous().view(torch.int32)).sum())
print("FDGS_NORMALIZE_MODE=%s: cov3D words differing %d of %d, radii differing %d, image words differing %d of %d" % (
    os.environ.get("FDGS_NORMALIZE_MODE", "0"), diff(a[5][vis], b[5][vis]), int(vis.sum()) * 6, diff(a[1], b[1]),
    diff(a[0], b[0]), a[0].numel()))

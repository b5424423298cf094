#!/usr/bin/env python3
"""GPU bring-up / parity explorer (development tool, run under gpurun).

For a list of synthetic configurations runs the compiled reference rasterizer (oracle/_ref) and
the fdgs CUDA path on identical inputs and prints, field by field, how many elements differ
bitwise and the max-norm relative error; then quick CUDA-event timings of both.
Usage: python tools/first_light.py [--configs small,mid,cfg2,cfg3] [--golden DIR]
"""
import argparse
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "4d-gaussian-splatting_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fdgs  # noqa: E402
from fdgs import synth  # noqa: E402
import oracle_py  # noqa: E402

import helpers  # noqa: E402  (tests/helpers.py: the shared configuration table and argument builders)

CONFIGS = helpers.CONFIGS


def bitdiff(a, b):
    a = a.contiguous().view(-1)
    b = b.contiguous().view(-1)
    if a.dtype.is_floating_point:
        return int((a.view(torch.int32) != b.view(torch.int32)).sum())
    return int((a != b).sum())


def relerr(a, b):
    a = a.float()
    b = b.float()
    d = (a - b).abs().max().item()
    n = b.abs().max().item()
    return d / n if n > 0 else d


NO_REF = False
NO_TIMING = False

GRAD_NAMES = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dflows", "dL_dts",
              "dL_dscales", "dL_dscales_t", "dL_drot", "dL_drot_r"]


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def run_config(name, cfg, golden_dir=None):
    print("=" * 100)
    print("CONFIG", name, cfg, flush=True)
    dev = torch.device("cuda:0")
    cfg, cam, sc, st = helpers.build(cfg, device=dev)
    C = fdgs.ext()
    ref = oracle_py.ref_module() if (oracle_py.ref_available() and not NO_REF) else None
    P, W, H = cfg["P"], cfg["W"], cfg["H"]
    a = helpers.fwd_args(st, sc, cfg)
    mine = C.rasterize_gaussians(*a)
    torch.cuda.synchronize()
    print("mine: num_rendered", mine[0], "visible", int((mine[5] > 0).sum()), flush=True)
    grads = helpers.pixel_grads(cfg, device=dev)
    mine_b = C.rasterize_gaussians_backward(*helpers.bwd_args(st, sc, cfg, mine, grads))
    torch.cuda.synchronize()
    # colour-only backward (the training-loss case), so that profiles of a --no-timing run cover it
    e_ = torch.empty(0, device=dev)
    C.rasterize_gaussians_backward(*helpers.bwd_args(st, sc, cfg, mine, (grads[0], e_, e_, e_)))
    torch.cuda.synchronize()
    mg = C.debug_export_geom(mine[6], P)
    mb = C.debug_export_binning(mine[7], mine[8], mine[0], W, H)
    ncm = mb[2]
    print("mine: mean n_contrib %.2f  mean T %.4f  color mean %.4f" % (ncm.float().mean().item(), mine[4].mean().item(),
                                                                        mine[1].mean().item()), flush=True)
    if ref is not None:
        rf = ref.rasterize_gaussians(*a)
        torch.cuda.synchronize()
        rb = ref.rasterize_gaussians_backward(*helpers.bwd_args(st, sc, cfg, rf, grads))
        torch.cuda.synchronize()
        print("ref : num_rendered", rf[0], "visible", int((rf[5] > 0).sum()))
        vis = rf[5] > 0
        rg = oracle_py.ref_geom_views(rf[6], P)
        print("-- per-Gaussian (over reference-visible rows unless noted): #bitwise-different / maxrel")
        print("radii            diff=%d" % bitdiff(mine[5], rf[5]))
        print("tiles_touched    diff=%d" % bitdiff(mg[5], rg["tiles_touched"]))
        print("out_means3D      diff=%d (all rows)" % bitdiff(mine[10], rf[10]))
        for nm, m_, r_ in (("depths", mg[0], rg["depths"]), ("means2D", mg[1], rg["means2D"]),
                           ("conic_opacity", mg[2], rg["conic_opacity"]), ("rgb", mg[3], rg["rgb"]),
                           ("cov3D", mine[9], rf[9])):
            print("%-16s diff=%d  maxrel=%.3e" % (nm, bitdiff(m_[vis], r_[vis]), relerr(m_[vis], r_[vis])))
        mclamp = torch.stack([(mg[4] >> i) & 1 for i in range(3)], 1)
        print("clamped          diff=%d" % int((mclamp[vis] != rg["clamped"][vis]).sum()))
        if mine[0] == rf[0] and rf[0] > 0:
            rpl = oracle_py.ref_binning_point_list(rf[7], rf[0])
            print("point_list       diff=%d of %d" % (bitdiff(mb[0], rpl), rf[0]))
            ri = oracle_py.ref_image_views(rf[8], W * H)
            print("n_contrib        diff=%d" % bitdiff(ncm.view(-1), ri["n_contrib"]))
            nt = ((W + 15) // 16) * ((H + 15) // 16)
            print("ranges           diff=%d" % bitdiff(mb[1], ri["ranges"][:nt]))
        print("-- images: #bitwise-different / maxrel")
        for nm, i in (("color", 1), ("flow", 2), ("depth", 3), ("T", 4)):
            print("%-16s diff=%d  maxrel=%.3e" % (nm, bitdiff(mine[i], rf[i]), relerr(mine[i], rf[i])))
        rb2 = ref.rasterize_gaussians_backward(*helpers.bwd_args(st, sc, cfg, rf, grads))
        print("-- gradients: maxrel(mine vs ref)   [reference self-noise run-to-run]")
        for nm, m_, r_, r2 in zip(GRAD_NAMES, mine_b, rb, rb2):
            if r_.numel() == 0:
                continue
            print("%-16s %.3e   [%.3e]   max|ref|=%.3e" % (nm, relerr(m_, r_), relerr(r2, r_), r_.abs().max().item()))
    else:
        print("(reference module not available)")

    if golden_dir and ref is not None and P <= 10000:
        os.makedirs(golden_dir, exist_ok=True)
        rpl = oracle_py.ref_binning_point_list(rf[7], rf[0]) if rf[0] > 0 else torch.zeros(0, dtype=torch.int32)
        ri = oracle_py.ref_image_views(rf[8], W * H)
        nt = ((W + 15) // 16) * ((H + 15) // 16)
        np.savez_compressed(
            os.path.join(golden_dir, "golden_%s.npz" % name),
            cfg=np.array(repr(cfg)), num_rendered=np.int64(rf[0]), color=rf[1].cpu().numpy(), flow=rf[2].cpu().numpy(),
            depth=rf[3].cpu().numpy(), T=rf[4].cpu().numpy(), radii=rf[5].cpu().numpy(),
            covs3D=rf[9].cpu().numpy(), out_means3D=rf[10].cpu().numpy(),
            depths=rg["depths"].cpu().numpy(), means2D=rg["means2D"].cpu().numpy(),
            conic_opacity=rg["conic_opacity"].cpu().numpy(), rgb=rg["rgb"].cpu().numpy(),
            clamped=rg["clamped"].cpu().numpy(), tiles_touched=rg["tiles_touched"].cpu().numpy(),
            point_list=rpl.cpu().numpy(), n_contrib=ri["n_contrib"].cpu().numpy(),
            ranges=ri["ranges"][:nt].cpu().numpy(),
            **{("grad_" + n): t.cpu().numpy() for n, t in zip(GRAD_NAMES, rb)})
        print("golden written", flush=True)

    # timings
    if NO_TIMING:
        return
    try:
        t_mf = timeit(lambda: C.rasterize_gaussians(*a))
        t_mb = timeit(lambda: C.rasterize_gaussians_backward(*helpers.bwd_args(st, sc, cfg, mine, grads)))
        e = torch.empty(0, device=dev)
        lean = (grads[0], e, e, e)   # colour-image loss only (the reference's training loss)
        t_ml = timeit(lambda: C.rasterize_gaussians_backward(*helpers.bwd_args(st, sc, cfg, mine, lean)))
        line = "TIMING %s mine fwd %.3f ms bwd %.3f ms (colour-only bwd %.3f ms)" % (name, t_mf, t_mb, t_ml)
        if ref is not None:
            t_rf = timeit(lambda: ref.rasterize_gaussians(*a))
            t_rb = timeit(lambda: ref.rasterize_gaussians_backward(*helpers.bwd_args(st, sc, cfg, rf, grads)))
            line += " | ref fwd %.3f ms bwd %.3f ms | speedup fwd %.2fx bwd %.2fx total %.2fx" % (
                t_rf, t_rb, t_rf / t_mf, t_rb / t_mb, (t_rf + t_rb) / (t_mf + t_mb))
        print(line, flush=True)
    except Exception:
        traceback.print_exc()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="tiny,small,flowbg,negfov,ragged,sh3d,dim3,norot4d,deg1,m16,prefilter,mid,cfg2,cfg3")
    ap.add_argument("--golden", default=None)
    ap.add_argument("--no-ref", action="store_true", help="skip the reference (e.g. under ncu)")
    ap.add_argument("--no-timing", action="store_true")
    args = ap.parse_args()
    global NO_REF, NO_TIMING
    NO_REF, NO_TIMING = args.no_ref, args.no_timing
    print(torch.cuda.get_device_name(0), torch.version.cuda, flush=True)
    for name in args.configs.split(","):
        try:
            run_config(name, CONFIGS[name], args.golden)
        except Exception:
            traceback.print_exc()
            torch.cuda.synchronize()
        sys.stdout.flush()


if __name__ == "__main__":
    main()

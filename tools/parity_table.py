#!/usr/bin/env python3
"""Per-gradient error-vs-noise table (run under gpurun; writes markdown to stdout).

For each configuration: our gradients against the MEAN of K runs of the compiled, unmodified reference
(oracle/_ref), next to the reference's own run-to-run spread (unordered fp32 atomics), in the L2 and the
max-norm sense; forward tensors bit-compared; PSNR of the image.  For sizes the CPU oracle finishes in
seconds the deterministic oracle (serial fp32 sums) is a third column: ours-vs-oracle and reference-vs-oracle.
Usage: python tools/parity_table.py [--configs mid,q250k,cfg2,cfg3] [--reruns 8] [--colour-only]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("4d-gaussian-splatting_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))

import fdgs  # noqa: E402
import helpers  # noqa: E402
import oracle_py  # noqa: E402


def rel(a, m):
    nm, sc = m.norm().item(), m.abs().max().item()
    if nm == 0 or sc == 0:
        return 0.0, 0.0
    d = a.double() - m
    return d.norm().item() / nm, d.abs().max().item() / sc


def table(name, K, colour_only, oracle_limit):
    dev = "cuda:0"
    C = fdgs.ext()
    ref = oracle_py.ref_module()
    cfg, cam, sc, st = helpers.build(name, device=dev)
    gc, gd, ga, gf = helpers.pixel_grads(cfg, device=dev)
    e = torch.empty(0, device=dev)
    fw = C.rasterize_gaussians(*helpers.fwd_args(st, sc, cfg))
    ours = C.rasterize_gaussians_backward(*helpers.bwd_args(st, sc, cfg, fw, (gc, e, e, e) if colour_only else (gc, gd, ga, gf)))
    rf = ref.rasterize_gaussians(*helpers.fwd_args(st, sc, cfg))
    up = (gc, 0 * gd, 0 * ga, 0 * gf) if colour_only else (gc, gd, ga, gf)
    runs = [ref.rasterize_gaussians_backward(*helpers.bwd_args(st, sc, cfg, rf, up)) for _ in range(K)]
    torch.cuda.synchronize()
    eq = lambda a, b: int((a.contiguous().view(torch.int32) != b.contiguous().view(torch.int32)).sum())
    mse = ((fw[1].double() - rf[1].double()) ** 2).mean().item()
    psnr = float("inf") if mse == 0 else 10 * np.log10(1.0 / mse)
    print("\n### %s  (P=%d, %dx%d, %s upstream gradient, K=%d reference reruns)\n" %
          (name, cfg["P"], cfg["W"], cfg["H"], "colour-only" if colour_only else "colour+depth+alpha+flow", K))
    print("forward: num_rendered %d vs %d; differing words: radii %d, image %d, depth %d, flow %d, T %d, out_means3D %d; "
          "PSNR(image) %s dB\n" % (fw[0], rf[0], eq(fw[5], rf[5]), eq(fw[1], rf[1]), eq(fw[3], rf[3]), eq(fw[2], rf[2]),
                                   eq(fw[4], rf[4]), eq(fw[10], rf[10]), "inf (bit-identical)" if mse == 0 else "%.1f" % psnr))
    orc = None
    if cfg["P"] <= oracle_limit:
        cfg_c, _, sc_c, st_c = helpers.build(name)
        inp = helpers.oracle_inputs(st_c, sc_c, cfg_c)
        f = oracle_py.forward(inp)
        z = lambda t: np.zeros_like(t.cpu().numpy())
        g = oracle_py.backward(inp, f, gc.cpu().numpy(), *([z(gd), z(ga), z(gf)] if colour_only else
                                                             [gd.cpu().numpy(), ga.cpu().numpy(), gf.cpu().numpy()]))
        orc = [torch.from_numpy(np.ascontiguousarray(g[k])).to(dev) for k in helpers.ORACLE_GRAD_KEYS]
    hdr = "| gradient | ours vs ref-mean L2 | ref spread L2 (1 run) | ours vs ref-mean max | ref spread max |"
    if orc:
        hdr += " ours vs oracle L2 | ref-mean vs oracle L2 |"
    print(hdr)
    print("|" + "---|" * (hdr.count("|") - 1))
    for k, gname in enumerate(helpers.GRAD_NAMES):
        rs = [r[k] for r in runs]
        if rs[0].numel() == 0:
            continue
        m = torch.stack([r.double() for r in rs]).mean(0)
        l2, mx = rel(ours[k], m)
        sp = [rel(r, m) for r in rs]
        sl2 = float(np.sqrt(np.mean([s[0] ** 2 for s in sp])))
        smx = max(s[1] for s in sp)
        row = "| %s | %.2e | %.2e | %.2e | %.2e |" % (gname, l2, sl2, mx, smx)
        if orc:
            o = orc[k].double().reshape(m.shape)
            row += " %.2e | %.2e |" % (rel(ours[k], o)[0], rel(m, o)[0])
        print(row)
    sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="mid,mid_rotcam,mid_dur10,q250k,cfg5,cfg2,cfg3")
    ap.add_argument("--reruns", type=int, default=8)
    ap.add_argument("--oracle-limit", type=int, default=300000)
    args = ap.parse_args()
    print("# Parity table: gradients vs the compiled reference (mean of K runs) and vs the deterministic CPU oracle\n")
    print("GPU: %s.  L2 = ||a-b||2/||b||2, max = ||a-b||inf/||b||inf.  'ref spread' = deviation of single reference runs "
          "from the mean of K (its unordered-atomics noise); the mean itself is known to spread/sqrt(K)." % torch.cuda.get_device_name(0))
    for name in args.configs.split(","):
        for colour_only in (True, False):
            table(name, args.reruns, colour_only, args.oracle_limit)


if __name__ == "__main__":
    main()

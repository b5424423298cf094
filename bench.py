#!/usr/bin/env python3
"""bench.py -- forward+backward Mpixels/s of the differentiable 4D Gaussian rasterizer.

Metric (BASELINE.json): fwd+bwd Mpixels/s @1352x1014 with 2M 4D Gaussians (SH degree 3, temporal
degree 2, 48 coefficients), one view per step per GPU.  A "step" = one render + one backward of a
view of the synthetic scene of SURVEY.md section 8(d), through the reference-facing Python API
(GaussianRasterizer + autograd over the C-ABI library).

  value   inputs resident in HBM, CUDA events, K steps after W warm-ups, max over ranks
  e2e     same call with HOST buffers: per step the camera matrices and the upstream gradient image
          are copied host->device from pinned memory (the image on a copy stream, overlapping the
          forward, both arms alike) and the loss is read back device->host (asynchronously into pinned
          memory, collected one step later and before the clock stops, so the launch queue never drains)
  N > 1   one view per rank (weak scaling), replicated Gaussians, SUM all-reduce (NCCL) of the
          per-Gaussian parameter gradients (rows of the union of the ranks' rendered Gaussians, one flat
          buffer) + the densification statistics inside the step
  --impl reference   the UNMODIFIED reference CUDA rasterizer (oracle/_ref, built from
          /root/reference by oracle/build_ref.py) on the same workload, same metric; if that .so did
          not travel, the CPU oracle port on a bounded sample.  Rank 0 only.

One JSON line on stdout (rank 0).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "4d-gaussian-splatting_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

WORKLOADS = {
    # name: (P, W, H, seed)  -- BASELINE.json configs[2] is the headline
    "cfg3": dict(P=2_000_000, W=1352, H=1014, seed=1237, desc="2M 4D Gaussians, 1352x1014, fwd+bwd, SH degree 3"),
    "cfg2": dict(P=500_000, W=1352, H=1014, seed=1236, desc="500k 4D Gaussians, 1352x1014"),
    "mid": dict(P=100_000, W=640, H=480, seed=1236, desc="100k 4D Gaussians, 640x480 (smoke)"),
}


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 300] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def make_view(W, H, rank):
    """View of rank `rank`: small yaw about the scene centre + its own timestamp (rank 0 = the parity view)."""
    from fdgs import synth
    if rank == 0:
        return synth.make_camera(W, H, timestamp=0.5)
    ang = math.radians(1.5 * ((rank + 1) // 2) * (1 if rank % 2 else -1))
    c, s = math.cos(ang), math.sin(ang)
    R = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
    centre = torch.tensor([0.0, 0.0, 6.0])
    T = centre - R.t() @ centre   # rotate about the scene centre
    return synth.make_camera(W, H, timestamp=0.5 + 0.01 * rank, R=R, T=T)


class Workload:
    def __init__(self, name, device, rank):
        from fdgs import synth
        w = WORKLOADS[name]
        self.name, self.P, self.W, self.H = name, w["P"], w["W"], w["H"]
        self.cam = make_view(self.W, self.H, rank)
        base_cam = synth.make_camera(self.W, self.H)
        self.scene_cpu = synth.make_scene(self.P, base_cam, w["seed"])       # identical on every rank
        self.scene = self.scene_cpu.to(device)
        self.device = device
        g = torch.Generator().manual_seed(w["seed"] + 999 + rank)
        self.G_host = torch.randn(3, self.H, self.W, generator=g).pin_memory() if device != "cpu" else torch.randn(3, self.H, self.W, generator=g)
        self.params = {k: v.clone().requires_grad_(True) for k, v in self.scene.tensors().items() if k != "flow_2d"}
        self.settings = synth.raster_settings(self.cam, self.scene_cpu, device=device)
        # pinned host copies of the per-step inputs (camera) for the e2e leg
        self.cam_host = {k: self.settings[k].cpu().pin_memory() for k in ("viewmatrix", "projmatrix", "campos")} if device != "cpu" else {}


class Runner:
    """One render+backward step through a rasterizer API (ours or the reference's)."""

    def __init__(self, wl: Workload, impl: str):
        self.wl = wl
        self.impl = impl
        if impl == "ours":
            from gaussian_renderer import GaussianRasterizationSettings, GaussianRasterizer
            self.Settings, self.Rasterizer = GaussianRasterizationSettings, GaussianRasterizer
        else:
            import ref_api
            self.ref_api = ref_api
        self.G_dev = wl.G_host.to(wl.device)
        self.copy_stream = torch.cuda.Stream(device=wl.device)
        self.loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.loss_ready = torch.cuda.Event()
        self.pending = False
        self.last = None

    def _raster(self, settings):
        p, sc = self.wl.params, self.wl.scene
        means2D = torch.zeros_like(p["means3D"], requires_grad=True)
        if self.impl == "ours":
            out = self.Rasterizer(self.Settings(**settings))(
                means3D=p["means3D"], means2D=means2D, opacities=p["opacities"], shs=p["shs"], flow_2d=sc.flow_2d,
                ts=p["ts"], scales=p["scales"], scales_t=p["scales_t"], rotations=p["rotations"],
                rotations_r=p["rotations_r"])
        else:
            out = self.ref_api.rasterize(settings, p["means3D"], means2D, p["opacities"], p["shs"], sc.flow_2d, p["ts"],
                                         p["scales"], p["scales_t"], p["rotations"], p["rotations_r"])
        return out, means2D

    def step_resident(self):
        for v in self.wl.params.values():
            v.grad = None
        (color, radii, depth, alpha, flow, covs), means2D = self._raster(self.wl.settings)
        loss = (color * self.G_dev).sum()
        loss.backward()
        self.last = (loss, radii, means2D)
        return loss

    def step_e2e(self):
        """Host buffers in, host scalar out: H2D of the camera + upstream gradient image, D2H of the loss."""
        wl = self.wl
        for v in wl.params.values():
            v.grad = None
        st = dict(wl.settings)
        for k, h in wl.cam_host.items():
            st[k] = h.to(wl.device, non_blocking=True)
        # the upstream gradient image (the stand-in for the ground-truth image of a training step) is
        # uploaded on a copy stream while the forward runs, like a data loader prefetching the next view
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            G = wl.G_host.to(wl.device, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        (color, radii, depth, alpha, flow, covs), means2D = self._raster(st)
        cur.wait_event(ready)
        G.record_stream(cur)
        loss = (color * G).sum()
        loss.backward()
        self.last = (loss, radii, means2D)
        # device->host read of the step's result: an asynchronous copy into pinned memory, collected at the start
        # of the next step (and by drain() before the clock stops) -- the way a training loop logs its loss
        # without stalling the launch queue.  Every step's result is read inside the timed region.
        prev = self.drain()
        self.loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        self.loss_ready.record(cur)
        self.pending = True
        return prev

    def drain(self):
        """Wait for and return the most recent step's loss (None if nothing is pending)."""
        if not getattr(self, "pending", False):
            return None
        self.loss_ready.synchronize()
        self.pending = False
        return float(self.loss_host[0])

    def h2d_bytes(self):
        return int(self.wl.G_host.numel() * 4 + sum(h.numel() * 4 for h in self.wl.cam_host.values()))

    def grads(self):
        return [v.grad for v in self.wl.params.values()]


def timed(fn, steps, warmup, device, world, finish=None):
    for _ in range(warmup):
        fn()
    if finish is not None:
        finish()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    if finish is not None:
        finish()      # e.g. collect the last step's host read-back before the clock stops
    e1.record()
    torch.cuda.synchronize(device)
    wall = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps, wall / steps * 1e3


def cpu_port_baseline(max_seconds=40.0):
    """The CPU oracle (port of the reference algorithm) on a bounded sample: a 1/8-scale replica of the
    workload with the same Gaussians-per-pixel density (250k Gaussians at 478x358), all stages,
    forward + backward, every host core the oracle can use (OpenMP in preprocess / forward blend)."""
    import oracle_py
    import helpers
    cfg = dict(P=250_000, W=478, H=358, seed=1237)
    cfg, cam, sc, st = helpers.build(cfg)
    inp = helpers.oracle_inputs(st, sc, cfg)
    gc, gd, ga, gf = helpers.pixel_grads(cfg)
    z = lambda t: np.zeros_like(t.numpy())
    t0 = time.perf_counter()
    f = oracle_py.forward(inp)
    t1 = time.perf_counter()
    oracle_py.backward(inp, f, gc.numpy(), z(gd), z(ga), z(gf))
    t2 = time.perf_counter()
    mpix = cfg["W"] * cfg["H"] / 1e6
    return {"value": mpix / (t2 - t0), "unit": "Mpixels/s", "cores": oracle_py.num_threads(), "kind": "port",
            "sample": "1/8-scale replica of the workload (250k Gaussians, 478x358, same density), all stages, fwd %.2fs + bwd %.2fs"
                      % (t1 - t0, t2 - t1)}


def python_preprocess_baseline(wl: Workload, max_points=2_000_000):
    """The reference's Python preprocess branches (compute_cov3D_python + convert_SHs_python,
    gaussian_renderer/__init__.py:73-81,98-111) restated device-agnostically and timed on the host cores."""
    from gaussian_renderer import pyprep
    sc = wl.scene_cpu
    n = min(sc.P, max_points)
    xyzt = torch.cat([sc.scales[:n], sc.scales_t[:n]], 1)
    t0 = time.perf_counter()
    with torch.no_grad():
        pyprep.python_preprocess(sc.means3D[:n], sc.ts[:n], xyzt, sc.rotations[:n], sc.rotations_r[:n], sc.opacities[:n],
                                 sc.shs[:n], wl.cam.camera_center, wl.cam.timestamp, sc.time_duration, 3, 2)
    dt = time.perf_counter() - t0
    return {"ms": dt * 1e3, "gaussians": n, "threads": torch.get_num_threads(), "cores": os.cpu_count()}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0                      # the reference arm runs on rank 0 alone
    ref_solo = args.impl == "reference"
    if world > 1 and not ref_solo:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    eff_world = 1 if ref_solo else world

    import oracle_py
    if args.impl == "reference" and not (torch.cuda.is_available() and oracle_py.ref_available()):
        # the compiled reference did not travel (or no GPU): time the CPU port on its bounded sample
        cb = cpu_port_baseline()
        line = {"impl": "reference", "metric": "fwd+bwd Mpixels/s @1352x1014, 2M 4D Gaussians", "value": cb["value"],
                "unit": "Mpixels/s", "n_gpus": 0, "steps": 1, "warmup": 0, "ms_per_step": None, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "note": "oracle/_ref not available: CPU port of the reference"},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "Mpixels/s", "h2d_bytes_per_step": 0,
                                            "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    assert torch.cuda.is_available(), "bench.py needs a GPU (the product has no CPU path)"
    device = "cuda:%d" % local
    torch.cuda.set_device(device)
    wl = Workload(args.workload, device, rank if not ref_solo else 0)
    runner = Runner(wl, "ours" if args.impl == "ours" else "ref")
    mpix = wl.W * wl.H / 1e6

    if eff_world > 1:
        from fdgs.dist import allreduce_gradients, ViewBatchStats

        dense = os.environ.get("FDGS_DENSE_ALLREDUCE") is not None

        def sync_grads():
            stats = ViewBatchStats(wl.P, device)
            stats.add_view(runner.last[2].grad, runner.last[1])
            stats.reduce()
            # rows of Gaussians no rank rendered are zero everywhere: exchange the union's rows only
            allreduce_gradients(runner.grads(), union_visible=None if dense else stats.max_radii > 0)
    else:
        def sync_grads():
            return None

    def step_value():
        runner.step_resident()
        sync_grads()

    def step_e2e():
        runner.step_e2e()
        sync_grads()

    # ---- value: resident inputs -------------------------------------------------------------------
    import fdgs
    launches0 = fdgs.launch_count() if args.impl == "ours" else 0
    sampler = ClockSampler(local)
    sampler.start()
    ms_step, wall_ms = timed(step_value, args.steps, args.warmup, device, eff_world)
    clocks = sampler.stop()
    launches = (fdgs.launch_count() - launches0) if args.impl == "ours" else None
    value = eff_world * mpix / (ms_step * 1e-3)

    # ---- e2e: host buffers ---------------------------------------------------------------------------
    ms_e2e, _ = timed(step_e2e, args.steps, max(3, args.warmup // 2), device, eff_world, finish=runner.drain)
    e2e = {"value": eff_world * mpix / (ms_e2e * 1e-3), "unit": "Mpixels/s", "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": runner.h2d_bytes(), "d2h_bytes_per_step": 4}

    # ---- scene statistics + per-stage times (separate short run, not part of the numbers above) -------
    loss, radii, _ = runner.last
    P_vis = int((radii > 0).sum().item())
    stats = {"P": wl.P, "P_vis": P_vis, "N_pixels": wl.W * wl.H}
    roofline = None
    stage_ms = None
    if args.impl == "ours":
        fdgs.profile_enable(True)
        for _ in range(3):
            runner.step_resident()
        torch.cuda.synchronize(device)
        prof = fdgs.profile_read()
        fdgs.profile_enable(False)
        stage_ms = {k: (v[0] / max(v[1], 1)) for k, v in prof.items() if v[1] > 0}
        C = fdgs.ext()
        with torch.no_grad():
            import helpers
            fw = C.rasterize_gaussians(*helpers.fwd_args(wl.settings, wl.scene, {}))
            ncontrib = C.debug_export_binning(fw[7], fw[8], fw[0], wl.W, wl.H)[2]
            R = int(fw[0])
            stats.update(R=R, mean_n_contrib=float(ncontrib.float().mean().item()))
            del fw
        N = wl.W * wl.H
        peak, peak_src = measured_peak()
        # algorithmic bytes: blend backward = 52 B per instance + 36 B per pixel
        # + 52 B per visible Gaussian (its blend-gradient row, written once)
        dom = max(stage_ms, key=stage_ms.get)
        alg = {"blend_bwd": 52.0 * R + 36.0 * N + 52.0 * P_vis, "blend_fwd": 28.0 * R + 32.0 * N,
               "preprocess_bwd": 1376.0 * P_vis + 88.0 * wl.P, "preprocess_fwd": 84.0 * wl.P + 655.0 * P_vis,
               "tile_sort_pack": 80.0 * R, "bin_scatter": 8.0 * R + 4.0 * wl.P + 32.0 * P_vis}.get(dom, 0.0)
        ach = alg / (stage_ms[dom] * 1e-3) / 1e9
        b_fwd = 84.0 * wl.P + 719.0 * P_vis + 88.0 * R + 32.0 * N
        b_bwd = 52.0 * R + 36.0 * N + 1428.0 * P_vis
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of that kernel, per launch, from the committed ncu capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[dom]
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "algorithmic_bytes": alg, "peak_source": peak_src,
                    "note": "dominant kernel is instruction-issue / shared-memory bound, not HBM bound (profiles/r01_blend_ncu_v11.md)",
                    "frame": {"algorithmic_bytes": b_fwd + b_bwd, "achieved": (b_fwd + b_bwd) / (ms_step * 1e-3) / 1e9,
                              "frac": (b_fwd + b_bwd) / (ms_step * 1e-3) / 1e9 / peak}}

    # ---- CPU baselines (rank 0, N = 1 only) -------------------------------------------------------------
    cpu_baseline = None
    py_pre = None
    if rank == 0 and eff_world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_port_baseline()
        py_pre = python_preprocess_baseline(wl)

    if rank == 0:
        line = {
            "metric": "fwd+bwd Mpixels/s @1352x1014, 2M 4D Gaussians" if args.workload == "cfg3" else "fwd+bwd Mpixels/s",
            "value": value, "unit": "Mpixels/s", "n_gpus": eff_world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.workload, WORKLOADS[args.workload]["desc"]),
                       "views_per_step": eff_world, "parallelism": "dp%d over views" % eff_world,
                       "l2": "inputs exceed L2 (SH rows alone are %.2f GB)" % (wl.P * 576 / 1e9),
                       "api": "GaussianRasterizer + autograd (reference-facing Python API over the C-ABI)"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "stats": stats,
            "wall_ms_per_step": wall_ms,
        }
        if args.impl == "reference":
            line["impl"] = "reference"
            line["config"]["reference"] = "unmodified reference CUDA rasterizer (oracle/_ref/ref_rasterizer.so, sm_100a)"
            line["gpu_launches"] = None
            line["cpu_baseline"] = cpu_baseline or {"value": None, "kind": "reference", "cores": 0,
                                                    "sample": "reference arm ran its CUDA path on the GPU"}
        else:
            line["roofline"] = roofline
            line["stage_ms"] = stage_ms
            line["cpu_baseline"] = cpu_baseline
        if py_pre is not None:
            line["python_preprocess_host"] = py_pre
        print(json.dumps(line))
    if world > 1 and not ref_solo:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

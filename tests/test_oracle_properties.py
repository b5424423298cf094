"""CPU: known-answer and consistency tests of the oracle that do not need the reference:
closed-form single-Gaussian profile, temporal slice identities, depth ordering, agreement with
the reference's *second* implementation of the same maths (the PyTorch "python preprocess" path,
restated in gaussian_renderer/pyprep.py), and linearity / locality properties of the backward."""
import math

import numpy as np
import torch

import helpers
import oracle_py
from fdgs import synth
from gaussian_renderer import pyprep


def _scene(P, W=64, H=64, **kw):
    cam = synth.make_camera(W, H)
    sc = synth.make_scene(P, cam, 1, **kw)
    return cam, sc


def _run(cam, sc, bg=None):
    st = synth.raster_settings(cam, sc, bg=bg)
    four_d = sc.gaussian_dim == 4
    inp = oracle_py.OracleInputs(st, sc.means3D, sc.opacities, shs=sc.shs, flow_2d=sc.flow_2d, ts=sc.ts if four_d else None,
                                 scales=sc.scales, scales_t=sc.scales_t if four_d else None, rotations=sc.rotations,
                                 rotations_r=sc.rotations_r if sc.rot_4d else None)
    return inp, oracle_py.forward(inp)


def test_single_isotropic_gaussian_closed_form():
    W = H = 64
    cam, sc = _scene(1, W, H, gaussian_dim=3, rot_4d=False, force_sh_3d=True, sh_degree=0, M=16)
    z, s, o = 4.0, 0.15, 0.8
    sc.means3D[:] = torch.tensor([[0.0, 0.0, z]])
    sc.scales[:] = s
    sc.rotations[:] = torch.tensor([[1.0, 0.0, 0.0, 0.0]])
    sc.opacities[:] = o
    sc.shs.zero_()
    sc.shs[0, 0, :] = torch.tensor([0.2, 0.5, 0.8]) / 0.28209479177387814   # rgb = dc*C0 + 0.5
    inp, f = _run(cam, sc)
    fl = 0.54 * W
    var = (fl * s / z) ** 2 + 0.3                      # EWA low-pass (forward.cu:234-235)
    cx = ((0.0 + 1.0) * W - 1.0) * 0.5                 # ndc2Pix
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    r2 = (xs - cx) ** 2 + (ys - cx) ** 2
    alpha = np.minimum(0.99, o * np.exp(-0.5 * r2 / var))
    alpha[alpha < 1.0 / 255.0] = 0.0
    radius = math.ceil(3.0 * math.sqrt(var))
    assert f["radii"][0] == radius
    # outside the tiles the Gaussian touches nothing is blended; inside, the profile is the closed form
    got = 1.0 - f["final_T"]
    x0, x1 = int((cx - radius) // 16) * 16, (int((cx + radius + 15) // 16)) * 16
    mask = np.zeros_like(alpha, bool)
    mask[max(0, x0):x1, max(0, x0):x1] = True
    np.testing.assert_allclose(got[mask], alpha[mask], atol=2e-6)
    assert np.all(got[~mask] == 0)
    expect_rgb = np.array([0.7, 1.0, 1.3])
    np.testing.assert_allclose(f["color"][:, 32, 32], alpha[32, 32] * expect_rgb, rtol=1e-5)
    np.testing.assert_allclose(f["depth"][0, 32, 32], alpha[32, 32] * z, rtol=1e-5)


def test_time_slice_identities():
    cam, sc = _scene(500)
    # dt = 0: marginal = 1, no mean shift, opacity untouched
    sc.ts[:] = cam.timestamp
    inp, f = _run(cam, sc)
    vis = f["radii"] > 0
    assert vis.sum() > 100
    np.testing.assert_array_equal(f["out_means3D"], sc.means3D.numpy())
    np.testing.assert_allclose(f["conic_opacity"][vis, 3], sc.opacities.numpy()[vis, 0], rtol=1e-7)
    # identity quaternions: Sigma is diagonal, the conditional covariance is diag(s^2), no shift for any dt
    cam, sc = _scene(500)
    sc.rotations[:] = torch.tensor([1.0, 0, 0, 0])
    sc.rotations_r[:] = torch.tensor([1.0, 0, 0, 0])
    inp, f = _run(cam, sc)
    vis = f["radii"] > 0
    s2 = (sc.scales.numpy() ** 2)[vis]
    np.testing.assert_allclose(f["cov3D"][vis][:, [0, 3, 5]], s2, rtol=1e-6)
    assert np.all(f["cov3D"][vis][:, [1, 2, 4]] == 0)
    np.testing.assert_array_equal(f["out_means3D"][vis], sc.means3D.numpy()[vis])
    # far-away timestamps are culled by the marginal test (forward.cu:334)
    dt = np.abs(sc.ts.numpy()[:, 0] - cam.timestamp)
    st2 = sc.scales_t.numpy()[:, 0] ** 2
    culled = np.exp(-0.5 * dt * dt / st2) <= 0.049
    assert culled.sum() > 10 and np.all(f["radii"][culled] == 0)


def test_depth_order_and_alpha_compositing():
    W = H = 32
    cam, sc = _scene(2, W, H, gaussian_dim=3, rot_4d=False, force_sh_3d=True, sh_degree=0, M=16)
    sc.means3D[:] = torch.tensor([[0.0, 0.0, 6.0], [0.0, 0.0, 3.0]])   # index 0 is BEHIND index 1
    sc.scales[:] = torch.tensor([[0.9] * 3, [0.45] * 3])               # same screen footprint
    sc.rotations[:] = torch.tensor([1.0, 0, 0, 0])
    sc.opacities[:] = torch.tensor([[0.9], [0.5]])
    sc.shs.zero_()
    sc.shs[0, 0, :] = (torch.tensor([1.0, 0.0, 0.0]) - 0.5) / 0.28209479177387814
    sc.shs[1, 0, :] = (torch.tensor([0.0, 1.0, 0.0]) - 0.5) / 0.28209479177387814
    inp, f = _run(cam, sc)
    first_tile = f["point_list"][f["ranges"][0, 0]:f["ranges"][0, 1]]
    assert list(first_tile) == [1, 0]                                   # sorted front to back
    py, px = 15, 15
    co = f["conic_opacity"]
    d = lambda i: (f["means2D"][i, 0] - px, f["means2D"][i, 1] - py)
    al = []
    for i in (1, 0):
        dx, dy = d(i)
        al.append(min(0.99, co[i, 3] * math.exp(-0.5 * (co[i, 0] * dx * dx + co[i, 2] * dy * dy) - co[i, 1] * dx * dy)))
    expect = np.array([0.0, 1.0, 0.0]) * al[0] + np.array([1.0, 0.0, 0.0]) * al[1] * (1 - al[0])
    np.testing.assert_allclose(f["color"][:, py, px], expect, rtol=1e-5, atol=1e-6)
    assert f["n_contrib"][py, px] == 2


def test_agrees_with_python_preprocess_path():
    """The reference's own second implementation (gaussian_model.py:34-47,230-251, sh_utils.py:115-223)
    must agree with the CUDA-path restatement up to the documented quirks."""
    cam, sc = _scene(4000, 128, 128)
    inp, f = _run(cam, sc)
    vis = f["radii"] > 0
    xyzt = torch.cat([sc.scales, sc.scales_t], 1)
    dt = cam.timestamp - sc.ts
    cov, delta = pyprep.conditional_covariance_and_offset(xyzt, 1.0, sc.rotations, sc.rotations_r, dt)
    marg = pyprep.marginal_t(xyzt, 1.0, sc.rotations, sc.rotations_r, sc.ts, cam.timestamp)
    assert helpers.max_rel(f["cov3D"][vis], cov.numpy()[vis]) < 2e-6
    np.testing.assert_allclose(f["out_means3D"][vis], (sc.means3D + delta).numpy()[vis], atol=2e-6)
    np.testing.assert_allclose(f["conic_opacity"][vis, 3], (sc.opacities * marg).numpy()[vis, 0], rtol=2e-6)
    assert np.all((marg.numpy()[:, 0] > 0.05)[vis])
    # colour: CUDA evaluates SH at the UNSHIFTED mean (quirk, forward.cu:480,482) -> use that direction
    dirs = sc.means3D - cam.camera_center[None]
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    rgb = pyprep.eval_shfs_4d(3, 2, sc.shs.transpose(1, 2), dirs, sc.ts - cam.timestamp, sc.time_duration)
    rgb = torch.clamp_min(rgb + 0.5, 0.0).numpy()
    np.testing.assert_allclose(f["rgb"][vis], rgb[vis], atol=3e-6)


def test_backward_linearity_and_locality():
    """Properties of the backward the domain offers without a reference: it is linear in the upstream
    pixel gradients, zero upstream gradients give zero parameter gradients, and Gaussians that were
    not rendered receive exactly zero.  (A finite-difference check is NOT meaningful here: the hard
    alpha < 1/255 cut-off makes the rendered image discontinuous in the parameters, and the reference's
    analytic gradient ignores those boundary terms by design.)"""
    cam, sc = _scene(3000, 96, 80, flow=True)
    inp, f = _run(cam, sc, bg=torch.tensor([0.2, 0.4, 0.6]))
    rng = np.random.RandomState(0)
    H, W = 80, 96
    G1 = [rng.randn(3, H, W).astype(np.float32), rng.randn(1, H, W).astype(np.float32),
          rng.randn(1, H, W).astype(np.float32), rng.randn(2, H, W).astype(np.float32)]
    G2 = [rng.randn(*g.shape).astype(np.float32) for g in G1]
    ga = oracle_py.backward(inp, f, *G1)
    gb = oracle_py.backward(inp, f, *G2)
    gab = oracle_py.backward(inp, f, *[2.0 * a - 0.5 * b for a, b in zip(G1, G2)])
    gz = oracle_py.backward(inp, f, *[np.zeros_like(a) for a in G1])
    invisible = f["radii"] <= 0
    assert invisible.sum() > 100
    for k in ga:
        lin = 2.0 * ga[k] - 0.5 * gb[k]
        assert helpers.l2_rel(gab[k], lin) < 1e-4, k
        assert not np.any(gz[k]), k
        assert not np.any(ga[k][invisible]), k

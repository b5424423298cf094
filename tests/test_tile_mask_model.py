"""The tile mask of the binning (preprocess_fwd.cu) restated in numpy and checked by brute force.

The CUDA kernel lists a Gaussian only in the tiles its alpha >= 1/255 ellipse { d : 0.5 d^T Q d <= qc } can reach:
bounding box of the ellipse, then per tile row the x-interval of the ellipse inside the row's y-band.  This file
re-implements exactly that arithmetic (float32, same margins) and verifies on random ellipses that it is CONSERVATIVE:
every tile that contains a pixel centre with q <= qc is kept.  (The kernel itself is covered on the GPU by the
bit-identical-image tests and by test_tile_lists_sorted_and_complete; this test pins the algorithm.)"""
import numpy as np

TILE = 16
f32 = np.float32


def tile_mask_model(px, py, A, B, C, qc, rect):
    """rect = (rx0, ry0, rx1, ry1): the reference's tile rectangle.  Returns the set of (tx, ty) tiles kept."""
    px, py, A, B, C = f32(px), f32(py), f32(A), f32(B), f32(C)
    rx0, ry0, rx1, ry1 = rect
    det = f32(A * C - B * B)
    pmin = f32(-(qc + 0.02))
    if not (det > f32(2.0e-3) * A * C):
        return {(tx, ty) for ty in range(ry0, ry1) for tx in range(rx0, rx1)}      # thin: the reference's rectangle
    q2 = f32(-2.0) * pmin
    q2_det = f32(q2 / det)
    ex = f32(np.sqrt(q2_det * C) * f32(1.0001) + f32(0.01))
    ey = f32(np.sqrt(q2_det * A) * f32(1.0001) + f32(0.01))
    tx0 = max(rx0, int(np.floor((px - ex) / TILE)))
    tx1 = min(rx1, int(np.floor((px + ex) / TILE)) + 1)
    ty0 = max(ry0, int(np.floor((py - ey) / TILE)))
    ty1 = min(ry1, int(np.floor((py + ey) / TILE)) + 1)
    if tx1 <= tx0 or ty1 <= ty0:
        return set()
    tw, tn = tx1 - tx0, (tx1 - tx0) * (ty1 - ty0)
    if not (tn <= 32 and tw > 1 and (ty1 - ty0) > 1):
        return {(tx, ty) for ty in range(ty0, ty1) for tx in range(tx0, tx1)}
    keep = set()
    sbc = f32(-B / C)
    dy_hi, dy_lo = f32(sbc * ex), f32(-sbc * ex)
    for cy in range(ty0, ty1):
        a0 = f32(cy * TILE) - py
        b0 = f32(a0 + (TILE - 1))
        U, L, hit = ex, f32(-ex), True
        if not (a0 <= dy_hi <= b0):
            dyc = min(max(dy_hi, a0), b0)
            disc = f32(q2 * A - det * dyc * dyc)
            hit = disc >= f32(-1.0e-3) * q2 * A
            U = f32((-B * dyc + np.sqrt(max(disc, f32(0)))) / A)
        if hit and not (a0 <= dy_lo <= b0):
            dyc = min(max(dy_lo, a0), b0)
            disc = f32(q2 * A - det * dyc * dyc)
            hit = disc >= f32(-1.0e-3) * q2 * A
            L = f32((-B * dyc - np.sqrt(max(disc, f32(0)))) / A)
        if not hit:
            continue
        U = f32(U + f32(1.0e-3) * abs(U) + f32(0.01))
        L = f32(L - (f32(1.0e-3) * abs(L) + f32(0.01)))
        c0 = max(tx0, int(np.floor((px + L) / TILE)))
        c1 = min(tx1 - 1, int(np.floor((px + U) / TILE)))
        for tx in range(c0, c1 + 1):
            keep.add((tx, cy))
    return keep


def brute_force_tiles(px, py, A, B, C, qc, grid):
    """tiles of the grid x grid tile image holding a pixel centre with q <= qc (float64)."""
    n = grid * TILE
    ys, xs = np.mgrid[0:n, 0:n].astype(np.float64)
    dx, dy = xs - px, ys - py
    q = 0.5 * (A * dx * dx + C * dy * dy) + B * dx * dy
    hit = q <= qc
    ty, tx = np.nonzero(hit.reshape(grid, TILE, grid, TILE).any(axis=(1, 3)))
    return set(zip(tx.tolist(), ty.tolist()))


def test_tile_mask_is_conservative_and_tight():
    rng = np.random.default_rng(7)
    grid = 10
    kept_total = ref_total = needed_total = 0
    for it in range(1500):
        # covariance of a screen-space Gaussian: sigmas 0.6 .. 30 px, any rotation, + the 0.3 dilation
        s1, s2 = np.exp(rng.uniform(np.log(0.6), np.log(30.0), 2))
        th = rng.uniform(0, np.pi)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        cov = R @ np.diag([s1 * s1, s2 * s2]) @ R.T + 0.3 * np.eye(2)
        Q = np.linalg.inv(cov)
        A, B, C = Q[0, 0], Q[0, 1], Q[1, 1]
        opacity = rng.uniform(0.005, 0.99)
        if opacity < 1.0 / 255.0:
            continue
        qc = float(np.log(255.0 * opacity))
        px, py = rng.uniform(-20, grid * TILE + 20, 2)
        # the reference's rectangle: ceil(3 sigma_max) square, clipped to the grid (auxiliary.h:46-59)
        lam = 0.5 * (cov[0, 0] + cov[1, 1]) + np.sqrt(max(0.1, (0.5 * (cov[0, 0] + cov[1, 1])) ** 2 - np.linalg.det(cov)))
        rad = int(np.ceil(3.0 * np.sqrt(lam)))
        clip = lambda v: min(grid, max(0, int(v)))
        rect = (clip((px - rad) / TILE), clip((py - rad) / TILE), clip((px + rad + TILE - 1) / TILE), clip((py + rad + TILE - 1) / TILE))
        ref_tiles = {(tx, ty) for ty in range(rect[1], rect[3]) for tx in range(rect[0], rect[2])}
        keep = tile_mask_model(px, py, A, B, C, qc, rect)
        need = brute_force_tiles(px, py, A, B, C, qc, grid) & ref_tiles     # the reference only blends inside its rectangle
        assert keep <= ref_tiles
        assert need <= keep, (it, sorted(need - keep), px, py, A, B, C, qc)
        kept_total += len(keep); ref_total += len(ref_tiles); needed_total += len(need)
    # and it is worth something: clearly fewer tiles than the reference's squares, close to what is needed
    assert kept_total < 0.75 * ref_total
    assert kept_total < 1.35 * needed_total + 50


def test_tile_mask_extremes():
    """Large Gaussians (rectangles above 32 tiles: bounding box only), needle-thin ones (axis ratio above ~2000: the
    reference's rectangle is kept), centres far outside the image, opacity at the 1/255 edge."""
    rng = np.random.default_rng(11)
    grid = 12
    for it in range(600):
        kind = it % 4
        if kind == 0:      # huge
            s1, s2 = np.exp(rng.uniform(np.log(20.0), np.log(200.0), 2))
        elif kind == 1:    # needle
            s1, s2 = rng.uniform(40.0, 400.0), rng.uniform(0.01, 0.2)
        elif kind == 2:    # tiny
            s1, s2 = rng.uniform(0.05, 0.6, 2)
        else:
            s1, s2 = np.exp(rng.uniform(np.log(0.6), np.log(30.0), 2))
        th = rng.uniform(0, np.pi)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        cov = R @ np.diag([s1 * s1, s2 * s2]) @ R.T + 0.3 * np.eye(2)
        Q = np.linalg.inv(cov)
        A, B, C = Q[0, 0], Q[0, 1], Q[1, 1]
        opacity = (1.0 / 255.0) * (1.0 + rng.uniform(0, 0.05)) if it % 7 == 0 else rng.uniform(0.005, 0.99)
        if opacity < 1.0 / 255.0:
            continue
        qc = float(np.log(255.0 * opacity))
        px, py = rng.uniform(-150, grid * TILE + 150, 2)
        lam = 0.5 * (cov[0, 0] + cov[1, 1]) + np.sqrt(max(0.1, (0.5 * (cov[0, 0] + cov[1, 1])) ** 2 - np.linalg.det(cov)))
        rad = int(np.ceil(3.0 * np.sqrt(lam)))
        clip = lambda v: min(grid, max(0, int(np.floor(v))))
        rect = (clip((px - rad) / TILE), clip((py - rad) / TILE), clip((px + rad + TILE - 1) / TILE), clip((py + rad + TILE - 1) / TILE))
        ref_tiles = {(tx, ty) for ty in range(rect[1], rect[3]) for tx in range(rect[0], rect[2])}
        keep = tile_mask_model(px, py, A, B, C, qc, rect)
        need = brute_force_tiles(px, py, A, B, C, qc, grid) & ref_tiles
        assert keep <= ref_tiles
        assert need <= keep, (it, kind, sorted(need - keep), px, py, A, B, C, qc)

"""k nearest neighbours among the Gaussian centres (csrc/knn.cu over the C-ABI).

Mirrors `knn(x, src, k)` of the reference's utils/general_utils.py:170-184 (which wraps pointops2's brute-force
knnquery) for the case train.py:132-152 uses -- a point cloud against itself: returns (idx int64 [b,n,k],
dist [b,n,k] = SQUARED distances ascending, the point itself first), found with a uniform-grid search instead of an
O(n^2) scan.  CUDA only.
"""
import torch

import fdgs


def knn(x: torch.Tensor, src: torch.Tensor, k: int, transpose: bool = False, brute_force: bool = False):
    if transpose:
        x, src = x.transpose(1, 2).contiguous(), src.transpose(1, 2).contiguous()
    if x.data_ptr() != src.data_ptr() and not torch.equal(x, src):
        raise NotImplementedError("fdgs.knn: neighbours of a cloud among itself only (the rigidity loss of train.py:132-152)")
    C = fdgs.ext()
    idx, d2 = [], []
    for b in range(x.shape[0]):
        i, d = C.knn(x[b].detach().contiguous(), int(k), bool(brute_force))
        idx.append(i.long())
        d2.append(d)
    return torch.stack(idx), torch.stack(d2)


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """Mean squared distance of every point to its 3 nearest OTHER points: the initial-scale heuristic of the
    reference's simple-knn extension (simple-knn/simple_knn.cu:139-178 boxMeanDist, called from
    scene/gaussian_model.py:273 create_from_pcd).  Exact 3-NN like the reference (which prunes Morton-sorted boxes);
    here on the same uniform grid as knn()."""
    x = points.detach().contiguous().float()
    n = x.shape[0]
    k = min(4, n)
    idx, d2 = fdgs.ext().knn(x, k, False)
    own = idx == torch.arange(n, device=x.device, dtype=idx.dtype)[:, None]
    # drop the point itself (first entry unless an exact duplicate with a lower index precedes it)
    first_own = own.float().argmax(dim=1)
    keep = torch.ones_like(own)
    keep[torch.arange(n, device=x.device), first_own] = False
    others = d2[keep].view(n, k - 1)
    if k - 1 < 3:
        others = torch.cat([others, others.new_full((n, 3 - (k - 1)), 3.4028234663852886e38)], 1)   # FLT_MAX like :148
    return (others[:, 0] + others[:, 1] + others[:, 2]) / 3.0

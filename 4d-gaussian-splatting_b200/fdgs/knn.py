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

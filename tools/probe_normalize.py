#!/usr/bin/env python3
"""Are the in-kernel activations of the raw-parameter entry bit-identical to torch's?  (run under gpurun)
exp / sigmoid: one candidate each; F.normalize: three summation orders of the 4-element norm (fdgs_common.cuh)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "4d-gaussian-splatting_b200"))
import fdgs  # noqa: E402

C = fdgs.ext()
dev = "cuda:0"
g = torch.Generator().manual_seed(1)
n = 1 << 20
log_s = (torch.randn(n, generator=g) * 1.5 - 4.0).to(dev)
logit = (torch.randn(n, generator=g) * 3.0).to(dev)
quat = (torch.randn(n, 4, generator=g) * (0.2 + 2 * torch.rand(n, 1, generator=g))).to(dev)
diff = lambda x, y: int((x.contiguous().view(torch.int32) != y.contiguous().view(torch.int32)).sum())
for mode in (0, 1, 2):
    s, o, q = C.debug_activate(log_s, logit, quat, mode)
    qn = torch.nn.functional.normalize(quat)
    print("mode %d: exp differing %d, sigmoid differing %d, normalize differing %d of %d (max |d| %.3e)" % (
        mode, diff(s, torch.exp(log_s)), diff(o, torch.sigmoid(logit)), diff(q, qn), 4 * n, float((q - qn).abs().max())))
# which norm does torch produce?
nrm = torch.linalg.vector_norm(quat, 2, dim=1)
x, y, z, w = [quat[:, i].double() for i in range(4)]
f = lambda t: t.float()
cands = {"pairwise": f(f(f(x * x) + f(y * y)).double() + f(f(z * z) + f(w * w)).double()),
         "sequential": f(f(f(f(x * x) + f(y * y)).double() + f(z * z).double()).double() + f(w * w).double()),
         "exact(double)": f(x * x + y * y + z * z + w * w)}
for k, v in cands.items():
    print("vector_norm vs sqrt(%s): differing %d" % (k, diff(nrm, torch.sqrt(v))))

# more candidates for the 4-element sum of squares, emulated in double (products of floats are exact in double)
X = [quat[:, i].double() for i in range(4)]
sq = [v * v for v in X]
fl = lambda t: t.float().double()
import itertools
cand = {}
for perm in itertools.permutations(range(4)):
    a, b, c, d = perm
    cand["fma-chain %s" % (perm,)] = fl(fl(fl(fl(sq[a]) + sq[b]) + sq[c]) + sq[d])
    cand["seq %s" % (perm,)] = fl(fl(fl(fl(sq[a]) + fl(sq[b])) + fl(sq[c])) + fl(sq[d]))
    cand["pair-fma %s" % (perm,)] = fl(fl(fl(sq[a]) + sq[b]) + fl(fl(sq[c]) + sq[d]))
best = sorted(((diff(nrm, torch.sqrt(v.float())), k) for k, v in cand.items()))[:6]
print("closest candidates to torch.linalg.vector_norm:", best)
n2 = (quat * quat).sum(1)
print("vector_norm vs sqrt((q*q).sum(1)): differing", diff(nrm, torch.sqrt(n2)))

# plain float32 arithmetic (torch ops round every product and every sum to fp32): all association orders
q2 = [quat[:, i] * quat[:, i] for i in range(4)]
res = []
for perm in itertools.permutations(range(4)):
    a, b, c, d = perm
    res.append((diff(nrm, torch.sqrt(((q2[a] + q2[b]) + q2[c]) + q2[d])), "f32 seq %s" % (perm,)))
    res.append((diff(nrm, torch.sqrt((q2[a] + q2[b]) + (q2[c] + q2[d]))), "f32 pair %s" % (perm,)))
print("plain fp32 orders closest to vector_norm:", sorted(res)[:4])

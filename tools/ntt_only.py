"""Runs a few batched forward/inverse NTT launches and one batch of rotations (for ncu captures of single kernels)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
pkg = common.package()
N = 16384
q = common.Q_16384
ctx = pkg.Context(N, common.T, q, device=0)
L, K = ctx.L, ctx.K
B = int(os.environ.get("PROBE_B", 148))
rng = np.random.default_rng(0)
a = np.empty((B, 2, L, N), dtype=np.uint64)
for i in range(L):
    a[:, :, i, :] = rng.integers(0, int(q[i]), (B, 2, N), dtype=np.uint64)
ksk = np.empty((L, 2, K, N), dtype=np.uint64)
for k in range(K):
    ksk[:, :, k, :] = rng.integers(0, int(q[k]), (L, 2, N), dtype=np.uint64)
ctx.load_ksk(0, ctx.galois_elt(-1), ksk)
d_a = ctx.dev_alloc(a.nbytes); d_o = ctx.dev_alloc(a.nbytes)
ctx.dev_upload(d_a, a)
for _ in range(3):
    ctx.dev_ntt(0, False, d_a, B * 2 * L)
    ctx.dev_ntt(0, True, d_a, B * 2 * L)
    ctx.dev_rotate_rows(d_a, -1, 0, d_o, B)
ctx.sync()
print("ok")

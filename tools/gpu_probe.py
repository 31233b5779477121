"""Device-side timing probe (CUDA events through torch on the context's stream). Not a bench: orientation numbers."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402

pkg = common.package()
N = int(os.environ.get("PROBE_N", 16384))
q = common.Q_16384 if N == 16384 else common.small_params(N, 6)
B = int(os.environ.get("PROBE_B", 148))
stream = torch.cuda.Stream()
ctx = pkg.Context(N, common.T, q, device=0, stream=stream.cuda_stream)
L, K = ctx.L, ctx.K
rng = np.random.default_rng(0)


def rnd_ct(count, size=2):
    out = np.empty((count, size, L, N), dtype=np.uint64)
    for i in range(L):
        out[:, :, i, :] = rng.integers(0, int(q[i]), (count, size, N), dtype=np.uint64)
    return out


def rnd_ksk():
    out = np.empty((L, 2, K, N), dtype=np.uint64)
    for k in range(K):
        out[:, :, k, :] = rng.integers(0, int(q[k]), (L, 2, N), dtype=np.uint64)
    return out


for step in (0, -1, 128):
    ctx.load_ksk(0, ctx.galois_elt(step), rnd_ksk())
ctx.load_ksk(2, 0, rnd_ksk())


def timed(fn, reps=3):
    fn()
    ctx.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(reps):
            fn()
        ev1.record(stream)
    ctx.sync()
    return ev0.elapsed_time(ev1) / reps


res = {"N": N, "B": B}
a = rnd_ct(B)
d_a = ctx.dev_alloc(a.nbytes)
d_o = ctx.dev_alloc(a.nbytes // 2 * 3)
ctx.dev_upload(d_a, a)
limbs = B * 2 * L
ms = timed(lambda: ctx.dev_ntt(0, False, d_a, limbs))
res["ntt_fwd_us_per_limb"] = ms * 1e3 / limbs
res["ntt_fwd_GBs"] = limbs * 16 * N / (ms * 1e-3) / 1e9
ms = timed(lambda: ctx.dev_ntt(0, True, d_a, limbs))
res["ntt_inv_GBs"] = limbs * 16 * N / (ms * 1e-3) / 1e9
ctx.dev_upload(d_a, a)
ms = timed(lambda: ctx.dev_rotate_rows(d_a, -1, 0, d_o, B))
res["rotate_ms_per_ct"] = ms / B
ms = timed(lambda: ctx.dev_multiply(d_a, d_a, d_o, B))
res["multiply_ms_per_ct"] = ms / B
a3 = rnd_ct(B, 3)
d_a3 = ctx.dev_alloc(a3.nbytes)
ctx.dev_upload(d_a3, a3)
ms = timed(lambda: ctx.dev_relinearize(d_a3, d_o, B))
res["relin_ms_per_ct"] = ms / B
print(json.dumps(res), flush=True)

for bsgs in (False, True):
    sym = rng.integers(0, common.T, (B, 128), dtype=np.uint64)
    d_sym = ctx.dev_alloc(sym.nbytes)
    ctx.dev_upload(d_sym, sym)
    d_key = ctx.dev_alloc(a[0].nbytes)
    ctx.dev_upload(d_key, a[0])
    if bsgs:
        for k in range(1, 8):
            ctx.load_ksk(0, ctx.galois_elt(-16 * k), rnd_ksk())
    lens = np.full(B, 128, dtype=np.uint32)
    ctr = np.arange(B, dtype=np.uint64)
    t0 = time.time()
    l0 = ctx.launch_count()
    ms = timed(lambda: ctx.dev_pasta3_decompose(d_key, d_sym, lens, ctr, common.NONCE, bsgs, d_o), reps=1)
    res = {"bsgs": bsgs, "decompose_ms_per_batch": ms, "blocks_per_s": B / (ms * 1e-3), "launches_per_batch": (ctx.launch_count() - l0) // 2,
           "wall_s": time.time() - t0}
    print(json.dumps(res), flush=True)

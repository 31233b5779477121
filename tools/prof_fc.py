"""Per-kernel time of the encrypted FC layer (multiply + relinearize + encrypted_vec_sum(n)) on a batch of samples: random operands and keys
(timing only; parity lives in tests/test_gpu_fc.py). PROBE_S=296 PROBE_LEN=128 python tools/prof_fc.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from tools import bench_configs as BC
pkg = common.package()
N, q = 16384, common.Q_16384
ctx = pkg.Context(N, common.T, q, device=0)
L, K = ctx.L, ctx.K
S, n = int(os.environ.get("PROBE_S", 296)), int(os.environ.get("PROBE_LEN", 128))
rng = np.random.default_rng(0)
ksk = BC._rnd_ksk(rng, q, L, K, N)
for e in range(13):
    for sgn in (1, -1):
        ctx.load_ksk(1, ctx.galois_elt(sgn * (1 << e)), ksk)
ctx.load_ksk(2, 0, ksk)
x, w = BC._rnd_ct(rng, q, L, N, S), BC._rnd_ct(rng, q, L, N, 1)
import ctypes as C
import torch
u64p = C.POINTER(C.c_uint64)


def call(xa, wa, oa):
    t0 = time.perf_counter()
    rc = ctx.lib.hhe_fc_rows(ctx.h, xa.ctypes.data_as(u64p), C.c_size_t(S), wa.ctypes.data_as(u64p), C.c_size_t(1), C.c_size_t(n), 1,
                             oa.ctypes.data_as(u64p))
    assert rc == 0
    return time.perf_counter() - t0


out_pg = np.zeros((S, 2, L, N), dtype=np.uint64)
call(x, w, out_pg)  # warm-up: staging buffers, scratch
ctx.profile(True); ctx.profile_reset()
dt = call(x, w, out_pg)
rep = ctx.profile_report()
ctx.profile(False)
print(f"fc_rows pageable in/out: {S} samples x 1 row of {n}: {dt*1e3:.1f} ms host clock ({dt/S*1e3:.3f} ms per sample)")
print(json.dumps({k: [round(v["ms"], 2), v.get("launches")] for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])}))
t0 = time.perf_counter(); fresh = np.zeros((S, 2, L, N), dtype=np.uint64); dtf = call(x, w, fresh)
print(f"fc_rows pageable in, fresh np.zeros out (first touch inside the call): {dtf*1e3:.1f} ms")
xp = torch.empty(x.shape, dtype=torch.int64).pin_memory(); xp.numpy()[...] = x.view(np.int64)
op = torch.empty(out_pg.shape, dtype=torch.int64).pin_memory()
call(xp.numpy().view(np.uint64), w, op.numpy().view(np.uint64))
dtp = call(xp.numpy().view(np.uint64), w, op.numpy().view(np.uint64))
print(f"fc_rows pinned in/out: {dtp*1e3:.1f} ms ({dtp/S*1e3:.3f} ms per sample)")
assert np.array_equal(op.numpy().view(np.uint64), out_pg)
ctx.set_batch(74)
call(xp.numpy().view(np.uint64), w, op.numpy().view(np.uint64))
dtc = call(xp.numpy().view(np.uint64), w, op.numpy().view(np.uint64))
print(f"fc_rows pinned in/out, 4 chunks of 74 (copy-out overlapped): {dtc*1e3:.1f} ms")
dtc2 = call(x, w, out_pg)
print(f"fc_rows pageable in/out, 4 chunks of 74: {dtc2*1e3:.1f} ms")

"""Per-kernel time of the config-5 primitives at one ring size (live CUDA-event profile of the engine): PROBE_N=32768 python tools/prof_prims.py"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from tools import bench_configs as BC
pkg = common.package()
N = int(os.environ.get("PROBE_N", 32768))
q = {8192: common.Q_8192, 16384: common.Q_16384, 32768: BC.Q_32768}[N]
ctx = pkg.Context(N, common.T, q, device=0)
L, K = ctx.L, ctx.K
B = int(os.environ.get("PROBE_B", 148))
rng = np.random.default_rng(0)
qmin = [min(int(v) for v in q)] * len(q)
a, a3 = BC._rnd_ct(rng, qmin, L, N, B), BC._rnd_ct(rng, qmin, L, N, B, 3)
ctx.load_ksk(0, ctx.galois_elt(-1), BC._rnd_ksk(rng, q, L, K, N))
ctx.load_ksk(2, 0, BC._rnd_ksk(rng, q, L, K, N))
d_a, d_a3, d_o = ctx.dev_alloc(a.nbytes), ctx.dev_alloc(a3.nbytes), ctx.dev_alloc(a3.nbytes)
ctx.dev_upload(d_a, a); ctx.dev_upload(d_a3, a3)
ops = {"ntt_fwd": lambda: ctx.dev_ntt(0, False, d_a, B * 2 * L), "ntt_inv": lambda: ctx.dev_ntt(0, True, d_a, B * 2 * L),
       "rotate": lambda: ctx.dev_rotate_rows(d_a, -1, 0, d_o, B), "relinearize": lambda: ctx.dev_relinearize(d_a3, d_o, B),
       "multiply": lambda: ctx.dev_multiply(d_a, d_a, d_o, B)}
for name, fn in ops.items():
    fn(); ctx.sync()
    ctx.profile(True); ctx.profile_reset()
    for _ in range(3):
        fn()
    ctx.sync()
    rep = ctx.profile_report()
    ctx.profile(False)
    print(name, json.dumps({k: round(v["ms"] / 3, 3) for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])}))

"""Secondary measurements for BASELINE.json configs[0..4] other than the headline bench (which is configs[3]-shaped):
  config1: one block latency (diagonal + BSGS), vs the reference on one core
  config2: MNIST 784->10 sample: transcipher 7 blocks + mask + flatten + 10 x (multiply, relin, vec_sum 784)
  config3: ECG 128->1, batch of samples (counter 0 each): transcipher + multiply + relin + vec_sum 128
  config5: primitive sweep N=8192/16384: NTT fwd/inv GB/s, rotate, relinearize per second (N=32768: not supported yet)
Prints one JSON object per config. CUDA-event timing on the engine's stream, device-resident unless noted."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from oracle import refshim as R
pkg = common.package()
T = common.T
stream = torch.cuda.Stream()

def timed(ctx, fn, reps=3):
    fn(); ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); ctx.sync()
    return e0.elapsed_time(e1) / reps

def rnd_ct(rng, q, L, N, count, size=2):
    out = np.empty((count, size, L, N), dtype=np.uint64)
    for i in range(L): out[:, :, i, :] = rng.integers(0, int(q[i]), (count, size, N), dtype=np.uint64)
    return out

def rnd_ksk(rng, q, L, K, N):
    out = np.empty((L, 2, K, N), dtype=np.uint64)
    for k in range(K): out[:, :, k, :] = rng.integers(0, int(q[k]), (L, 2, N), dtype=np.uint64)
    return out

def config5():
    rng = np.random.default_rng(0)
    Q_32768 = [36028797017456641, 36028797014704129, 36028797014573057, 36028797014376449, 36028797013327873, 36028797013000193,
               36028797012606977, 36028797010444289, 36028797009985537, 36028797005856769, 36028797005529089, 36028797005135873,
               36028797003694081, 36028797003563009, 36028797001138177, 72057594037338113]  # BFVDefault(32768), SURVEY.md B.1
    for N, q in ((8192, common.Q_8192), (16384, common.Q_16384), (32768, Q_32768)):
        ctx = pkg.Context(N, T, q, device=0, stream=stream.cuda_stream)
        L, K = ctx.L, ctx.K
        B = 296 if N < 32768 else 148
        ctx.load_ksk(0, ctx.galois_elt(-1), rnd_ksk(rng, q, L, K, N)); ctx.load_ksk(2, 0, rnd_ksk(rng, q, L, K, N))
        a = rnd_ct(rng, q, L, N, B); a3 = rnd_ct(rng, q, L, N, B, 3)
        d_a, d_a3, d_o = ctx.dev_alloc(a.nbytes), ctx.dev_alloc(a3.nbytes), ctx.dev_alloc(a3.nbytes)
        ctx.dev_upload(d_a, a); ctx.dev_upload(d_a3, a3)
        limbs = B * 2 * L
        f = timed(ctx, lambda: ctx.dev_ntt(0, False, d_a, limbs)); i = timed(ctx, lambda: ctx.dev_ntt(0, True, d_a, limbs))
        ctx.dev_upload(d_a, a)
        rot = timed(ctx, lambda: ctx.dev_rotate_rows(d_a, -1, 0, d_o, B)); rel = timed(ctx, lambda: ctx.dev_relinearize(d_a3, d_o, B))
        mul = timed(ctx, lambda: ctx.dev_multiply(d_a, d_a, d_o, B))
        print(json.dumps({"config": 5, "N": N, "L": L, "batch": B, "ntt_fwd_GBs": limbs * 16 * N / f / 1e6, "ntt_inv_GBs": limbs * 16 * N / i / 1e6,
                          "ntt_fwd_per_s": limbs / f * 1e3, "rotate_per_s": B / rot * 1e3, "relinearize_per_s": B / rel * 1e3, "multiply_per_s": B / mul * 1e3,
                          "fp64_moduli": ctx.info()["fp64_moduli"]}), flush=True)
        for p in (d_a, d_a3, d_o): ctx.dev_free(p)
        ctx.close()

def configs123():
    N = 16384
    steps = (0, -1, 128) + tuple(-128 * i for i in range(1, 7)) + tuple(-16 * k for k in range(1, 8))
    ref = R.Ref(N, T, None, seed=21, steps=steps, default_gk=True)
    ctx = pkg.Context(N, T, ref.q, device=0, stream=stream.cuda_stream)
    t0 = time.time(); common.load_keys_from_ref(ctx, ref, keysets=(0, 1)); t_keys = time.time() - t0
    rng = np.random.default_rng(8)
    key = rng.integers(0, T, 256, dtype=np.uint64)
    enc_key = ref.encrypt(common.pack_key(key, N))
    from oracle import oracle as O
    # ---- config 1: one block, host API end to end (H2D + D2H inside) ----
    sym = O.pasta_plain(key, T, rng.integers(0, T, 128, dtype=np.uint64))
    for bsgs in (False, True):
        ctx.pasta3_decompose(enc_key, sym, use_bsgs=bsgs)
        t0 = time.perf_counter(); out = ctx.pasta3_decompose(enc_key, sym, use_bsgs=bsgs); dt = time.perf_counter() - t0
        t0 = time.perf_counter(); want = ref.pasta_decompose(enc_key, sym, bsgs); dr = time.perf_counter() - t0
        print(json.dumps({"config": 1, "use_bsgs": bsgs, "b200_latency_s": dt, "reference_1core_s": dr, "speedup": dr / dt,
                          "bit_exact": bool(np.array_equal(out, want))}), flush=True)
    # ---- config 2: MNIST-shaped sample ----
    x = rng.integers(0, 256, 784, dtype=np.uint64); W = rng.integers(-8, 9, (10, 784))
    symx = O.pasta_plain(key, T, x)
    enc_w = np.stack([ref.encrypt(np.mod(W[r], T).astype(np.uint64)) for r in range(10)])
    import importlib
    host = importlib.import_module(common.PKG + ".host")
    hhe = host.PASTA_SEAL(ctx)
    def run2():
        flat = host.decompose(hhe, [symx], [enc_key], 784, mask_in_place=True)[0]
        return flat, host.evaluate_model(ctx, [flat], enc_w, 784)[0]
    run2()
    t0 = time.perf_counter(); flat, outs = run2(); dt = time.perf_counter() - t0
    logits = [int(ref.decrypt(outs[r])[0][783]) for r in range(10)]
    want = [int(v) % T for v in W @ x.astype(np.int64)]
    # reference cost model from its own per-op timings on this box (full run would take ~20 min of SEAL on one core)
    rot_s = ref.bench_primitive(enc_key, 2, 3); mul_s = ref.bench_primitive(enc_key, 5, 2); rel_s = ref.bench_primitive(enc_key, 3, 2)
    blk_s = ref.bench_decompose(enc_key, 1, 1, False)
    ref_est = 7 * blk_s + 6 * rot_s + 10 * (mul_s + rel_s + 2875 * rot_s)
    print(json.dumps({"config": 2, "b200_e2e_s": dt, "logits_match_plaintext": logits == want, "reference_1core_estimate_s": ref_est,
                      "reference_ops": {"block_s": blk_s, "rotate_s": rot_s, "multiply_s": mul_s, "relinearize_s": rel_s, "key_switches_per_row": 2875},
                      "speedup_vs_1core": ref_est / dt, "key_upload_s": t_keys}), flush=True)
    # ---- config 3: ECG batch (counter 0 for every sample) ----
    S = int(os.environ.get("ECG_BATCH", 1024))
    xs = rng.integers(0, 256, (S, 128), dtype=np.uint64); w = rng.integers(-128, 128, 128)
    syms = np.stack([O.pasta_plain(key, T, xs[i]) for i in range(S)])
    enc_w1 = ref.encrypt(np.mod(w, T).astype(np.uint64))[None]
    def run3():
        cts = ctx.pasta3_decompose(enc_key, syms.reshape(-1), records=S)
        return ctx.fc_rows(cts, enc_w1, 128)
    t0 = time.perf_counter(); outs = run3(); dt = time.perf_counter() - t0
    ok = all(int(ref.decrypt(outs[i, 0])[0][127]) == int(np.dot(xs[i].astype(np.int64), w)) % T for i in (0, S // 2, S - 1))
    ref_est = S * (blk_s + mul_s + rel_s + 355 * rot_s)
    print(json.dumps({"config": 3, "samples": S, "b200_e2e_s": dt, "samples_per_s": S / dt, "spot_check_dot_products": ok,
                      "reference_1core_estimate_s": ref_est, "speedup_vs_1core": ref_est / dt}), flush=True)

def config4_siesta():
    """SURVEY.md 8d config 4, real-data variant: SIESTA-shaped records (300 words in [0,31] -> 3 blocks, counters 0..2 restarting per
    record) through the service call hhe_csp_decompose (host buffers: transcipher + flatten per record). The engine regroups the blocks
    by counter, so the round material and the diagonals of each of the 3 counters are computed once per batch."""
    N = 16384
    ref = R.Ref(N, T, None, seed=31, steps=(0, -1, 128, -128, -256), default_gk=False)
    ctx = pkg.Context(N, T, ref.q, device=0, stream=stream.cuda_stream)
    common.load_keys_from_ref(ctx, ref, keysets=(0,))
    for s_ in (-128, -256):
        ctx.load_ksk(1, ref.galois_elt(s_), ref.ksk(0, ref.galois_elt(s_)))
    rng = np.random.default_rng(9)
    key = rng.integers(0, T, 256, dtype=np.uint64)
    enc_key = ref.encrypt(common.pack_key(key, N))
    from oracle import oracle as O
    Rn = int(os.environ.get("SIESTA_RECORDS", 98))  # 98 x 3 = 294 blocks: one lock-step wave
    recs = rng.integers(0, 32, (Rn, 300), dtype=np.uint64)
    syms = np.stack([O.pasta_plain(key, T, r) for r in recs])
    ctx.csp_decompose(enc_key, syms.reshape(-1), records=Rn, flatten_keys=1)
    t0 = time.perf_counter(); out = ctx.csp_decompose(enc_key, syms.reshape(-1), records=Rn, flatten_keys=1); dt = time.perf_counter() - t0
    ok = all(np.array_equal(ref.decrypt(out[r])[0][:300], recs[r]) for r in (0, Rn // 2, Rn - 1))
    blk_s = ref.bench_decompose(enc_key, 1, 1, False); rot_s = ref.bench_primitive(enc_key, 2, 3)
    print(json.dumps({"config": "4-siesta", "records": Rn, "blocks": 3 * Rn, "b200_e2e_s": dt, "records_per_s": Rn / dt, "blocks_per_s": 3 * Rn / dt,
                      "records_decrypt_to_input": ok, "reference_1core_estimate_s": Rn * (3 * blk_s + 2 * rot_s),
                      "speedup_vs_1core": Rn * (3 * blk_s + 2 * rot_s) / dt}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["5", "123", "4"]
    if "5" in which: config5()
    if "123" in which: configs123()
    if "4" in which: config4_siesta()

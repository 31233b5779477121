"""BASELINE.json configs other than the headline one, as functions bench.py calls (its `configs` object) and as a script.

  config 1  one PASTA-3 block, host API end to end (diagonal and BSGS), limb parity with the reference, reference time on this box
  bsgs      blocks/s of the BSGS affine layer on the headline batch (device resident)
  config 2  MNIST 784 -> 10 sample: 7 blocks + mask + flatten + 10 x (multiply, relinearize, encrypted_vec_sum(784))
  config 3  ECG 128 -> 1 batch (every record restarts at counter 0): transcipher + multiply + relinearize + encrypted_vec_sum(128)
  config 5  primitive sweep N = 8192 / 16384 / 32768: NTT fwd / inv, rotate_rows(-1), relinearize, multiply
  next rows SURVEY.md section 8(f): service handlers on SIESTA-shaped records, SEAL wire format, second FC layer, client-side encryption

Every entry carries a parity flag, the reference's time for the same work on this box's host cores (measured, or composed from its
own per-operation timings where the whole run would take tens of minutes -- the entry says which) and the fraction of the HBM roofline
computed from SURVEY.md section 8(d)'s algorithmic-byte formulas (operands read once + result written once per SEAL-level operation).
CUDA-event timing on the engine's stream for device-resident figures, host clock around the blocking call for end-to-end ones.
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import common  # noqa: E402

T = common.T
MIB = 1 << 20
BYTES_PER_BLOCK = {False: 7_751_991_296, True: 5_990_383_616}  # diagonal / BSGS
KS_COUNT = {128: 355, 300: 938, 784: 2875}                      # NAF key switches of encrypted_vec_sum(n), SURVEY.md App. D
Q_32768 = [36028797017456641, 36028797014704129, 36028797014573057, 36028797014376449, 36028797013327873, 36028797013000193,
           36028797012606977, 36028797010444289, 36028797009985537, 36028797005856769, 36028797005529089, 36028797005135873,
           36028797003694081, 36028797003563009, 36028797001138177, 72057594037338113]  # BFVDefault(32768), SURVEY.md B.1
# credited bytes per operation (evaluation keys excluded) for N = 8192 / 16384 / 32768
PRIM_BYTES = {8192: {"ntt": 131072, "rotate": 1048576, "relinearize": 1310720},
              16384: {"ntt": 262144, "rotate": 4194304, "relinearize": 5242880},
              32768: {"ntt": 524288, "rotate": 15728640, "relinearize": 19660800}}
BSGS_STEPS = tuple(-16 * k for k in range(1, 8))
FLATTEN_STEPS = tuple(-128 * i for i in range(1, 7))
ALL_STEPS = (0, -1, 128) + FLATTEN_STEPS + BSGS_STEPS


def fc_row_bytes(n, ct_mib=2):
    """multiply (2 ct + ct3) + relinearize (ct3 + ct) + (n-1) adds (3 ct) + KS(n) key switches (2 ct), SURVEY.md 8(d)"""
    return int((3.5 * ct_mib + 2.5 * ct_mib + (n - 1) * 3 * ct_mib + KS_COUNT[n] * 2 * ct_mib) * MIB)


def _timed(ctx, stream, fn, reps=3):
    import torch
    fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    ctx.sync()
    return e0.elapsed_time(e1) / reps * 1e-3


def _rnd_ct(rng, q, L, N, count, size=2):
    out = np.empty((count, size, L, N), dtype=np.uint64)
    for i in range(L):
        out[:, :, i, :] = rng.integers(0, int(q[i]), (count, size, N), dtype=np.uint64)
    return out


def _rnd_ksk(rng, q, L, K, N):
    out = np.empty((L, 2, K, N), dtype=np.uint64)
    for k in range(K):
        out[:, :, k, :] = rng.integers(0, int(q[k]), (L, 2, N), dtype=np.uint64)
    return out


def load_all_keys(ctx, ref):
    """PASTA + flatten + BSGS keys into keyset 0, the analyst's default power-of-two set into keyset 1, relin key."""
    common.load_keys_from_ref(ctx, ref, keysets=(0, 1))


def config1(ctx, ref, enc_key, key, rng, peak):
    from oracle import oracle as O
    sym = O.pasta_plain(key, T, rng.integers(0, T, 128, dtype=np.uint64))
    out = {}
    for bsgs in (False, True):
        ctx.pasta3_decompose(enc_key, sym, use_bsgs=bsgs)
        lat = []
        for _ in range(3):
            t0 = time.perf_counter()
            got = ctx.pasta3_decompose(enc_key, sym, use_bsgs=bsgs)
            lat.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        want = ref.pasta_decompose(enc_key, sym, bsgs)
        dr = time.perf_counter() - t0
        dt = min(lat)
        out["bsgs" if bsgs else "diagonal"] = {
            "b200_latency_s": dt, "reference_1core_s": dr, "speedup_vs_1core": dr / dt, "parity_limb_exact": bool(np.array_equal(got, want)),
            "algorithmic_bytes": BYTES_PER_BLOCK[bsgs], "frac_of_hbm_roofline": BYTES_PER_BLOCK[bsgs] / dt / 1e9 / peak,
            "note": "one block cannot fill 148 SMs (18 CTAs per key switch): latency is bound by the chain of dependent launches"}
    return out


def config2(ctx, ref, enc_key, key, rng, peak, ref_ops):
    from oracle import oracle as O
    host = importlib.import_module(common.PKG + ".host")
    x = rng.integers(0, 256, 784, dtype=np.uint64)
    W = rng.integers(-8, 9, (10, 784))
    symx = O.pasta_plain(key, T, x)
    enc_w = np.stack([ref.encrypt(np.mod(W[r], T).astype(np.uint64)) for r in range(10)])
    hhe = host.PASTA_SEAL(ctx)

    def run():
        flat = host.decompose(hhe, [symx], [enc_key], 784, mask_in_place=True)[0]
        return flat, host.evaluate_model(ctx, [flat], enc_w, 784)[0]
    run()
    t0 = time.perf_counter()
    flat, outs = run()
    dt = time.perf_counter() - t0
    logits = [int(ref.decrypt(outs[r])[0][783]) for r in range(10)]
    want = [int(v) % T for v in W @ x.astype(np.int64)]
    nbytes = 7 * BYTES_PER_BLOCK[False] + int(4.125 * MIB) + 6 * 10 * MIB + 10 * fc_row_bytes(784)
    ref_est = 7 * ref_ops["block_s"] + 6 * ref_ops["rotate_s"] + ref_ops["multiply_plain_s"] + 10 * (
        ref_ops["multiply_s"] + ref_ops["relinearize_s"] + KS_COUNT[784] * ref_ops["rotate_s"])
    return {"b200_e2e_s": dt, "parity_logits_equal_plaintext_dot_products": logits == want,
            "parity_note": "one output neuron limb-exact vs the reference: tests/test_gpu_fc.py::test_mnist_output_neuron_limb_exact",
            "reference_1core_s": ref_est, "reference_kind": "composed from the reference's per-operation times measured in this run "
            "(7 blocks + 6 rotations + 10 x (multiply + relinearize + 2,875 key switches)); the whole run is ~2.5 h of SEAL on one core",
            "speedup_vs_1core": ref_est / dt, "algorithmic_bytes": nbytes, "frac_of_hbm_roofline": nbytes / dt / 1e9 / peak}


def config3(ctx, ref, enc_key, key, rng, peak, ref_ops, samples=1024, rank=0, world=1):
    """ECG batch; with world > 1 every rank takes a contiguous share of the samples (no data-path collective)."""
    from oracle import oracle as O
    S = samples // world
    xs = rng.integers(0, 256, (samples, 128), dtype=np.uint64)[rank * S:(rank + 1) * S]
    w = rng.integers(-128, 128, 128)
    syms = np.stack([O.pasta_plain(key, T, xs[i]) for i in range(S)])
    enc_w1 = ref.encrypt(np.mod(w, T).astype(np.uint64))[None] if ref is not None else None
    return xs, w, syms, enc_w1


def config3_run(ctx, enc_key, syms, enc_w1, bufs=None):
    """decompose (records -> ciphertexts in host memory) then the FC layer on them (host in, host out), like the CSP's two handlers.
    bufs: two caller-owned host arrays for the two results (bench.py passes pinned memory, reused between calls)."""
    S = syms.shape[0]
    t0 = time.perf_counter()
    cts = ctx.pasta3_decompose(enc_key, syms.reshape(-1), records=S, out=None if bufs is None else bufs[0][:S])
    outs = ctx.fc_rows(cts, enc_w1, 128, out=None if bufs is None else bufs[1][:S])
    return outs, time.perf_counter() - t0


def next_rows(ctx, ref, enc_key, key, rng, ref_ops, records=98, eval_samples=32, enc_count=64, plain_blocks=4096):
    """SURVEY.md section 8(f) rows, each measured through the call a user makes (host buffers, host clock) with a parity flag checked
    on the spot and the reference's time for the same work on one host core."""
    from oracle import oracle as O
    pkg = common.package()
    host = importlib.import_module(common.PKG + ".host")
    seal_io = importlib.import_module(common.PKG + ".seal_io")
    N, L = ctx.N, ctx.L
    out = {}

    def best(fn, reps=2):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            r = fn()
            ts.append(time.perf_counter() - t0)
        return r, min(ts)

    # f.1 service handlers: BaseCSP::decompose on SIESTA-shaped records (300 words in [0,31] -> 3 blocks, counters 0..2, masked
    # and flattened into one ciphertext per record), then evaluateModel (one weight row of length 300) on 32 of them
    R_, n = records, 300
    xs = rng.integers(0, 32, (R_, n), dtype=np.uint64)
    syms = np.stack([O.pasta_plain(key, T, xs[i]) for i in range(R_)])
    recs, dt = best(lambda: ctx.csp_decompose(enc_key, syms.reshape(-1), records=R_, apply_mask=True, flatten_keys=pkg.KEYSET_0))
    ok = all(np.array_equal(ref.decrypt(recs[i])[0][:n], xs[i]) for i in (0, R_ - 1))
    w = rng.integers(-8, 9, n)
    enc_w = ref.encrypt(np.mod(w, T).astype(np.uint64))[None]
    S_ = eval_samples
    outs, dte = best(lambda: ctx.csp_evaluate_model(recs[:S_], enc_w, n))
    dece = [ref.decrypt(outs[i][0]) for i in (0, S_ - 1)]
    oke = all(int(d[0][n - 1]) == int(w @ xs[i].astype(np.int64)) % T for d, i in zip(dece, (0, S_ - 1)))
    ref_rec = 3 * ref_ops["block_s"] + 2 * ref_ops["rotate_s"] + ref_ops["multiply_plain_s"]
    ref_eval = ref_ops["multiply_s"] + ref_ops["relinearize_s"] + KS_COUNT[n] * ref_ops["rotate_s"]
    out["f1_service_handlers"] = {
        "csp_decompose": {"records": R_, "words_per_record": n, "records_per_s": R_ / dt, "blocks_per_s": 3 * R_ / dt,
                          "parity_records_decrypt_to_input": bool(ok), "reference_1core_s_per_record": ref_rec,
                          "note": "every record uses counters 0..2: their three keystream ciphertexts are evaluated once per call and "
                                  "shared by all records (bit-identical to per-record evaluation, HHE_NO_SHARED_KEYSTREAM=1 restores it: "
                                  "52.7 records/s); what remains per record is encode + add_plain, mask, flatten and the copy-out",
                          "speedup_vs_1core": R_ / dt * ref_rec},
        "csp_evaluate_model": {"samples": S_, "row_length": n, "samples_per_s": S_ / dte, "parity_slot_equals_dot_product": bool(oke),
                               "noise_budget_left_bits": min(int(d[1]) for d in dece),
                               "reference_1core_s_per_sample": ref_eval, "speedup_vs_1core": S_ / dte * ref_eval},
        "parity_note": "limb-exact vs the reference: tests/test_gpu_fc.py::test_siesta_record_flattened_limb_exact, test_gpu_dropin",
        "reference_kind": "composed from the reference's per-operation times measured in this run"}

    # f.2 SEAL wire format at the boundary: the codec alone (host code) and a transciphering call that takes / returns serialized bytes
    ring = seal_io.Ring(N, T, ctx.q, lib=ctx.lib)
    ct = recs[0]
    res = {}
    for name, mode in (("none", 0), ("zstd", 2)):
        blob, ts = best(lambda m=mode: ring.save_ciphertext(ct, m), 3)
        (back, _), tl = best(lambda b=blob: ring.load_ciphertext(b), 3)
        t0 = time.perf_counter()
        want = ref.ct_save(ct, mode)
        tr = time.perf_counter() - t0
        res[name] = {"bytes": len(blob), "save_MBps": ct.nbytes / ts / 1e6, "load_MBps": ct.nbytes / tl / 1e6,
                     "seal_save_MBps_1core": ct.nbytes / tr / 1e6,
                     "parity_bytes_equal_seal": bool(blob == want) if mode == 0 else None,
                     "parity_seal_loads_ours_and_round_trip": bool(np.array_equal(ref.ct_load(blob)[0], ct) and np.array_equal(back, ct))}
    key_blob = ref.ct_save(enc_key, 2)
    nb8 = min(8, syms.size // 128)
    sym8 = syms.reshape(-1)[:nb8 * 128]
    plain8, tp = best(lambda: ctx.pasta3_decompose(enc_key, sym8), 1)
    ser8, tsr = best(lambda: ctx.pasta3_decompose_serialized(key_blob, sym8), 1)
    res["decompose_serialized"] = {"blocks": nb8, "seconds": tsr, "seconds_raw_buffers": tp,
                                   "parity_equals_raw_call": bool(all(np.array_equal(ref.ct_load(ser8[b])[0], plain8[b]) for b in range(nb8)))}
    res["parity_note"] = "tests/test_seal_codec.py (pinned against libseal-4.0.a both ways, committed SEAL-written fixture)"
    out["f2_seal_wire_format"] = res

    # f.3 second FC layer: 128 -> 8 -> 2 network on one decomposed record (fc1 rows, square activation, fc2 with plain weights)
    H, n1 = 8, 128
    x1 = rng.integers(0, 16, n1, dtype=np.uint64)
    W1 = rng.integers(-2, 3, (H, n1))
    W2 = rng.integers(-3, 4, (2, H))
    W2[W2 == 0] = 1
    rec1 = ctx.pasta3_decompose(enc_key, O.pasta_plain(key, T, x1))[0]
    enc_w1 = np.stack([ref.encrypt(np.mod(W1[r], T).astype(np.uint64)) for r in range(H)])
    o2, t2 = best(lambda: host.evaluate_model_2fc(ctx, [rec1], enc_w1, n1, W2), 1)
    hidden = W1 @ x1.astype(np.int64)
    want2 = [int(v) % T for v in W2 @ (hidden * hidden)]
    dec = [ref.decrypt(o2[0][k]) for k in range(2)]
    ref2 = H * (ref_ops["multiply_s"] + ref_ops["relinearize_s"] + KS_COUNT[n1] * ref_ops["rotate_s"]) + H * (
        ref_ops["multiply_s"] + ref_ops["relinearize_s"]) + 2 * H * ref_ops["multiply_plain_s"]
    out["f3_second_fc_layer"] = {
        "network": "128 -> 8 -> x^2 -> 2", "seconds_per_sample": t2, "parity_decrypts_to_plaintext_network": [int(d[0][n1 - 1]) for d in dec] == want2,
        "noise_budget_left_bits": min(int(d[1]) for d in dec), "reference_1core_s": ref2, "speedup_vs_1core": ref2 / t2,
        "reference_kind": "composed from the reference's per-operation times measured in this run",
        "parity_note": "limb-exact vs SEAL's operation sequence: tests/test_gpu_fc.py::test_two_fc_layers_limb_exact_and_decrypt"}

    # f.4 client side: Encryptor::encrypt of batch-encoded rows and plain PASTA-3 encryption of a word stream
    C_ = enc_count
    rows = rng.integers(0, T, (C_, 128), dtype=np.uint64)
    pk = ref.public_key()
    seeds = rng.integers(0, 1 << 63, (C_, 8), dtype=np.uint64)
    cts, te = best(lambda: ctx.encrypt(pk, slots=rows, seeds=seeds))
    oke = all(np.array_equal(ref.decrypt(cts[i])[0][:128], rows[i]) for i in (0, C_ - 1))
    t0 = time.perf_counter()
    for i in range(min(4, C_)):
        ref.encrypt(rows[i])
    tre = (time.perf_counter() - t0) / min(4, C_)
    words = rng.integers(0, T, 128 * plain_blocks, dtype=np.uint64)
    enc, tpl = best(lambda: ctx.pasta3_plain(key, words))
    t0 = time.perf_counter()
    want_p = O.pasta_plain(key, T, words[:128 * min(64, plain_blocks)])
    tro = (time.perf_counter() - t0) / min(64, plain_blocks)
    out["f4_client_side"] = {
        "encrypt": {"ciphertexts": C_, "per_s": C_ / te, "parity_decrypts_to_input": bool(oke), "seal_1core_per_s": 1 / tre,
                    "speedup_vs_1core": C_ / te * tre,
                    "note": "host buffers in and out: the 2 MiB device-to-host copy per ciphertext into pageable memory is most of the call",
                    "parity_note": "bit-identical to seal::Encryptor::encrypt for the same seed: tests/test_encrypt.py"},
        "pasta3_plain": {"blocks": plain_blocks, "words_per_s": words.size / tpl, "parity_equals_cpu_restatement": bool(np.array_equal(enc[:want_p.size], want_p)),
                         "cpu_port_1core_words_per_s": 128 / tro, "speedup_vs_1core": words.size / tpl * tro / 128,
                         "parity_note": "reference KAT: tests/test_gpu_*::test_plain_pasta3_matches_reference_kat_and_oracle"}}
    return out


def config5(stream, peak, ref_factory=None, sizes=(8192, 16384, 32768)):
    import torch  # noqa: F401
    pkg = common.package()
    rng = np.random.default_rng(0)
    res = {}
    for N, q in ((8192, common.Q_8192), (16384, common.Q_16384), (32768, Q_32768)):
        if N not in sizes:
            continue
        ctx = pkg.Context(N, T, q, device=torch.cuda.current_device(), stream=stream.cuda_stream)
        L, K = ctx.L, ctx.K
        B = 296 if N < 32768 else 148
        ctx.load_ksk(0, ctx.galois_elt(-1), _rnd_ksk(rng, q, L, K, N))
        ctx.load_ksk(2, 0, _rnd_ksk(rng, q, L, K, N))
        qmin = [min(int(v) for v in q)] * len(q)  # residues valid for every limb: the NTT launches below use table 0 for all limbs
        a, a3 = _rnd_ct(rng, qmin, L, N, B), _rnd_ct(rng, qmin, L, N, B, 3)
        d_a, d_a3, d_o = ctx.dev_alloc(a.nbytes), ctx.dev_alloc(a3.nbytes), ctx.dev_alloc(a3.nbytes)
        ctx.dev_upload(d_a, a)
        ctx.dev_upload(d_a3, a3)
        limbs = B * 2 * L
        f = _timed(ctx, stream, lambda: ctx.dev_ntt(0, False, d_a, limbs))
        i = _timed(ctx, stream, lambda: ctx.dev_ntt(0, True, d_a, limbs))
        # round trip parity of the timed kernels on the spot: 8 forward + 8 inverse transforms ran, the data must be the input again
        back = np.empty_like(a)
        ctx.dev_download(d_a, back)
        ctx.dev_upload(d_a, a)
        rot = _timed(ctx, stream, lambda: ctx.dev_rotate_rows(d_a, -1, 0, d_o, B))
        rel = _timed(ctx, stream, lambda: ctx.dev_relinearize(d_a3, d_o, B))
        mul = _timed(ctx, stream, lambda: ctx.dev_multiply(d_a, d_a, d_o, B))
        pb = PRIM_BYTES[N]
        entry = {
            "L": L, "batch": B, "fp64_path_moduli": ctx.info()["fp64_moduli"], "parity_ntt_round_trip": bool(np.array_equal(back, a)),
            "parity_note": "limb parity of every primitive with SEAL at each size: tests/test_gpu_fullsize.py (N = 8192, 16384, 32768 cases)",
            "ntt_fwd": {"GBps": limbs * pb["ntt"] / f / 1e9, "limbs_per_s": limbs / f, "frac_of_hbm_roofline": limbs * pb["ntt"] / f / 1e9 / peak},
            "ntt_inv": {"GBps": limbs * pb["ntt"] / i / 1e9, "limbs_per_s": limbs / i, "frac_of_hbm_roofline": limbs * pb["ntt"] / i / 1e9 / peak},
            "rotate_rows": {"per_s": B / rot, "GBps": B * pb["rotate"] / rot / 1e9, "frac_of_hbm_roofline": B * pb["rotate"] / rot / 1e9 / peak},
            "relinearize": {"per_s": B / rel, "GBps": B * pb["relinearize"] / rel / 1e9,
                            "frac_of_hbm_roofline": B * pb["relinearize"] / rel / 1e9 / peak},
            "multiply": {"per_s": B / mul},
        }
        for p_ in (d_a, d_a3, d_o):
            ctx.dev_free(p_)
        ctx.close()
        if ref_factory is not None:
            r = ref_factory(N)
            ct = r.encrypt(np.arange(16, dtype=np.uint64))
            ntt_f, ntt_i = r.bench_primitive(ct, 0, 20), r.bench_primitive(ct, 1, 20)
            reps = 3 if N < 32768 else 1
            rr, rl, rm = r.bench_primitive(ct, 2, reps), r.bench_primitive(ct, 3, reps), r.bench_primitive(ct, 5, reps)
            entry["reference_1core"] = {"ntt_fwd_limbs_per_s": 1 / ntt_f, "ntt_inv_limbs_per_s": 1 / ntt_i, "rotate_rows_per_s": 1 / rr,
                                        "relinearize_per_s": 1 / rl, "multiply_per_s": 1 / rm}
            entry["speedup_vs_1core"] = {"ntt_fwd": limbs / f * ntt_f, "rotate_rows": B / rot * rr, "relinearize": B / rel * rl,
                                         "multiply": B / mul * rm}
            r.close()
        res[f"N{N}"] = entry
    return res


if __name__ == "__main__":
    import torch
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    from oracle import refshim as R
    stream = torch.cuda.Stream()
    fac = (lambda N: R.Ref(N, T, None, seed=5, steps=(0, -1), default_gk=False)) if R.available() else None
    print(json.dumps({"config5": config5(stream, peak, fac)}))

#!/usr/bin/env python
"""bench.py -- PASTA-3 blocks transciphered per second at BFV N=16384, t=65537 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--blocks B] [--bsgs] [--stream TOTAL] [--no-configs] [--impl reference]

A "step" transciphers B PASTA-3 blocks with distinct SHAKE counters per GPU (BASELINE.json configs[3], the
SpO2-stream case, sharded: rank r, step s owns counters [(s*G + r)*B, (s*G + r + 1)*B)); work per GPU is fixed as
N grows (weak scaling). Independent blocks need no data-path collective; rank 0 only gathers one 64-bit digest per
block over NCCL at the end of every step.

  value  : blocks/s with the symmetric ciphertext and the encrypted key already resident in HBM (CUDA events on the
           engine's stream, max over ranks).
  e2e    : the same metric through the host-buffer C ABI call hhe_pasta3_decompose (pinned host buffers; H2D of the
           inputs and D2H of every output ciphertext inside the timed region).
  roofline: the dominant kernel (ks_digits, the key-switch digit NTT + inner product) timed live with CUDA events.
  cpu_baseline / --impl reference: the UNMODIFIED reference (src/pasta + vendored libseal via oracle/_ref) on all host
           cores, one PASTA_SEAL per thread, one block per thread per step.
  strong : a FIXED stream of blocks (--stream TOTAL, e.g. 65536 = BASELINE configs[3] in full; default 8 x 296) split over the
           ranks in contiguous counter ranges: blocks/s with total work fixed, next to the weak-scaling `value`.
  fc     : BASELINE configs[2] (ECG 128 -> 1, batch 1024) with the samples sharded over the ranks: transcipher + encrypted FC.
  configs: (one GPU) configs 1, 2, 5 and the BSGS mode, each with a parity flag, the reference's time on this box and the
           fraction of the HBM roofline from SURVEY.md 8(d)'s byte formulas (tools/bench_configs.py).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402

N = 16384
METRIC = "pasta3_blocks_transciphered_per_s"
UNIT = "blocks/s"
MIB = 1 << 20
# algorithmic bytes per PASTA-3 block (operands read once + result written once per SEAL-level op)
BYTES_PER_BLOCK = {False: 7_751_991_296, True: 5_990_383_616}
KS_BYTES = 4 * MIB  # one key switch (rotate_rows / rotate_columns): 2 ciphertexts
# FP64-pipe instructions (per thread) ks_digits needs for one item at N=16384: per (key limb, half) CTA
# 7 digit transforms of 8192 residues (fold 23 x 4096, four radix-8 passes 84 x 1024 each + 3 half reductions 8 x 1024,
# multiply-accumulate with the key 2 x 6 x 8192) + the digit reused from the plaintext product (13 x 8192) + write-out.
KS_FP64_PER_ITEM = 18 * (7 * (23 * 4096 + 4 * 84 * 1024 + 3 * 8 * 1024 + 12 * 8192) + 13 * 8192 + 2 * 4 * 8192)
FP64_LANES_PER_SM_CLK = 64  # DFMA/DMUL/DADD: one warp instruction per 2 cycles per sub-partition (tools/microbench/pipes.cu)


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def config(args, world):
    return {
        "workload": f"BASELINE configs[3] shard: {args.blocks} distinct-counter PASTA-3 blocks per GPU per step, "
                    f"{'BSGS' if args.bsgs else 'diagonal (reference default use_bsgs=false)'} affine layers",
        "N": N, "t": common.T, "coeff_modulus": "BFVDefault(16384): L=8 data limbs + special prime",
        "blocks_per_gpu_per_step": args.blocks, "use_bsgs": bool(args.bsgs), "parallelism": f"shard{world}",
        "arith": "exact mod-q arithmetic on u64 residues: error-free FP64-pipe products for the 48/49-bit coefficient primes, integer Shoup products for the 61-bit BEHZ primes",
        "l2": "working set per step (>=0.6 GB of ciphertext state + 144 MB of key-switching keys) exceeds the 126 MB L2",
    }


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows), "reasons": sorted(reasons)}


def run_reference(args, rank, emit):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    from oracle import refshim as R
    if not R.available():
        emit({"impl": "reference", "unavailable": "oracle/_ref/libhhe_ref.so was not built (needs /root/reference at build time)"})
        return
    cores = os.cpu_count() or 1
    steps = [0, -1, 128] + ([-16 * k for k in range(1, 8)] if args.bsgs else [])
    ref = R.Ref(N, common.T, None, seed=4, steps=tuple(steps), default_gk=False)
    rng = np.random.default_rng(4)
    enc_key = ref.encrypt(common.pack_key(rng.integers(0, common.T, 256, dtype=np.uint64), N))
    per = max(1, args.ref_blocks_per_thread)
    for _ in range(args.warmup):
        ref.bench_decompose(enc_key, cores, per, args.bsgs)
    total = 0.0
    for _ in range(args.steps):
        total += ref.bench_decompose(enc_key, cores, per, args.bsgs)
    value = cores * per * args.steps / total
    sample = f"{cores} threads x {per} block(s) per step, one pasta::PASTA_SEAL per thread (src/pasta/pasta_3_seal.cpp:106-172 + libseal-4.0.a)"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--blocks", type=int, default=296, help="PASTA blocks per GPU per step")
    ap.add_argument("--bsgs", action="store_true", help="baby-step/giant-step affine layers (reference: use_bsgs=true)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-blocks-per-thread", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary configs (1, 2, 3, 5, BSGS, strong-scaling sample)")
    ap.add_argument("--stream", type=int, default=0, help="strong-scaling mode: total blocks of the fixed stream (default 8 x --blocks)")
    ap.add_argument("--ecg-samples", type=int, default=1024)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    # stdout carries exactly ONE JSON line (rank 0): everything libraries print at the C level (NCCL's version banner ignores
    # NCCL_DEBUG_FILE at NCCL_DEBUG=VERSION) is sent to stderr for the whole run; the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        return run_reference(args, rank, emit)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import refshim as R
    ref_ok = R.available()
    pkg = common.package()
    stream = torch.cuda.Stream()
    ctx = pkg.Context(N, common.T, common.Q_16384, device=local, stream=stream.cuda_stream)
    ctx.set_batch(args.blocks)
    L = ctx.L
    B = args.blocks

    # ---- inputs: rank 0 generates the key material, NCCL broadcasts it (one-time, untimed) ----
    # keyset 0: PASTA keys (+ the BSGS and flatten rotations when the secondary configs run); keyset 1: the analyst's default
    # power-of-two Galois keys for encrypted_vec_sum (FC line); relinearisation key.
    extras = not args.no_configs
    ref = None
    steps_ = [0, -1, 128] + ([-16 * k for k in range(1, 8)] if (args.bsgs or extras) else [])
    if extras and world == 1:
        steps_ += [-128 * i for i in range(1, 7)]
    meta_t = torch.zeros(2, dtype=torch.int64, device="cuda")
    ks = []
    if rank == 0:
        rng0 = np.random.default_rng(4)
        sym_key = rng0.integers(0, common.T, 256, dtype=np.uint64)
        if ref_ok:
            ref = R.Ref(N, common.T, None, seed=4, steps=tuple(steps_), default_gk=extras)
            ks = [(0, ref.galois_elt(s), ref.ksk(0, ref.galois_elt(s))) for s in steps_] + [(2, 0, ref.ksk(2))]
            if extras:
                ks += [(1, e, ref.ksk(1, e)) for e in ref.list_galois(1)]
            enc_key = ref.encrypt(common.pack_key(sym_key, N))
        else:
            q = common.Q_16384
            def rnd_key():
                k = np.empty((L, 2, L + 1, N), dtype=np.uint64)
                for i, m in enumerate(q):
                    k[:, :, i, :] = rng0.integers(0, m, (L, 2, N), dtype=np.uint64)
                return k
            ks = [(0, ctx.galois_elt(s), rnd_key()) for s in steps_] + [(2, 0, rnd_key())]
            enc_key = np.stack([np.stack([rng0.integers(0, m, N, dtype=np.uint64) for m in q[:L]]) for _ in range(2)])
        meta_t[0], meta_t[1] = len(ks), int(ref is not None)
    if world > 1:
        dist.broadcast(meta_t, 0)
    nkeys, have_ref = int(meta_t[0].item()), bool(meta_t[1].item())
    extras = extras and have_ref  # the secondary configs need real keys (their parity flags decrypt with SEAL)
    ek_t = torch.empty((2, L, N), dtype=torch.int64, device="cuda")
    sk_t = torch.empty(256, dtype=torch.int64, device="cuda")
    elt_t = torch.zeros((nkeys, 2), dtype=torch.int64, device="cuda")
    if rank == 0:
        for i, (kind, elt, _) in enumerate(ks):
            elt_t[i, 0], elt_t[i, 1] = kind, elt
        ek_t.copy_(torch.from_numpy(enc_key.view(np.int64)))
        sk_t.copy_(torch.from_numpy(sym_key.view(np.int64)))
    if world > 1:
        for t_ in (elt_t, ek_t, sk_t):
            dist.broadcast(t_, 0)
    elts = elt_t.cpu().numpy()
    key_t = torch.empty((L, 2, L + 1, N), dtype=torch.int64, device="cuda")
    for i in range(nkeys):  # one key (18 MiB) at a time: broadcast over NCCL, upload into the engine
        if rank == 0:
            key_t.copy_(torch.from_numpy(ks[i][2].view(np.int64)))
        if world > 1:
            dist.broadcast(key_t, 0)
        ctx.load_ksk(int(elts[i, 0]), int(elts[i, 1]), key_t.cpu().numpy().view(np.uint64))
    del key_t, ks
    enc_key = ek_t.cpu().numpy().view(np.uint64)
    sym_key = sk_t.cpu().numpy().view(np.uint64)

    # synthetic symmetric ciphertext stream for this rank (uniform words < t, seed 4 + rank)
    rng = np.random.default_rng(1000 + rank)
    sym_host = torch.empty((B, 128), dtype=torch.int64).pin_memory()
    sym_np = sym_host.numpy().view(np.uint64)
    sym_np[:] = rng.integers(0, common.T, (B, 128), dtype=np.uint64)
    out_host = torch.empty((B, 2, L, N), dtype=torch.int64).pin_memory()
    d_sym = sym_host.cuda()
    d_key = ek_t
    d_out = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda")
    lens = np.full(B, 128, dtype=np.uint32)
    digests = torch.empty((world, B), dtype=torch.int64, device="cuda") if rank == 0 else None
    import ctypes as C
    ptr = lambda t_: C.c_void_p(t_.data_ptr())  # noqa: E731

    def counters(step):
        base = (step * world + rank) * B
        return np.arange(base, base + B, dtype=np.uint64)

    def device_step(step):
        ctx.dev_pasta3_decompose(ptr(d_key), ptr(d_sym), lens, counters(step), common.NONCE, args.bsgs, ptr(d_out))
        with torch.cuda.stream(stream):
            dig = d_out.view(B, -1).sum(dim=1)
            if world > 1:
                dist.gather(dig, list(digests.unbind(0)) if rank == 0 else None, dst=0)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- correctness spot-check (untimed): block 0 of step 0 must decrypt to the PASTA plaintext ----
    checked = "skipped (oracle/_ref absent: random key material)"
    for s in range(args.warmup):
        device_step(s)
        if s == 0 and rank == 0 and ref is not None and not os.environ.get("HHE_BENCH_NO_VERIFY"):  # (timing-only kernel experiments)
            ctx.sync()
            from oracle import oracle as O
            got = d_out[0].cpu().numpy().view(np.uint64)
            slots, budget = ref.decrypt(got)
            want = O.pasta_plain(sym_key, common.T, sym_np[0], decrypt=True)  # counter 0
            assert np.array_equal(slots[:128], want), "bench output does not decrypt to the PASTA plaintext"
            checked = f"block 0 decrypts (SEAL) to the PASTA plaintext, noise budget {budget} bits"

    # ---- timed region: K device-resident steps ----
    ctx.profile(True)
    ctx.profile_reset()
    l0 = ctx.launch_count()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for s in range(args.steps):
        device_step(args.warmup + s)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count() - l0
    prof = ctx.profile_report()
    ctx.profile(False)
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    def host_step(step):
        first = int(counters(step)[0])
        rc = ctx.lib.hhe_pasta3_decompose(ctx.h, C.cast(C.c_void_p(ek_host.data_ptr()), C.POINTER(C.c_uint64)),
                                          C.cast(C.c_void_p(sym_host.data_ptr()), C.POINTER(C.c_uint64)), C.c_size_t(B * 128),
                                          C.c_uint64(common.NONCE), C.c_uint64(first), int(args.bsgs),
                                          C.cast(C.c_void_p(out_host.data_ptr()), C.POINTER(C.c_uint64)))
        ctx._chk(rc)
    ek_host = torch.from_numpy(enc_key.view(np.int64).copy()).pin_memory()
    host_step(0)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        host_step(args.warmup + s)
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t_e.item())
    if ref is not None and rank == 0:
        same = np.array_equal(out_host[0].numpy().view(np.uint64), d_out[0].cpu().numpy().view(np.uint64))
        assert same, "host-API and device-API outputs differ"

    # ---- "NTT GB/s" (the second figure BASELINE.json's metric names): forward / inverse negacyclic NTT of every limb of the output batch (621 MB at 296 blocks,
    # larger than L2) in HBM, 2 x 8N bytes per limb-transform; CUDA events on the engine's stream, rank 0 only ----
    ntt_info = None
    if rank == 0:
        limbs = 2 * L * B
        d_l = d_out.view(-1)[: limbs * N]
        rngl = torch.Generator(device="cuda").manual_seed(7)
        d_l.copy_(torch.randint(0, int(common.Q_16384[0]), (limbs * N,), dtype=torch.int64, device="cuda", generator=rngl))
        torch.cuda.synchronize()
        res = {}
        for name, inv in (("fwd", False), ("inv", True)):
            for _ in range(3):
                ctx.dev_ntt(0, inv, ptr(d_l), limbs)
            ctx.sync()
            n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0.record(stream)
            for _ in range(10):
                ctx.dev_ntt(0, inv, ptr(d_l), limbs)
            n1.record(stream)
            ctx.sync()
            t = n0.elapsed_time(n1) / 10
            res[name] = {"GBps": limbs * 16 * N / t / 1e6, "limb_transforms_per_s": limbs / t * 1e3, "frac_of_hbm_peak": limbs * 16 * N / t / 1e6 / peak_hbm()[0]}
        ntt_info = {"limbs_per_launch": limbs, "bytes_per_limb": 16 * N, "fwd": res["fwd"], "inv": res["inv"],
                    "note": "exact 49-bit modular transforms on the FP64 pipe (7 FP64 + 1 FRND per butterfly): pipe ceiling = 23 M limbs/s = 6.1 TB/s equivalent"}
    # ---- sharding invariance on the hardware: the digests rank 0 gathered in the last timed step came from
    # every rank's own GPU; rank 0 now transciphers the first blocks of every OTHER rank's shard itself (same counters, same words,
    # regenerated from that rank's seed) and the 64-bit digests must agree: a block's ciphertext depends on (counter, words) only ----
    invariance = None
    if world > 1:
        device_step(args.warmup + args.steps - 1)  # digests of a known step (the e2e loop reused d_out)
        barrier()
        if rank == 0:
            per = 2
            last = args.warmup + args.steps - 1
            ctrs, words = [], []
            for r in range(1, world):
                base = (last * world + r) * B
                ctrs += list(range(base, base + per))
                words.append(np.random.default_rng(1000 + r).integers(0, common.T, (B, 128), dtype=np.uint64)[:per])
            words = np.concatenate(words)
            d_w = torch.from_numpy(words.view(np.int64)).cuda()
            d_o = torch.empty((len(ctrs), 2, L, N), dtype=torch.int64, device="cuda")
            ctx.dev_pasta3_decompose(ptr(ek_t), ptr(d_w), np.full(len(ctrs), 128, dtype=np.uint32), np.array(ctrs, dtype=np.uint64),
                                     common.NONCE, args.bsgs, ptr(d_o))
            ctx.sync()
            mine = d_o.view(len(ctrs), -1).sum(dim=1).view(world - 1, per)
            theirs = digests[1:, :per]
            invariance = {"ranks_checked": world - 1, "blocks_per_rank": per, "digests_match": bool(torch.equal(mine, theirs)),
                          "what": "rank 0 re-transciphered blocks of every other rank's shard; 64-bit digests over all limbs compared"}
            del d_o, d_w
        barrier()

    # ---- strong scaling: a FIXED stream of blocks split over the ranks (BASELINE configs[3]: 65,536 blocks = --stream 65536) ----
    strong = None
    if not args.no_configs or args.stream:
        total = args.stream if args.stream else 8 * B
        per_rank = (total + world - 1) // world
        lo, hi = min(total, rank * per_rank), min(total, (rank + 1) * per_rank)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for off in range(lo, hi, B):
            nb = min(B, hi - off)
            ctx.dev_pasta3_decompose(ptr(d_key), ptr(d_sym), lens[:nb], np.arange(off, off + nb, dtype=np.uint64), common.NONCE, args.bsgs,
                                     ptr(d_out))
        s1.record(stream)
        barrier()
        st_ms = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(st_ms, op=dist.ReduceOp.MAX)
        strong = {"scaling": "strong", "total_blocks": total, "blocks_per_rank": per_rank, "ms": float(st_ms.item()),
                  "value": total / float(st_ms.item()) * 1e3, "unit": UNIT,
                  "note": "contiguous counter ranges per rank, no data-path collective; BASELINE configs[3] in full is --stream 65536"}

    # ---- FC line: BASELINE configs[2] (ECG 128 -> 1, batch of 1024 one-block records that all restart at counter 0), samples sharded
    # over the ranks; per sample: transcipher, packed_enc_multiply, relinearize, encrypted_vec_sum(128) with the default Galois keys ----
    fc = None
    if extras:
        from tools import bench_configs as BC
        rngf = np.random.default_rng(33)
        S_all = args.ecg_samples - args.ecg_samples % world
        xs_all = rngf.integers(0, 256, (S_all, 128), dtype=np.uint64)
        wts = rngf.integers(-128, 128, 128)
        S = S_all // world
        xs = xs_all[rank * S:(rank + 1) * S]
        from oracle import oracle as O
        syms = np.stack([O.pasta_plain(sym_key, common.T, xs[i]) for i in range(S)])
        w_t = torch.empty((1, 2, L, N), dtype=torch.int64, device="cuda")
        if rank == 0:
            w_t.copy_(torch.from_numpy(ref.encrypt(np.mod(wts, common.T).astype(np.uint64)).view(np.int64)[None]))
        if world > 1:
            dist.broadcast(w_t, 0)
        enc_w1 = w_t.cpu().numpy().view(np.uint64)
        ctx.set_batch(0)
        # pinned host buffers for the two results (the transciphered records and the FC outputs), as for `e2e`
        fc_pin = [torch.empty((S, 2, L, N), dtype=torch.int64).pin_memory() for _ in range(2)]
        fc_bufs = [t_.numpy().view(np.uint64) for t_ in fc_pin]
        BC.config3_run(ctx, enc_key, syms[:min(S, 8)], enc_w1, fc_bufs)  # warm-up (arena, key paths)
        barrier()
        outs, dt = BC.config3_run(ctx, enc_key, syms, enc_w1, fc_bufs)
        outs = outs.copy()
        barrier()
        # the same call with every record transciphered on its own (the reference's behaviour: 1,024 evaluations of the same keystream
        # circuit), on a sample of the batch -- the default computes the keystream ciphertext once per distinct counter per call
        S_cmp = min(S, 296)
        os.environ["HHE_NO_SHARED_KEYSTREAM"] = "1"
        outs_cmp, dt_cmp = BC.config3_run(ctx, enc_key, syms[:S_cmp], enc_w1, fc_bufs)
        del os.environ["HHE_NO_SHARED_KEYSTREAM"]
        same_bits = bool(np.array_equal(outs_cmp, outs[:S_cmp]))
        del outs_cmp, fc_bufs, fc_pin
        t_f = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_f, op=dist.ReduceOp.MAX)
        ok = None
        if rank == 0:
            ok = all(int(ref.decrypt(outs[i, 0])[0][127]) == int(np.dot(xs[i].astype(np.int64), wts)) % common.T for i in (0, S // 2, S - 1))
            # credited bytes follow the operations actually executed: one keystream evaluation per rank per call (shared by its S
            # records), and per record the final add_plain (2 ct + pt) and the FC row
            nbytes = BYTES_PER_BLOCK[False] / S + int(4.125 * (1 << 20)) + BC.fc_row_bytes(128)
            fc = {"workload": "BASELINE configs[2]: ECG 128->1, samples sharded over the ranks, pinned host buffers in and out (records -> host -> FC layer -> host)", "samples": S_all,
                  "samples_per_rank": S, "s": float(t_f.item()), "value": S_all / float(t_f.item()), "unit": "samples/s",
                  "parity_decrypted_dot_products": bool(ok),
                  "keystream_sharing": {
                      "what": "every ECG record restarts at counter 0 (CSP.cpp:247-252), so the keystream ciphertext -- which depends on "
                              "(encrypted key, nonce, counter) only, and SEAL's evaluation is deterministic -- is the same object for all "
                              "records of a call: it is evaluated once per distinct counter and each record only adds its own encoded words "
                              "(bit-identical outputs). The reference evaluates the keystream circuit once per record.",
                      "per_record_keystream_samples_per_s_rank0": S_cmp / dt_cmp, "per_record_sample": S_cmp,
                      "outputs_bit_identical_to_per_record_evaluation": same_bits},
                  "parity_note": "FC row limb-exact vs the reference: tests/test_gpu_fc.py::test_ecg_row_bit_exact_with_seal",
                  "algorithmic_bytes_per_sample": nbytes, "reference_op_sequence_bytes_per_sample": BYTES_PER_BLOCK[False] + BC.fc_row_bytes(128),
                  "frac_of_hbm_roofline": nbytes * S_all / world / float(t_f.item()) / 1e9 / peak_hbm()[0]}
        ctx.set_batch(args.blocks)
        del outs

    # ---- the final gather of the result ciphertexts themselves to rank 0 over NCCL/NVLink, outside the timed steps:
    # the steps gather digests only (65,536 result ciphertexts are 128 GiB); a caller that wants the ciphertexts pays this once ----
    gather_info = None
    if world > 1:
        shard = __import__("importlib").import_module(common.PKG + ".shard")
        shard.gather_ciphertexts(d_out, world)  # warm-up (communicator setup)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        got = shard.gather_ciphertexts(d_out, world)
        g1.record()
        torch.cuda.synchronize()
        g_ms = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(g_ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            nbytes = (world - 1) * d_out.numel() * 8
            gather_info = {"ciphertexts": world * B, "bytes_received_by_rank0": nbytes, "ms": float(g_ms.item()),
                           "GBps": nbytes / float(g_ms.item()) / 1e6, "rank0_block0_matches": bool(torch.equal(got[0], d_out[0]))}
        del got
    # ---- BASELINE configs[4] at N GPUs: the primitive sweep is replicated (every rank runs independent batches of the same shape on
    # its own GPU, no communication); rates are summed over the ranks. At one GPU the sweep is part of `configs` below. ----
    sweep_agg = None
    if world > 1 and not args.no_configs:
        from tools import bench_configs as BC
        barrier()
        loc = BC.config5(stream, peak_hbm()[0], None, sizes=(8192, 16384, 32768))
        names = [(n, op, f) for n in sorted(loc) for op, f in (("ntt_fwd", "GBps"), ("ntt_inv", "GBps"), ("rotate_rows", "per_s"),
                                                                ("relinearize", "per_s"), ("multiply", "per_s"))]
        vals = torch.tensor([loc[n][op][f] for n, op, f in names], dtype=torch.float64, device="cuda")
        dist.all_reduce(vals, op=dist.ReduceOp.SUM)
        if rank == 0:
            sweep_agg = {"ranks": world, "scaling": "weak (replicated batches, rates summed over ranks)"}
            for (n, op, f), v in zip(names, vals.tolist()):
                sweep_agg.setdefault(n, {})[f"{op}_{f}"] = v
                sweep_agg[n][f"{op}_{f}_rank0"] = loc[n][op][f]
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline of the dominant kernel ----
    # ks_digits is NOT HBM-bound: it is bound by the FP64 pipe (exact 49-bit modular products are 5 FP64 + 1 FRND instructions) with
    # the L2->SM path as co-limiter (profiles/r2_ksdigits_ncu.txt). The headline fraction is therefore the FP64-pipe one; the HBM view
    # (SURVEY.md 8(d)'s algorithmic bytes: one key switch = 2 ciphertexts = 4 MiB per item) is reported beside it.
    peak, peak_src = peak_hbm()
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
    ks = prof.get("ks_digits", {"launches": 0, "ms": 0.0})
    avg_ms = ks["ms"] / max(1, ks["launches"])
    hbm_achieved = KS_BYTES * B / (avg_ms * 1e-3) / 1e9 if avg_ms else 0.0
    ncu = None
    tpath = os.path.join(ROOT, "profiles", "ks_digits_traffic.json")
    if os.path.exists(tpath):
        ncu = json.load(open(tpath))
    cap_items = (ncu or {}).get("capture_items", 148)
    if ncu and "dram_bytes_per_launch" not in ncu and "dram_bytes_per_launch_at_capture_batch_148" in ncu:  # round-1 file layout
        ncu["dram_bytes_per_launch"] = ncu["dram_bytes_per_launch_at_capture_batch_148"]
    # the capture ran at `capture_items` items per launch; DRAM traffic scales with the items of a launch
    traffic = int(ncu["dram_bytes_per_launch"] * B / cap_items) if ncu and "dram_bytes_per_launch" in ncu else None
    # FP64 warp instructions per item: counted by ncu (smsp__inst_executed_pipe_fp64.sum of the captured launch / its items) when the
    # capture carries it, otherwise the count derived from the kernel's pass structure (they agree within 2 %)
    fp64_warp_per_item = (ncu or {}).get("fp64_warp_inst_per_item") or KS_FP64_PER_ITEM / 32.0
    fp64_src = "ncu smsp__inst_executed_pipe_fp64.sum (profiles/ks_digits_traffic.json)" if (ncu or {}).get("fp64_warp_inst_per_item") \
        else "derived from the kernel's pass structure"
    kernel_ms = sum(v["ms"] for v in prof.values())
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp64_peak = 148 * 4 * 0.5 * sm_mhz * 1e6  # warp instructions/s: one per 2 cycles per SM sub-partition (tools/microbench/pipes.cu)
    fp64_ach = fp64_warp_per_item * B / (avg_ms * 1e-3) if avg_ms else 0.0
    roofline = {
        "bound": "fp64_pipe", "kernel": "ks_digits (key-switch digit NTT + key inner product)",
        "achieved": fp64_ach / 1e12, "peak": fp64_peak / 1e12, "unit": "T warp-inst/s (FP64 pipe)", "frac": fp64_ach / fp64_peak if avg_ms else None,
        "traffic": traffic, "peak_source": "148 SMs x 4 sub-partitions x 1 FP64 warp instruction per 2 cycles x the SM clock sampled during the run "
                                           "(microbench: profiles/r1_pipes_microbench.txt)",
        "fp64_warp_inst_per_item": fp64_warp_per_item, "fp64_count_source": fp64_src,
        "avg_launch_ms": avg_ms, "launches": ks["launches"], "items_per_launch": B,
        "share_of_kernel_time": ks["ms"] / kernel_ms if kernel_ms else None, "dominant_by_time": dom[0],
        "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": KS_BYTES * B, "peak_source": peak_src,
                "note": "algorithmic bytes (one key switch = 2 ciphertexts per item) over the live launch time: small by construction, the "
                        "kernel's operands are L2-resident"},
        "l2_to_sm": ({"bytes_per_launch": int(ncu["l2_to_sm_bytes_per_launch"] * B / cap_items),
                      "GBps": ncu["l2_to_sm_bytes_per_launch"] * B / cap_items / (avg_ms * 1e-3) / 1e9 if avg_ms else None,
                      "note": "every CTA streams its digits (128 KiB), key slice (128 KiB) and last-pass twiddles (56 KiB) per digit from the L2: "
                              "the second limiter next to the FP64 pipe"} if ncu and "l2_to_sm_bytes_per_launch" in ncu else None),
        "ncu": ({k: ncu[k] for k in ("fp64_pipe_active_pct", "issue_active_pct", "lsu_wavefront_pipe_pct", "dram_throughput_pct", "source") if k in ncu}
                if ncu else None),
        "step_level": {"algorithmic_bytes_per_block": BYTES_PER_BLOCK[bool(args.bsgs)],
                       "achieved": BYTES_PER_BLOCK[bool(args.bsgs)] * value / world / 1e9,
                       "frac": BYTES_PER_BLOCK[bool(args.bsgs)] * value / world / 1e9 / peak, "unit": "GB/s", "peak": peak},
        "kernel_ms": {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
    }
    # ---- CPU baseline: the reference itself on this box's host cores (bounded sample) ----
    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        if ref is not None:
            secs = ref.bench_decompose(enc_key, cores, 1, args.bsgs)
            cpu = {"value": cores / secs, "unit": UNIT, "cores": cores, "kind": "reference", "sample_s": secs,
                   "sample": f"{cores} threads x 1 block, one pasta::PASTA_SEAL per thread (reference sources + libseal-4.0.a), {secs:.1f} s"}
        else:
            from oracle import oracle as O
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "oracle/_ref absent; run tests for the port"}
    # ---- secondary configs (one GPU): 1, 2, 5 and the BSGS mode (config 3 is the `fc` line above, config 4 the headline) ----
    configs = None
    if extras and world == 1:
        from tools import bench_configs as BC
        rngc = np.random.default_rng(8)
        ctx.set_batch(0)
        blk_s = cpu["sample_s"] if cpu and cpu.get("sample_s") else ref.bench_decompose(enc_key, 1, 1, False)
        ref_ops = {"block_s": blk_s, "rotate_s": ref.bench_primitive(enc_key, 2, 3), "relinearize_s": ref.bench_primitive(enc_key, 3, 2),
                   "multiply_plain_s": ref.bench_primitive(enc_key, 4, 2), "multiply_s": ref.bench_primitive(enc_key, 5, 2)}
        configs = {"reference_ops_1core_s": ref_ops, "config1_one_block": BC.config1(ctx, ref, enc_key, sym_key, rngc, peak)}
        # BSGS mode on the headline batch, device resident
        ctx.set_batch(args.blocks)
        from oracle import oracle as O
        for _ in range(2):
            ctx.dev_pasta3_decompose(ptr(d_key), ptr(d_sym), lens, counters(0), common.NONCE, True, ptr(d_out))
        ctx.sync()
        got0 = d_out[0].cpu().numpy().view(np.uint64)  # counter 0 (one GPU): must decrypt to the PASTA plaintext of block 0's words
        dec_ok = bool(np.array_equal(ref.decrypt(got0)[0][:128], O.pasta_plain(sym_key, common.T, sym_np[0], decrypt=True)))
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(stream)
        for s_ in range(2):
            ctx.dev_pasta3_decompose(ptr(d_key), ptr(d_sym), lens, counters(1 + s_), common.NONCE, True, ptr(d_out))
        b1.record(stream)
        ctx.sync()
        bs = b0.elapsed_time(b1) * 1e-3 / 2
        configs["bsgs_mode"] = {"blocks_per_s": B / bs, "ms_per_step": bs * 1e3, "blocks": B, "parity_note": "limb-exact vs the reference with "
                                "use_bsgs=true: config1_one_block.bsgs here and tests/test_gpu_fullsize.py::test_bsgs_block_bit_exact_with_seal",
                                "block_decrypts": dec_ok, "algorithmic_bytes_per_block": BYTES_PER_BLOCK[True],
                                "frac_of_hbm_roofline": BYTES_PER_BLOCK[True] * B / bs / 1e9 / peak}
        ctx.set_batch(0)
        configs["config2_mnist_sample"] = BC.config2(ctx, ref, enc_key, sym_key, rngc, peak, ref_ops)
        configs["next_rows"] = BC.next_rows(ctx, ref, enc_key, sym_key, rngc, ref_ops)
        if fc:
            per_sample = ref_ops["block_s"] + ref_ops["multiply_s"] + ref_ops["relinearize_s"] + BC.KS_COUNT[128] * ref_ops["rotate_s"]
            fc["reference_1core_s_per_sample"] = per_sample
            fc["reference_kind"] = "composed from the reference's per-operation times measured in this run (block + multiply + relinearize + 355 key switches)"
            fc["speedup_vs_1core"] = fc["value"] * per_sample
            configs["config3_ecg_batch"] = "see the top-level `fc` object (the same measurement, sharded over the ranks at N > 1)"
        ctx.close()
        fac = lambda n_: R.Ref(n_, common.T, None, seed=5, steps=(0, -1), default_gk=False)  # noqa: E731
        configs["config5_primitive_sweep"] = BC.config5(stream, peak, fac)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": config(args, world), "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ek_host.numel() * 8 + sym_host.numel() * 8 + B * 12),
                "d2h_bytes_per_step": int(out_host.numel() * 8),
                "note": "host clock around the blocking C-ABI call on pinned host buffers (H2D of key + words, D2H of every output ciphertext "
                        "inside); `value`'s timed region additionally records two CUDA events per kernel launch for the live per-kernel "
                        "table (about 1 %), which is why e2e can come out marginally above it"},
        "gpu_launches": int(launches), "clocks": clocks, "verified": checked, "ntt": ntt_info, "gather": gather_info,
        "sharding_invariance": invariance, "strong": strong, "fc": fc, "configs": configs, "primitive_sweep_all_ranks": sweep_agg,
    }
    emit(out)


if __name__ == "__main__":
    main()

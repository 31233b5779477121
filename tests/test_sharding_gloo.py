"""world_size-2 (gloo, CPU) test of the multi-GPU path's host logic: contiguous counter ranges per rank, each rank
transciphers its shard independently (here through the emulation harness), rank 0 gathers per-block digests; the
result must equal the single-rank run block for block (sharding invariance)."""
import os
import subprocess
import sys

import numpy as np

import common

WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["HHE_ROOT"]); sys.path.insert(0, os.path.join(os.environ["HHE_ROOT"], "tests"))
import common, importlib
pkg = common.package(); shard = importlib.import_module(common.PKG + ".shard")
fx = np.load(os.path.join(common.ROOT, "tests", "golden", "pasta_n512.npz"))
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["HHE_PORT"], rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
ctx = pkg.Context(int(fx["N"]), int(fx["t"]), fx["q"], lib_path=os.path.join(common.ROOT, "tests", "emul", "libhhe_emul.so"), emulation_harness=True)
for name, kind in (("gk_m1", 0), ("gk_p128", 0), ("gk_col", 0), ("rk", 2)):
    ctx.load_ksk(kind, int(fx[name + "_elt"]), fx[name])
words, first, nblk = shard.shard_stream(fx["sym_ct"], rank, world)
out = ctx.pasta3_decompose(fx["enc_key"], words, first_counter=first) if nblk else np.zeros((0, 2, ctx.L, ctx.N), dtype=np.uint64)
dig = shard.gather_digests(shard.digest(out) if nblk else np.zeros(0, dtype=np.uint64), world)
cts = shard.gather_ciphertexts(out, world)  # the final gather of the result ciphertexts themselves
if rank == 0:
    np.save(os.environ["HHE_OUT"], dig)
    np.save(os.environ["HHE_OUT"] + ".cts.npy", cts)
else:
    assert cts is None
dist.destroy_process_group()
'''


def test_block_ranges_partition_everything():
    shard = __import__("importlib").import_module(common.PKG + ".shard")
    for total in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            rs = [shard.block_range(total, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_equal_one_rank(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(common.ROOT, "tests", "emul")])
    fx = np.load(os.path.join(common.ROOT, "tests", "golden", "pasta_n512.npz"))
    shard = __import__("importlib").import_module(common.PKG + ".shard")
    want = shard.digest(fx["decomposed"])  # golden: the reference's own output for the whole stream
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "digests.npy"
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", HHE_ROOT=common.ROOT, HHE_PORT="29653", HHE_OUT=str(out))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    for p in procs:
        assert p.wait(timeout=600) == 0
    got = np.load(out)
    assert np.array_equal(got, want)
    assert np.array_equal(np.load(str(out) + ".cts.npy"), fx["decomposed"])  # gathered ciphertexts == the reference's, in block order


RAGGED_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["HHE_ROOT"]); sys.path.insert(0, os.path.join(os.environ["HHE_ROOT"], "tests"))
import common, importlib
shard = importlib.import_module(common.PKG + ".shard")
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["HHE_PORT"], rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
total = 5  # 5 units over 2 ranks: 3 + 2
lo, hi = shard.block_range(total, rank, world)
local = (np.arange(lo, hi, dtype=np.uint64)[:, None, None] * 1000 + np.arange(6, dtype=np.uint64).reshape(2, 3)[None])
out = shard.gather_ciphertexts(local, world)
empty = shard.gather_ciphertexts(local[:0] if rank else local, world)  # a rank with nothing to contribute
if rank == 0:
    want = np.arange(total, dtype=np.uint64)[:, None, None] * 1000 + np.arange(6, dtype=np.uint64).reshape(2, 3)[None]
    assert np.array_equal(out, want), out
    assert np.array_equal(empty, want[:3]), empty
    open(os.environ["HHE_OUT"], "w").write("ok")
dist.destroy_process_group()
'''


def test_gather_ciphertexts_ragged_shards(tmp_path):
    script = tmp_path / "ragged.py"
    script.write_text(RAGGED_WORKER)
    out = tmp_path / "ok.txt"
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", HHE_ROOT=common.ROOT, HHE_PORT="29654", HHE_OUT=str(out))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    assert out.read_text() == "ok"


RECORDS_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["HHE_ROOT"]); sys.path.insert(0, os.path.join(os.environ["HHE_ROOT"], "tests"))
import common, importlib
pkg = common.package(); shard = importlib.import_module(common.PKG + ".shard")
fx = np.load(os.path.join(common.ROOT, "tests", "golden", "pasta_n512.npz"))
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["HHE_PORT"], rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
ctx = pkg.Context(int(fx["N"]), int(fx["t"]), fx["q"], lib_path=os.path.join(common.ROOT, "tests", "emul", "libhhe_emul.so"), emulation_harness=True)
for name, kind in (("gk_m1", 0), ("gk_p128", 0), ("gk_col", 0), ("rk", 2)):
    ctx.load_ksk(kind, int(fx[name + "_elt"]), fx[name])
# 5 one-block records that all restart at counter 0 (the ECG batch of BASELINE configs[2]), sharded by record: 3 + 2
R, n = 5, 100
recs = np.random.default_rng(5).integers(0, int(fx["t"]), (R, n), dtype=np.uint64)
lo, hi = shard.block_range(R, rank, world)
mine = ctx.pasta3_decompose(fx["enc_key"], recs[lo:hi].reshape(-1), records=hi - lo)  # keystream shared by this rank's records
got = shard.gather_ciphertexts(mine, world)
if rank == 0:
    whole = ctx.pasta3_decompose(fx["enc_key"], recs.reshape(-1), records=R)
    single = np.stack([ctx.pasta3_decompose(fx["enc_key"], recs[r])[0] for r in range(R)])  # one call per record: no sharing at all
    assert np.array_equal(got, whole) and np.array_equal(got, single)
    open(os.environ["HHE_OUT"], "w").write("ok")
dist.destroy_process_group()
'''


def test_records_sharded_over_two_ranks(tmp_path):
    """Sample sharding (the `fc` line of bench.py at N > 1): records that restart their counters are split over the ranks, every rank
    evaluates the shared keystream for its own call; the gathered ciphertexts equal the one-rank call and per-record calls bit for bit."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(common.ROOT, "tests", "emul")])
    script = tmp_path / "records.py"
    script.write_text(RECORDS_WORKER)
    out = tmp_path / "ok.txt"
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", HHE_ROOT=common.ROOT, HHE_PORT="29655", HHE_OUT=str(out))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    for p in procs:
        assert p.wait(timeout=900) == 0
    assert out.read_text() == "ok"

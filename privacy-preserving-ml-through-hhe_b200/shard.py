"""Sharding of independent PASTA blocks / samples over the ranks of one node: contiguous ranges, no
data-path collective; only per-unit digests (or, on request, result ciphertexts) are gathered to rank 0."""
import numpy as np


def block_range(total, rank, world):
    """Contiguous range [lo, hi) of `total` units owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_stream(sym_ct, rank, world, first_counter=0):
    """One symmetric-ciphertext stream -> this rank's words, its first SHAKE counter and its block count."""
    sym_ct = np.asarray(sym_ct, dtype=np.uint64)
    nblocks = (len(sym_ct) + 127) // 128
    lo, hi = block_range(nblocks, rank, world)
    return sym_ct[lo * 128:min(hi * 128, len(sym_ct))], first_counter + lo, hi - lo


def digest(cts):
    """64-bit digest per ciphertext (wrapping sum of its words): what rank 0 gathers instead of 2 MiB ciphertexts."""
    a = np.asarray(cts, dtype=np.uint64)
    return a.reshape(a.shape[0], -1).sum(axis=1, dtype=np.uint64)


def gather_digests(local, world):
    """torch.distributed gather of per-block digests to rank 0 (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(local).view(np.int64))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.numel()], dtype=torch.int64, device=t.device))
    bufs = [torch.zeros(int(s.item()), dtype=torch.int64, device=t.device) for s in sizes]
    dist.all_gather(bufs, t) if len({int(s.item()) for s in sizes}) == 1 else _uneven_all_gather(bufs, t)
    return np.concatenate([b.cpu().numpy().view(np.uint64) for b in bufs])


def gather_ciphertexts(local, world, dst=0):
    """The final gather of result ciphertexts to rank `dst` (NCCL over NVLink on GPUs, gloo in the CPU tests).
    `local`: this rank's ciphertexts, a numpy uint64 array or a torch int64 tensor (device tensors stay on the device),
    shape [n_r, ...]. Ranks may hold different counts. Returns the concatenation in rank order on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist

    as_numpy = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local).view(np.int64)) if as_numpy else local.contiguous()
    if dist.get_backend() == "nccl" and not t.is_cuda:
        t = t.cuda()
    rank = dist.get_rank()
    counts = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device))
    counts = [int(c.item()) for c in counts]
    tail = tuple(t.shape[1:])
    if len(set(counts)) == 1:
        bufs = [torch.empty((counts[0],) + tail, dtype=t.dtype, device=t.device) for _ in range(world)] if rank == dst else None
        dist.gather(t, bufs, dst=dst)
    else:  # ragged shards: point-to-point into rank dst
        bufs = None
        if rank == dst:
            bufs = [torch.empty((c,) + tail, dtype=t.dtype, device=t.device) for c in counts]
            bufs[dst].copy_(t)
            for src in range(world):
                if src != dst and counts[src]:
                    dist.recv(bufs[src], src=src)
        elif t.shape[0]:
            dist.send(t, dst=dst)
    if rank != dst:
        return None
    out = torch.cat(bufs, dim=0)
    return out.cpu().numpy().view(np.uint64) if as_numpy else out


def _uneven_all_gather(bufs, t):
    import torch.distributed as dist

    for src, b in enumerate(bufs):
        if src == dist.get_rank():
            b.copy_(t)
        dist.broadcast(b, src)

"""Read-sharded multi-GPU driver: anchors are independent units, the packed reference is
replicated on every GPU, each rank extends a contiguous range of reads and rank 0 concatenates the results.
There is no data-path collective; torch.distributed only carries the final gather (and the bench's barrier).

`compute(anchors, hit_pool) -> (res, ops)` is the per-rank worker -- in production
`Processor.extender_body` of the rank's GPU; the CPU tests inject the oracle so the plumbing runs under gloo.
"""
import numpy as np

from . import abi


def split_reads(anchors, world):
    """Contiguous anchor ranges, cut only where read_num changes, balanced by anchor count.
    Returns [(lo, hi)] * world (possibly empty ranges)."""
    n = len(anchors)
    if n == 0:
        return [(0, 0)] * world
    rn = anchors["read_num"]
    cuts = np.flatnonzero(rn[1:] != rn[:-1]) + 1          # legal cut positions
    bounds = [0]
    for r in range(1, world):
        target = n * r // world
        if len(cuts) == 0:
            pos = n if target > 0 else 0
        else:
            k = int(np.searchsorted(cuts, target))
            cand = [int(cuts[min(k, len(cuts) - 1)])]
            if k > 0:
                cand.append(int(cuts[k - 1]))
            pos = min(cand, key=lambda c: abs(c - target))
        bounds.append(max(pos, bounds[-1]))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def local_shard(anchors, hit_pool, rank, world):
    """This rank's anchors with hit offsets rebased onto a compact private hit pool."""
    lo, hi = split_reads(anchors, world)[rank]
    a = anchors[lo:hi].copy()
    if len(a) == 0:
        return a, np.zeros(0, np.uint64), (lo, hi)
    # vectorised gather of the shard's hit lists: [left_0 | right_0 | left_1 | right_1 | ...]
    off = np.stack([a["left_hits_off"], a["right_hits_off"]], 1).reshape(-1).astype(np.int64)
    cnt = np.stack([a["left_hits_n"], a["right_hits_n"]], 1).reshape(-1).astype(np.int64)
    start = np.cumsum(cnt) - cnt
    total = int(cnt.sum())
    idx = np.repeat(off - start, cnt) + np.arange(total, dtype=np.int64)
    hp = hit_pool[idx] if total else np.zeros(0, np.uint64)
    a["left_hits_off"] = start[0::2]
    a["right_hits_off"] = start[1::2]
    return a, hp, (lo, hi)


def extend_sharded(compute, anchors, hit_pool, rank=0, world=1, group=None):
    """Run `compute` on this rank's shard; rank 0 returns (res, ops) for ALL anchors in input order, other ranks None."""
    a, hp, _ = local_shard(anchors, hit_pool, rank, world)
    if len(a):
        res, ops = compute(a, hp)
        used = int((res["ops_offset"] + res["n_ops"]).max()) if len(res) else 0
        ops = np.ascontiguousarray(ops[:used])
    else:
        res, ops = np.zeros(0, abi.ALN_RES), np.zeros(0, np.uint8)
    if world == 1:
        return res, ops
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")

    def gather_bytes(arr):
        raw = torch.from_numpy(np.frombuffer(arr.tobytes(), np.uint8).copy()).to(dev)
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([raw.numel()], dtype=torch.int64, device=dev), group=group)
        sizes = [int(s.item()) for s in sizes]
        mx = max(max(sizes), 1)
        padded = torch.zeros(mx, dtype=torch.uint8, device=dev)
        padded[:raw.numel()] = raw
        bufs = [torch.zeros(mx, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(padded, bufs, dst=0, group=group)
        if rank != 0:
            return None
        return [b[:s].cpu().numpy() for b, s in zip(bufs, sizes)]

    res_parts = gather_bytes(res)
    ops_parts = gather_bytes(ops)
    if rank != 0:
        return None
    out_res, out_ops, base = [], [], 0
    for rb, ob in zip(res_parts, ops_parts):
        r = rb.view(abi.ALN_RES).copy()
        r["ops_offset"] += base
        base += len(ob)
        out_res.append(r)
        out_ops.append(ob)
    return np.concatenate(out_res), np.concatenate(out_ops) if out_ops else np.zeros(0, np.uint8)

"""Read-sharded extension across the GPUs of one box.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/extend_sharded.py [n_reads]
Every rank uploads the full arena replica, extends its contiguous read range (darwin_b200.shard), rank 0 gathers and --
as a check -- recomputes everything on its own GPU and compares."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import darwin_b200  # noqa: E402
from darwin_b200 import abi, shard, synth  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
arena, anchors, hits = synth.anchor_batch(11, n_reads, 10000, 4000000)      # same seeded inputs on every rank
p = darwin_b200.Processor(len(arena), local)
p.InitializeScoringParameters(abi.Scoring.from_values())
p.InitializeReferenceMemory(0, arena)                                        # replica of the packed reference + reads
p.extender_body(anchors[:32], hits, 384, 64, 0)
kernel_ms = []


def compute(a, hp):
    out = p.extender_body(a, hp, 384, 64, 0)
    kernel_ms.append(p.stats().last_kernel_ms)
    return out


torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
got = shard.extend_sharded(compute, anchors, hits, rank, world)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
wall = time.perf_counter() - t0
t = torch.tensor([sum(kernel_ms)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    res, ops = got
    cells = float(res["cells"].sum())
    print("world %d: %d reads, %.3g cells: max-over-ranks kernel %.1f ms -> %.0f GCUPS, %.0f reads/s; wall incl. gather %.1f ms" % (
        world, n_reads, cells, float(t[0]), cells / float(t[0]) / 1e6, n_reads / float(t[0]) * 1e3, wall * 1e3))
    whole_res, whole_ops = p.extender_body(anchors, hits, 384, 64, 0)
    same = all(whole_res[k][f] == res[k][f] for k in range(len(res)) for f in ("n_ops", "score", "reference_start_offset", "query_end_offset", "flags"))
    same = same and all(np.array_equal(whole_ops[int(a["ops_offset"]):int(a["ops_offset"]) + int(a["n_ops"])],
                                       ops[int(b["ops_offset"]):int(b["ops_offset"]) + int(b["n_ops"])]) for a, b in zip(whole_res[::37], res[::37]))
    print("sharded result identical to single-GPU result:", same)
p.close()
if world > 1:
    dist.destroy_process_group()

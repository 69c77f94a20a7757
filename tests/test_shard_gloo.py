"""CPU: the N>1 path (read sharding + gather on rank 0) under torch.distributed/gloo with world_size 2.
The per-rank worker is the oracle here (no GPU in this container); on the GPU box the same driver runs with
Processor.extender_body (tests/test_gpu_parity.py::test_sharded_extend_matches_single)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import oracle
from darwin_b200 import abi, shard
from test_host_logic import synthetic_anchor_set
from conftest import alignments_equal, ALN_FIELDS_OURS


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arena, anchors, hits = synthetic_anchor_set(9, 10, 1200, ref_len=16000)
    port_ = oracle.port(abi.Scoring.from_values())
    prm = abi.ExtendParams(128, 32, 0, 0)

    def compute(a, hp):
        return port_.extend(arena, prm, a, hp, oracle.Port.STREAM)

    got = shard.extend_sharded(compute, anchors, hits, rank, world)
    if rank == 0:
        res, ops = got
        np.savez(out_path, res=res, ops=ops)
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_split_reads_cuts_on_read_boundaries():
    a = np.zeros(10, abi.ANCHOR)
    a["read_num"] = [0, 0, 0, 1, 1, 2, 2, 2, 2, 3]
    for world in (1, 2, 3, 4, 8):
        parts = shard.split_reads(a, world)
        assert parts[0][0] == 0 and parts[-1][1] == 10
        for (lo, hi), (lo2, _) in zip(parts, parts[1:]):
            assert hi == lo2
        for lo, hi in parts:
            if 0 < lo < 10:
                assert a["read_num"][lo] != a["read_num"][lo - 1]


def test_world2_gloo_matches_single_process(tmp_path):
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    arena, anchors, hits = synthetic_anchor_set(9, 10, 1200, ref_len=16000)
    res, ops = oracle.port(abi.Scoring.from_values()).extend(arena, abi.ExtendParams(128, 32, 0, 0), anchors, hits,
                                                             oracle.Port.STREAM)
    assert alignments_equal(res, ops, got["res"], got["ops"], ALN_FIELDS_OURS) == []

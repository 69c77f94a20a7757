"""GPU: D-SOFT seeding on the device (darwin_gpu_seed_index / darwin_gpu_seed, csrc/dsoft.cuh) against the oracle's
restatement and the compiled reference's own seeder_body: seed position table, anchors and chained hits identical."""
import os
import sys

import numpy as np
import pytest

import oracle
from darwin_b200 import abi, synth
from conftest import GOLDEN
from test_oracle_dsoft import same_seed_output, strand_views

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not built")]


def _case(seed, n_reads, overlap):
    sys.path.insert(0, GOLDEN)
    import make_golden
    ref, n_reads = make_golden.filter_case(seed, n_reads)
    ref.set_extend(384, 64, 2, overlap)
    return ref, n_reads


@pytest.mark.parametrize("seed,overlap", [(11, 0), (12, 1)])
def test_gpu_seeding_matches_reference(gpu, seed, overlap):
    ref, n_reads = _case(seed, 40, overlap)
    try:
        ref.seed(0, n_reads)
        begin, anchors, pool = ref.seed_anchors()
        arena = ref.arena().copy()
        prm = ref.seed_params()
        p = gpu(len(arena), abi.Scoring.from_values())
        p.InitializeReferenceMemory(0, arena)
        ref_size = int(ref.lib.dref_arena_reference_size())
        p.build_seed_index(prm, ref.chroms(), ref_size)
        # the table itself: buckets equal, positions equal wherever D-SOFT may read them
        dp = oracle.DsoftPort(np.concatenate([arena, np.full(256, ord("N"), np.uint8)]), ref.chroms(), ref_size, prm)
        wb, wp = dp.index_arrays()
        gb, gp, max_occ = p.seed_index_arrays()
        assert max_occ == dp.ix.kmer_max_occurence and np.array_equal(gb, wb) and len(gp) == len(wp)
        small = np.nonzero((np.diff(wb.astype(np.int64)) > 0) & (np.diff(wb.astype(np.int64)) <= max_occ))[0]
        for b in small[:: max(1, len(small) // 20000)]:
            assert np.array_equal(gp[wb[b]:wb[b + 1]], wp[wb[b]:wb[b + 1]])
        dp.close()
        reads = np.zeros(n_reads, abi.SEED_READ)
        for r in range(n_reads):
            reads[r]["read_addr"], reads[r]["read_len"] = ref.read_addr(r), ref.lib.dref_read_len(r)
        gbeg, ganc, gpool = p.seeder_body(reads)
        assert np.array_equal(gbeg, begin), (gbeg[:10], begin[:10])
        want, got = strand_views(begin, anchors, n_reads), strand_views(gbeg, ganc, n_reads)
        for r in range(n_reads):
            for s in (0, 1):
                assert same_seed_output(got[r][s], gpool, want[r][s], pool), (r, s)
        assert len(anchors) > 30
        p.close()
    finally:
        ref.set_extend(384, 64, 2, 0)

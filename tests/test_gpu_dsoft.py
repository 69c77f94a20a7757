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


@pytest.mark.parametrize("sort", ["smem", "cub"])
@pytest.mark.parametrize("seed,overlap", [(11, 0), (12, 1)])
def test_gpu_seeding_matches_reference(gpu, seed, overlap, sort, monkeypatch):
    """Both sort paths of the seeding call: one CTA per segment in shared memory (default) and CUB's segmented sort (the
    fallback for strands / windows that do not fit shared memory; DARWIN_GPU_SEED_SORT is read when the handle is made)."""
    monkeypatch.setenv("DARWIN_GPU_SEED_SORT", sort)
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


def test_align_reads_equals_staged_calls(gpu):
    """darwin_gpu_align_reads (resident pipeline) vs the same stages called one by one with the chained hits travelling
    through the host: seeder_body -> filter_body -> slope filter (oracle restatement) -> extender_body."""
    import darwin_b200
    ref, n_reads = _case(13, 36, 0)
    ref.seed(0, n_reads)
    arena = ref.arena().copy()
    sc = abi.Scoring.from_values()
    p = gpu(len(arena), sc)
    p.InitializeReferenceMemory(0, arena)
    chroms = ref.chroms()
    ref_size = int(ref.lib.dref_arena_reference_size())
    p.build_seed_index(ref.seed_params(), chroms, ref_size)
    reads = np.zeros(n_reads, abi.SEED_READ)
    for r in range(n_reads):
        reads[r]["read_addr"], reads[r]["read_len"] = ref.read_addr(r), ref.lib.dref_read_len(r)
    # staged
    begin, sanc, pool = p.seeder_body(reads)
    cands = np.zeros(len(sanc), abi.FILTER_CAND)
    rn = np.zeros(len(sanc), np.int32)
    starts = chroms["start"].astype(np.int64)
    padded = np.diff(np.concatenate([starts, [ref_size]]))
    for r in range(n_reads):
        for s in (0, 1):
            for i in range(begin[2 * r + s], begin[2 * r + s + 1]):
                hit, off = int(sanc[i]["hit_offset"]) >> 32, int(sanc[i]["hit_offset"]) & 0xFFFFFFFF
                c = int(np.searchsorted(starts, hit, side="right")) - 1
                cands[i] = (reads[r]["read_addr"], hit, off, starts[c], padded[c], reads[r]["read_len"], s, (0, 0, 0))
                rn[i] = r
    fres = p.filter_body(cands)
    port = oracle.port(sc)
    want_anchors, hp = [], []
    for s in (0, 1):
        m = np.nonzero((cands["strand"] == s) & ((fres["flags"] & 3) == 3))[0]
        for i in m[port.slope_filter(rn[m], fres["score"][m], fres["reference_pos"][m], fres["query_pos"][m])]:
            a = np.zeros(1, abi.ANCHOR)[0]
            a["read_addr"], a["read_len"], a["read_num"], a["strand"] = cands[i]["read_addr"], cands[i]["read_len"], rn[i], s
            a["reference_pos"], a["query_pos"], a["score"] = fres[i]["reference_pos"], fres[i]["query_pos"], fres[i]["score"]
            a["chr_start"], a["ref_len"] = cands[i]["chr_start"], cands[i]["chr_len"]
            a["chr_id"] = int(np.searchsorted(starts, int(cands[i]["hit"]), side="right")) - 1
            for side in ("left", "right"):
                lo, n = int(sanc[i][side + "_off"]), int(sanc[i][side + "_n"])
                a[side + "_hits_off"], a[side + "_hits_n"] = sum(len(x) for x in hp), n
                hp.append(pool[lo:lo + n])
            want_anchors.append(a)
    want_anchors = np.array(want_anchors, abi.ANCHOR)
    wres, wops = p.extender_body(want_anchors, np.concatenate(hp), 384, 64, 0)
    # resident
    ganc, gres, gops = p.align_reads(reads)
    assert len(ganc) == len(want_anchors) > n_reads // 2
    for f in ("read_addr", "reference_pos", "query_pos", "chr_start", "ref_len", "read_len", "read_num", "chr_id", "score",
              "left_hits_n", "right_hits_n", "strand"):
        assert np.array_equal(ganc[f], want_anchors[f]), f
    from conftest import alignments_equal, ALN_FIELDS_OURS
    assert alignments_equal(wres, wops, gres, gops, ALN_FIELDS_OURS) == []
    # caller-supplied buffers that are too small: the needed count comes back with DARWIN_ERR_CAPACITY
    small = (np.empty(2, abi.ANCHOR), np.empty(2, abi.ALN_RES), np.empty(1 << 20, np.uint8))
    with pytest.raises(darwin_b200.DarwinGpuError) as e:
        p.align_reads(reads, out=small)
    assert e.value.code == abi.ERR_CAPACITY
    p.close()


def test_gpu_seeding_long_reads_and_many_reads(gpu):
    """ONT-like 30-50 kbp reads (tens of minimizer chunks per strand, the stride-4 tail of the seed list) and reads shorter
    than one chunk, against the reference's seeder_body."""
    rng = np.random.default_rng(3)
    ref = oracle.reference("patched")
    ref.set_scoring(abi.Scoring.from_values())
    ref.set_dsoft_defaults()
    ref.set_extend(384, 64, 2, 0)
    ref.reset_arena()
    genome = synth.random_seq(rng, 400000)
    genome[250000:253000] = genome[50000:53000]                       # an exact repeat: multi-hit buckets
    ref.add_chr("chrL", genome.tobytes(), True)
    ref.build_index()
    lens = [50000, 41234, 30001, 2047, 2048, 2049, 300, 129, 100, 65 + 14]
    for k, L in enumerate(lens):
        p = int(rng.integers(0, len(genome) - L))
        r = synth.mutate_fast(rng, genome[p:p + L], 0.04, 0.03, 0.05)
        if k % 2:
            r = synth.revcomp(r)
        ref.add_read("r%d" % k, np.ascontiguousarray(r).tobytes())
    n_reads = ref.lib.dref_num_reads()
    ref.seed(0, n_reads)
    begin, anchors, pool = ref.seed_anchors()
    arena = ref.arena().copy()
    p = gpu(len(arena), abi.Scoring.from_values())
    p.InitializeReferenceMemory(0, arena)
    p.build_seed_index(ref.seed_params(), ref.chroms(), int(ref.lib.dref_arena_reference_size()))
    reads = np.zeros(n_reads, abi.SEED_READ)
    for r in range(n_reads):
        reads[r]["read_addr"], reads[r]["read_len"] = ref.read_addr(r), ref.lib.dref_read_len(r)
    gbeg, ganc, gpool = p.seeder_body(reads)
    assert np.array_equal(gbeg, begin)
    want, got = strand_views(begin, anchors, n_reads), strand_views(gbeg, ganc, n_reads)
    for r in range(n_reads):
        for s in (0, 1):
            assert same_seed_output(got[r][s], gpool, want[r][s], pool), (r, s, lens[r] if r < len(lens) else None)
    assert len(anchors) >= 6 and int(anchors["left_n"].max()) > 1000
    p.close()


@pytest.mark.parametrize("k,w,stride", [(12, 5, 4), (13, 9, 2)])
def test_gpu_seeding_other_seed_shapes(gpu, k, w, stride):
    from test_oracle_dsoft import small_case
    ref, n_reads = small_case(k, w, stride)
    try:
        ref.seed(0, n_reads)
        begin, anchors, pool = ref.seed_anchors()
        arena = ref.arena().copy()
        p = gpu(len(arena), abi.Scoring.from_values())
        p.InitializeReferenceMemory(0, arena)
        p.build_seed_index(ref.seed_params(), ref.chroms(), int(ref.lib.dref_arena_reference_size()))
        reads = np.zeros(n_reads, abi.SEED_READ)
        for r in range(n_reads):
            reads[r]["read_addr"], reads[r]["read_len"] = ref.read_addr(r), ref.lib.dref_read_len(r)
        gbeg, ganc, gpool = p.seeder_body(reads)
        assert np.array_equal(gbeg, begin)
        want, got = strand_views(begin, anchors, n_reads), strand_views(gbeg, ganc, n_reads)
        for r in range(n_reads):
            for s in (0, 1):
                assert same_seed_output(got[r][s], gpool, want[r][s], pool), (k, w, r, s)
        p.close()
    finally:
        ref.set_dsoft_defaults()


def test_seed_index_over_many_short_chromosomes(gpu):
    """De novo sized chromosome lists (more entries than a CUDA grid's y/z extent, 65 535): the table equals the CPU
    restatement's.  An empty list is refused."""
    import darwin_b200
    rng = np.random.default_rng(11)
    n_chr = 70001
    lens = rng.integers(40, 121, n_chr).astype(np.int64)
    lens[123] = 5000                                                    # one chromosome spans several chunks
    padded = (lens + 127) // 128 * 128
    starts = 128 + np.concatenate([[0], np.cumsum(padded)[:-1]])
    ref_size = int(starts[-1] + padded[-1])
    arena = np.full(ref_size, ord("N"), np.uint8)
    for s, L in zip(starts, lens):
        arena[s:s + L] = synth.random_seq(rng, int(L))
    chroms = np.zeros(n_chr, abi.CHROM)
    chroms["start"], chroms["len_unpadded"] = starts, lens
    prm = abi.SeedParams.stock()
    p = gpu(len(arena), abi.Scoring.from_values())
    p.InitializeReferenceMemory(0, arena)
    p.build_seed_index(prm, chroms, ref_size)
    dp = oracle.DsoftPort(np.concatenate([arena, np.full(256, ord("N"), np.uint8)]), chroms, ref_size, prm)
    wb, wp = dp.index_arrays()
    gb, gp, max_occ = p.seed_index_arrays()
    assert max_occ == dp.ix.kmer_max_occurence and np.array_equal(gb, wb) and len(gp) == len(wp)
    small = np.nonzero((np.diff(wb.astype(np.int64)) > 0) & (np.diff(wb.astype(np.int64)) <= max_occ))[0]
    for b in small[:: max(1, len(small) // 20000)]:
        assert np.array_equal(gp[wb[b]:wb[b + 1]], wp[wb[b]:wb[b + 1]])
    dp.close()
    with pytest.raises(darwin_b200.DarwinGpuError) as e:
        p.build_seed_index(prm, chroms[:0], ref_size)
    assert e.value.code == abi.ERR_INVALID


@pytest.mark.parametrize("overlap", [0, 1])
def test_seed_sort_paths_agree_at_scale(gpu, overlap, monkeypatch):
    """3 000 reads of 0.2 - 12 kbp against 3 Mbp with repeats (strands of very different hit counts, windows from a few to a
    few thousand hits): the shared-memory sorts and the CUB path give the same anchors and the same chained hits."""
    rng = np.random.default_rng(77 + overlap)
    G = 3_000_000
    genome = synth.random_seq(rng, G)
    for k in range(12):                                                # repeats: multi-hit buckets, crowded bins
        a, b = int(rng.integers(0, G - 5000)), int(rng.integers(0, G - 5000))
        genome[b:b + 4000] = genome[a:a + 4000]
    lens = rng.integers(200, 12000, 3000)
    stride = [int(L + ((-L) % 128)) for L in lens]
    ref_end = 128 + G + ((-G) % 128)
    arena = np.full(ref_end + sum(stride) + 128, ord("N"), np.uint8)
    arena[128:128 + G] = genome
    reads = np.zeros(len(lens), abi.SEED_READ)
    at = ref_end
    for k, L in enumerate(lens):
        p0 = int(rng.integers(0, G - L))
        r = synth.mutate_fast(rng, genome[p0:p0 + L], 0.03, 0.04, 0.04)[:L]
        if k & 1:
            r = synth.revcomp(r)
        arena[at:at + len(r)] = r
        reads[k]["read_addr"], reads[k]["read_len"] = at, len(r)
        at += stride[k]
    chroms = np.zeros(1, abi.CHROM); chroms["start"] = 128; chroms["len_unpadded"] = G
    out = {}
    for sort in ("smem", "cub"):
        monkeypatch.setenv("DARWIN_GPU_SEED_SORT", sort)
        p = gpu(len(arena), abi.Scoring.from_values())
        p.InitializeReferenceMemory(0, arena)
        p.build_seed_index(abi.SeedParams.stock(overlap), chroms, ref_end)
        out[sort] = p.seeder_body(reads)
        p.close()
    (b0, a0, p0), (b1, a1, p1) = out["smem"], out["cub"]
    assert len(a0) > 2000
    assert np.array_equal(b0, b1) and np.array_equal(a0, a1) and len(p0) == len(p1)
    # the pool region of a candidate is window + 1 entries long; only the two chains are defined
    used = np.zeros(len(p0) + 1, np.int64)
    for side in ("left", "right"):
        lo, n = a0[side + "_off"].astype(np.int64), a0[side + "_n"].astype(np.int64)
        np.add.at(used, lo, 1); np.add.at(used, lo + n, -1)
    used = np.cumsum(used)[:-1] > 0
    assert used.sum() > len(a0) and np.array_equal(p0[used], p1[used])

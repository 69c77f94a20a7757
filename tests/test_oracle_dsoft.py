"""CPU: the oracle's restatement of D-SOFT (oracle/dsoft_oracle.c: minimizers, seed position table, SeedPosTable::DSOFT ==
seed_pos_table.h:280-372, seed_pos_table.cpp:41-160, :252-553) against the compiled reference's own seeder_body."""
import numpy as np
import pytest

import oracle
from darwin_b200 import abi, synth
from conftest import GOLDEN


def same_seed_output(a, pa, b, pb):
    if len(a) != len(b) or not (np.array_equal(a["hit_offset"], b["hit_offset"]) and np.array_equal(a["left_n"], b["left_n"]) and
                                np.array_equal(a["right_n"], b["right_n"])):
        return False
    for x, y in zip(a, b):
        for off, n in (("left_off", "left_n"), ("right_off", "right_n")):
            if not np.array_equal(pa[int(x[off]):int(x[off]) + int(x[n])], pb[int(y[off]):int(y[off]) + int(y[n])]):
                return False
    return True


def strand_views(begin, anchors, n_reads):
    return [[anchors[begin[2 * r + s]:begin[2 * r + s + 1]] for s in (0, 1)] for r in range(n_reads)]


@pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,overlap", [(11, 0), (12, 1)])
def test_dsoft_port_matches_reference(seed, overlap):
    import sys
    sys.path.insert(0, GOLDEN)
    import make_golden
    ref, n_reads = make_golden.filter_case(seed, 40)
    ref.set_extend(384, 64, 2, overlap)                       # cfg.do_overlap: stop after N+1 seeds, SV window of one bin
    try:
        cands, rn = ref.seed(0, n_reads)
        begin, anchors, pool = ref.seed_anchors()
        arena = np.concatenate([ref.arena().copy(), np.full(256, ord("N"), np.uint8)])
        prm = ref.seed_params()
        assert prm.do_overlap == overlap
        dp = oracle.DsoftPort(arena, ref.chroms(), int(ref.lib.dref_arena_reference_size()), prm)
        views = strand_views(begin, anchors, n_reads)
        checked = 0
        for r in range(n_reads):
            L = ref.lib.dref_read_len(r)
            addr = ref.read_addr(r)
            fwd = arena[addr:addr + L]
            for strand in (0, 1):
                a, p = dp.query(np.ascontiguousarray(synth.revcomp(fwd) if strand else fwd))
                assert same_seed_output(a, p, views[r][strand], pool), (r, strand)
                checked += len(a)
        assert checked == len(anchors) and checked > 30
        dp.close()
    finally:
        ref.set_extend(384, 64, 2, 0)


def test_minimizer_rule():
    """iterate_minimizers_qw: window minimum of hash32 over w positions, emitted on change or every w positions."""
    prm = abi.SeedParams.stock()
    lib = oracle._load(oracle.os.path.join(oracle._HERE, "libgact_oracle.so"))
    lib.dsoft_minimizers.restype = oracle.C.c_uint64
    lib.dsoft_hash32.restype = oracle.C.c_uint32
    rng = np.random.default_rng(1)
    seq = np.concatenate([synth.random_seq(rng, 997), np.full(64, ord("N"), np.uint8)])
    out = np.zeros(2048, np.uint64)
    n = lib.dsoft_minimizers(abi.ptr(seq), oracle.C.c_uint32(997), prm.seed_size, prm.minimizer_window, abi.ptr(out))
    code = {65: 0, 67: 1, 71: 2, 84: 3, 78: 0}
    k, w = prm.seed_size, prm.minimizer_window
    hashes = []
    for p in range(1008 - k):
        seed = sum(code[int(seq[p + c])] << (2 * c) for c in range(k))
        hashes.append(lib.dsoft_hash32(oracle.C.c_uint32(seed), k))
    want, last_m, last_p = [], 0, 0
    for p in range(w - 1, 1008 - k):
        m = min(hashes[p - w + 1:p + 1])
        if m != last_m or p - last_p >= w:
            want.append((p << 32) | m)
            last_m, last_p = m, p
    assert list(out[:n]) == want and n > 300


def small_case(k, w, stride, n_reads=12, seed=None):
    """A reference driver loaded with a 200 kbp genome and a few reads, D-SOFT parameters (k, w, max_stride) set."""
    import ctypes as C
    rng = np.random.default_rng(seed if seed is not None else k)
    ref = oracle.reference("patched")
    ref.set_scoring(abi.Scoring.from_values())
    ref.lib.dref_set_dsoft(k, w, 64, 26, 300, 40, 1000, stride, 128, 60, 64, 1000, C.c_float(0.05))
    ref.set_extend(384, 64, 2, 0)
    ref.reset_arena()
    genome = synth.random_seq(rng, 200000)
    ref.add_chr("c", genome.tobytes(), True)
    ref.build_index()
    for r in range(n_reads):
        L = int(rng.integers(2000, 7000))
        p = int(rng.integers(0, len(genome) - L))
        s = synth.mutate_fast(rng, genome[p:p + L], 0.04, 0.04, 0.04)
        if r % 2:
            s = synth.revcomp(s)
        ref.add_read("r%d" % r, np.ascontiguousarray(s).tobytes())
    return ref, n_reads


@pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("k,w,stride", [(12, 5, 4), (13, 9, 2), (15, 4, 1)])
def test_dsoft_port_other_seed_shapes(k, w, stride):
    """w = 5 and w = 9 take the reference's AVX2 specialisations (seed_pos_table.h:374-522), other w the generic loop; all
    follow the same emission rule, which is what the port restates."""
    ref, n_reads = small_case(k, w, stride)
    try:
        ref.seed(0, n_reads)
        begin, anchors, pool = ref.seed_anchors()
        arena = np.concatenate([ref.arena().copy(), np.full(256, ord("N"), np.uint8)])
        dp = oracle.DsoftPort(arena, ref.chroms(), int(ref.lib.dref_arena_reference_size()), ref.seed_params())
        views = strand_views(begin, anchors, n_reads)
        for r in range(n_reads):
            L, addr = ref.lib.dref_read_len(r), ref.read_addr(r)
            fwd = arena[addr:addr + L]
            for s in (0, 1):
                a, p = dp.query(np.ascontiguousarray(synth.revcomp(fwd) if s else fwd))
                assert same_seed_output(a, p, views[r][s], pool), (k, w, r, s)
        assert len(anchors) >= n_reads
        dp.close()
    finally:
        ref.set_dsoft_defaults()

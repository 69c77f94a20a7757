"""GPU: the packed score-only first-tile path (gact_filter.cuh, two tiles per warp) and darwin_gpu_filter against the
oracle, the golden fixture of the compiled reference's filter_body, and -- through the C++ host adapter -- against the
reference's own filter_body on the same seeder output."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from darwin_b200 import abi, synth
from conftest import GOLDEN
from test_oracle_filter import locations_from, same_locations, check_custom

pytestmark = pytest.mark.gpu


def test_filter_body_matches_golden(gpu):
    g = np.load(os.path.join(GOLDEN, "filter_v1.npz"))
    sc = abi.Scoring.from_values(*[int(x) for x in g["scoring"]])
    p = gpu(len(g["arena"]), sc)
    p.InitializeReferenceMemory(0, g["arena"])
    st0 = p.stats()
    res = p.filter_body(g["cands"], 128, 60, 1000)
    st1 = p.stats()
    assert np.array_equal(res, g["port_res"])
    assert st1.tiles_filter - st0.tiles_filter == len(g["cands"])          # all of them on the packed path
    idx = locations_from(g["cands"], g["cand_read_num"], res, oracle.port(sc))
    assert same_locations(idx, g["cands"], g["cand_read_num"], res, g["anchors"])
    p.close()


def test_filter_body_custom_candidates(gpu):
    g = np.load(os.path.join(GOLDEN, "filter_v1.npz"))
    sc = abi.Scoring.from_values(*[int(x) for x in g["scoring"]])
    p = gpu(len(g["arena"]), sc)
    p.InitializeReferenceMemory(0, g["arena"])
    check_custom(g, lambda c, fts, thr, ovl: p.filter_body(c, fts, thr, ovl), oracle.port(sc))
    # thresholds: the flags follow filter.cpp:87 and :102-104
    res = p.filter_body(g["custom_cands"], 128, 60, 1000)
    want = oracle.port(sc).filter(g["arena"], g["custom_cands"], 128, 60, 1000)
    assert np.array_equal(res, want) and 0 < int((res["flags"] & 1).sum()) < len(res) and (res["flags"] & 2).any() and not (res["flags"] & 2).all()
    # a first_tile_size beyond the packed path's 128: every tile is handed over to the exact path inside the same call
    st0 = p.stats()
    res = p.filter_body(g["custom_cands"], 200, 60, 1000)
    st1 = p.stats()
    assert np.array_equal(res, oracle.port(sc).filter(g["arena"], g["custom_cands"], 200, 60, 1000))
    assert st1.tiles_exact - st0.tiles_exact > len(res) // 2
    p.close()


@pytest.mark.parametrize("n", [1, 2, 5001])
def test_filter_tiles_match_port(gpu, n):
    arena, req = synth.tile_batch_fast(17, n, 128, mode="filter")
    req["align_fields"][1::3] = abi.REVERSE_QUERY | abi.COMPLEMENT_QUERY    # the reverse-complement strand's flags
    sc = abi.Scoring.from_values()
    p = gpu(len(arena), sc)
    p.InitializeReferenceMemory(0, arena)
    st0 = p.stats()
    res, _ = p.BatchAlignmentSIMD(req, 0)
    st1 = p.stats()
    want, _, _ = oracle.port(sc).tiles(arena, req, 0, oracle.Port.STREAM, tb_words_per_req=1)
    assert np.array_equal(res, want)
    assert st1.tiles_filter - st0.tiles_filter == n and st1.tiles_exact == st0.tiles_exact
    p.close()


@pytest.mark.parametrize("vals", [(2, -6, -1, -4, -2, -25, -1), (1, -1, 0, -1, -1, -1, -1), (1, -1, 0, -2, -1, -4, 0),
                                  (3, -2, -1, -5, -1, -30, 0)])
def test_filter_mixed_shapes_flags_and_n(gpu, vals):
    """Ragged shapes (pairs with different geometry run alone), every flag set, N bases, tiles beyond the packed
    path's limits (exact-path hand-over), scoring schemes with saturated ties."""
    rng = np.random.default_rng(5 + vals[0])
    n = 700
    arena = [np.full(64, ord("N"), np.uint8)]
    pos, req = 64, np.zeros(n, abi.TILE_REQ)
    for k in range(n):
        R = int(rng.choice([128, 128, 128, 127, 96, 64, 33, 1, 160, 200])) if k % 4 else 128
        Q = int(rng.choice([128, 128, 128, 100, 31, 1, 129, 256])) if k % 4 else 128
        r = synth.random_seq(rng, R)
        q = synth.mutate(rng, r, 0.06, 0.04, 0.04, 0.01 if k % 50 == 7 else 0.0)
        q = np.concatenate([q, synth.random_seq(rng, Q)])[:Q]
        if k % 97 == 3:
            q = synth.random_seq(rng, Q)                                   # unrelated: score near zero
        for name, s in (("ref", r), ("query", q)):
            req[k][name + "_bases_start_addr"] = pos
            arena.append(s)
            pos += len(s)
        req[k]["ref_size"], req[k]["query_size"] = R, Q
        req[k]["max_tb_steps"] = 256
        req[k]["align_fields"] = int(rng.choice([0, 0, 0, 6, 6, 24, 30, 1, 4, 2, 16]))
        req[k]["index"] = k % 64
    arena = np.concatenate(arena + [np.full(64, ord("N"), np.uint8)])
    sc = abi.Scoring.from_values(*vals)
    p = gpu(len(arena), sc)
    p.InitializeReferenceMemory(0, arena)
    st0 = p.stats()
    res, _ = p.BatchAlignmentSIMD(req, 0)
    st1 = p.stats()
    want, _, _ = oracle.port(sc).tiles(arena, req, 0, oracle.Port.STREAM, tb_words_per_req=1)
    bad = [k for k in range(n) if res[k] != want[k]]
    assert bad == [], (bad[:5], res[bad[:5]], want[bad[:5]])
    packed, exact = st1.tiles_filter - st0.tiles_filter, st1.tiles_exact - st0.tiles_exact
    assert packed + exact == n and packed > n // 3 and exact > 0
    p.close()


def test_filter_rejects_inconsistent_candidates(gpu):
    arena, _ = synth.tile_batch_fast(1, 4, 128, mode="filter")
    p = gpu(len(arena), abi.Scoring.from_values())
    p.InitializeReferenceMemory(0, arena)
    c = np.zeros(1, abi.FILTER_CAND)
    c["chr_start"], c["chr_len"], c["hit"], c["read_addr"], c["read_len"], c["offset"] = 0, 512, 600, 512, 256, 0
    with pytest.raises(Exception):
        p.filter_body(c)                                                    # hit outside its chromosome
    assert len(p.filter_body(c[:0])) == 0
    p.close()


LIB = os.path.join(os.path.dirname(os.path.abspath(oracle.__file__)), "_ref", "libdarwin_ref_gpu.so")


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libdarwin_ref_gpu.so not built (needs /root/reference at build time)")
def test_gpu_filter_body_is_a_drop_in():
    """gpu_filter_body (darwin_b200/host) vs the reference's filter_body on the same seeder_body output: identical
    ExtendLocations incl. the chained hits; then the whole pipeline with both GPU stages vs the CPU pipeline."""
    ref = oracle.Reference.__new__(oracle.Reference)
    ref.lib = C.CDLL(LIB)
    L = ref.lib
    L.dref_arena.restype = C.c_void_p
    L.dref_arena_position.restype = C.c_uint64
    L.dref_add_chr.restype = C.c_uint64
    L.dref_anchor_hits_total.restype = C.c_uint64
    ref.set_scoring(abi.Scoring.from_values())
    ref.set_dsoft_defaults()
    ref.set_extend(384, 64, 2, 0)
    ref.reset_arena()
    rng = np.random.default_rng(77)
    genome = synth.random_seq(rng, 150000)
    rep = synth.mutate_fast(rng, genome[1000:4000], 0.03, 0.01, 0.01)[:2900]
    genome[90000:90000 + len(rep)] = rep
    ref.add_chr("chrS", genome.tobytes(), True)
    ref.build_index()
    nreads = 12
    for k in range(nreads):
        Lr = int(rng.integers(3000, 6000))
        s = int(rng.integers(0, len(genome) - Lr)) if k % 4 else len(genome) - Lr
        r = synth.mutate_fast(rng, genome[s:s + Lr], 0.05, 0.05, 0.05)
        if k % 2:
            r = synth.revcomp(r)
        ref.add_read("r%d" % k, np.ascontiguousarray(r).tobytes())
    cands, _ = ref.seed(0, nreads)
    a_cpu, h_cpu = ref.filter_last()
    assert L.dref_gpu_init(1) == 0
    try:
        a_gpu, h_gpu = ref.filter_last(gpu=True)
        assert len(cands) >= nreads and len(a_cpu) >= nreads
        assert np.array_equal(a_cpu, a_gpu) and np.array_equal(h_cpu, h_gpu)
        cap = 64 << 20
        buf_cpu, buf_gpu = C.create_string_buffer(cap), C.create_string_buffer(cap)
        n_gpu = L.dref_pipeline(0, nreads, 2, buf_gpu, C.c_uint64(cap))       # gpu_filter_body + gpu_extender_body
        L.dref_use_cpu_table()
        n_cpu = L.dref_pipeline(0, nreads, 0, buf_cpu, C.c_uint64(cap))
        assert n_cpu > 0 and n_gpu == n_cpu and buf_gpu.value == buf_cpu.value
    finally:
        L.dref_use_cpu_table()
        L.dref_gpu_shutdown()

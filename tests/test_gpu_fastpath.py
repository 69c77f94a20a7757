"""GPU: the packed fast path and its hand-over to the exact path (long-gap ties, band exits, N bases, ragged shapes,
truncated tracebacks), always bit-exact against the oracle's exact rule; plus size-independent properties on a
bench-sized batch."""
import numpy as np
import pytest

import oracle
from darwin_b200 import abi, synth
from conftest import tiles_equal

pytestmark = pytest.mark.gpu


def _ragged_batch(seed, n, tmax, long_indel=0, n_rate=0.0, small_tb=False):
    rng = np.random.default_rng(seed)
    arena, req, pos = [np.full(33, ord("N"), np.uint8)], np.zeros(n, abi.TILE_REQ), 33
    for k in range(n):
        R = tmax if k % 3 == 0 else int(rng.integers(1, tmax + 1))
        r = synth.random_seq(rng, R)
        q = synth.mutate(rng, r, 0.05, 0.05, 0.05, n_rate, (1, long_indel) if long_indel and k % 2 else None)
        Q = min(len(q), tmax) if k % 4 else min(len(q), int(rng.integers(1, tmax + 1)))
        q = q[:max(Q, 1)]
        req[k]["ref_bases_start_addr"], req[k]["ref_size"] = pos, R
        arena.append(r)
        pos += R
        req[k]["query_bases_start_addr"], req[k]["query_size"] = pos, len(q)
        arena.append(q)
        pos += len(q)
        req[k]["max_tb_steps"] = int(rng.choice([24, 100])) if (small_tb and k % 2) else 2 * tmax
        req[k]["align_fields"] = int(rng.choice([1, 21, 7, 19]))
        req[k]["index"] = k % 256
    return np.concatenate(arena + [np.full(64, ord("N"), np.uint8)]), req


def _run(gpu, sc, arena, req):
    p = gpu(len(arena), sc)
    p.InitializeReferenceMemory(0, arena)
    st0 = p.stats()
    res, tb = p.BatchAlignmentSIMD(req, 1)
    st1 = p.stats()
    pres, ptb, _ = oracle.port(sc).tiles(arena, req, 1, oracle.Port.STREAM, tb_words_per_req=tb.shape[1])
    assert tiles_equal(pres, ptb, res, tb) == []
    p.close()
    # (clean fast path, exact paths = unpacked + packed, fast tiles that had to be recomputed)
    return (st1.tiles_fast - st0.tiles_fast,
            (st1.tiles_exact - st0.tiles_exact) + (st1.tiles_xfast - st0.tiles_xfast), st1.tiles_rerun - st0.tiles_rerun)


@pytest.mark.parametrize("vals", [(2, -6, -1, -4, -2, -25, -1), (1, -1, 0, -1, -1, -1, -1), (1, -1, 0, -2, -1, -4, 0),
                                  (2, -3, -1, -3, -2, -8, -1), (5, -4, -1, -10, -1, -1000, -1)])
@pytest.mark.parametrize("tmax", [256, 320, 384])
def test_fast_path_all_schemes_ragged(gpu, vals, tmax):
    """K = 4/5/6 geometries; tie-heavy schemes force many exact reruns (long-gap candidates on the path)."""
    fast, exact, rerun = _run(gpu, abi.Scoring.from_values(*vals), *_ragged_batch(hash(vals) % 1000 + tmax, 600, tmax))
    assert fast + exact == 600
    if vals[5] == -1000:                    # bias would not fit 16 bits: exact path only
        assert fast == 0
    else:
        assert fast > 0
    if vals[5] == -1:                       # tie-saturated scoring: long gaps tie everywhere
        assert rerun > 0


def test_band_exit_falls_back_to_exact(gpu):
    """Indels longer than the stored band push the path out of shared memory -> exact recomputation."""
    fast, exact, rerun = _run(gpu, abi.Scoring.from_values(), *_ragged_batch(5, 500, 320, long_indel=90))
    assert rerun > 0 and fast > 0


def _tiles_with_n(arena, req):
    return sum(1 for r in req if (arena[int(r["ref_bases_start_addr"]):int(r["ref_bases_start_addr"]) + int(r["ref_size"])] == ord("N")).any()
               or (arena[int(r["query_bases_start_addr"]):int(r["query_bases_start_addr"]) + int(r["query_size"])] == ord("N")).any())


@pytest.mark.parametrize("vals", [(2, -6, -1, -4, -2, -25, -1), (2, -3, -3, -3, -2, -8, -1), (2, -6, 0, -4, -2, -25, -1)])
@pytest.mark.parametrize("tmax,n_rate", [(320, 0.004), (384, 0.02), (512, 0.004)])
def test_tiles_with_n_stay_on_packed_path(gpu, vals, tmax, n_rate):
    """N bases (Nt2Int -> 4, scored sub_N against anything, Processor.cpp:21-46, :50-74) no longer force the unpacked exact
    path: single N, N runs, N against N, on both sequences, single-strip and multi-strip geometries, sub_N = mismatch and 0."""
    arena, req = _ragged_batch(6 + tmax, 400, tmax, n_rate=n_rate)
    rng = np.random.default_rng(tmax)
    for k in range(0, len(req), 9):                            # N runs and N-vs-N columns on top of the sprinkled ones
        a, b = int(req[k]["ref_bases_start_addr"]), int(req[k]["query_bases_start_addr"])
        L = min(int(req[k]["ref_size"]), int(req[k]["query_size"]))
        if L > 40:
            p0 = int(rng.integers(0, L - 30))
            arena[a + p0:a + p0 + 12] = ord("N")
            arena[b + p0:b + p0 + 20] = ord("n") if k % 2 else ord("N")
    with_n = _tiles_with_n(arena, req)
    fast, exact, rerun = _run(gpu, abi.Scoring.from_values(*vals), arena, req)
    assert with_n > 100 and fast + exact == len(req)
    assert exact <= rerun                                      # only reruns (long-gap ties, band exits) leave the packed path
    assert fast >= len(req) - rerun


def test_tiles_with_n_unpackable_scoring_use_exact_path(gpu):
    """sub_N below the mismatch score cannot be expressed as a non-negative packed correction: such tiles keep the exact path."""
    arena, req = _ragged_batch(6, 300, 320, n_rate=0.004)
    fast, exact, rerun = _run(gpu, abi.Scoring.from_values(2, -6, -9, -4, -2, -25, -1), arena, req)
    assert exact >= _tiles_with_n(arena, req) > 0 and fast > 0


def test_truncated_traceback(gpu):
    """max_tb_steps smaller than the path (Processor.cpp:616) in both paths."""
    _run(gpu, abi.Scoring.from_values(), *_ragged_batch(7, 400, 320, small_tb=True))


def test_non_uniform_matrix_is_exact_only(gpu):
    m = dict(AA=3, AC=-5, AG=-2, AT=-5, CC=3, CG=-5, CT=-2, GG=3, GT=-5, TT=4)
    sc = abi.Scoring.from_values(matrix=m, sub_n=-1, go=-5, ge=-2, lgo=-20, lge=-1)
    fast, exact, rerun = _run(gpu, sc, *_ragged_batch(8, 300, 256))
    assert fast == 0 and exact == 300


def test_invalid_scoring_is_rejected(gpu):
    import darwin_b200
    p = darwin_b200.Processor(1 << 16)
    with pytest.raises(darwin_b200.DarwinGpuError) as e:
        p.InitializeScoringParameters(abi.Scoring.from_values(2, -6, -1, -1, -3, -25, -1))    # open cheaper than extend
    assert e.value.code == abi.ERR_INVALID
    p.close()


def test_bench_sized_batch_properties(gpu):
    """200k tiles of the bench workload: properties that need no oracle -- every op string is consistent with the
    reported offsets, ends where the reference's loop must end, and a 2 % sample is bit-exact against the oracle."""
    sc = abi.Scoring.from_values()
    n, T = 200000, 320
    arena, req = synth.tile_batch_fast(99, n, T)
    p = gpu(len(arena), sc)
    p.InitializeReferenceMemory(0, arena)
    res, tb = p.BatchAlignmentSIMD(req, 1)
    total = res["total_TB_pointers"].astype(np.int64)
    nw = tb.shape[1]
    shifts = (2 * np.arange(32, dtype=np.uint64))[None, None, :]
    cnt = np.zeros((n, 4), np.int64)
    for lo in range(0, n, 20000):
        hi = min(n, lo + 20000)
        ops = ((tb[lo:hi, :, None] >> shifts) & np.uint64(3)).reshape(hi - lo, nw * 32)
        valid = np.arange(nw * 32)[None, :] < total[lo:hi, None]
        for d in (1, 2, 3):
            cnt[lo:hi, d] = ((ops == d) & valid).sum(1)
        assert not ((ops == 0) & valid).any()                   # only I, D, M are ever emitted (Processor.cpp:570)
    assert np.array_equal(cnt[:, 3] + cnt[:, 1], res["query_offset"].astype(np.int64))     # M + I = query steps
    assert np.array_equal(cnt[:, 3] + cnt[:, 2], res["ref_offset"].astype(np.int64))       # M + D = reference steps
    assert (res["query_offset"] <= T).all() and (res["ref_offset"] <= T).all()
    assert (res["ref_max_pos"] == T - 1).all() and (res["query_max_pos"] == T - 1).all()   # corner start (Processor.cpp:544-547)
    assert (res["score"] >= 0).all() and np.median(res["score"]) > 100
    sample = np.arange(0, n, 50)
    pres, ptb, _ = oracle.port(sc).tiles(arena, req[sample], 1, oracle.Port.STREAM, tb_words_per_req=nw)
    assert tiles_equal(pres, ptb, res[sample], tb[sample]) == []
    p.close()


def test_sharded_extend_matches_single(gpu):
    """darwin_b200.shard with the GPU worker (world = 1 here; the world-2 plumbing is covered under gloo on CPU)."""
    from darwin_b200 import shard
    from test_host_logic import synthetic_anchor_set
    from conftest import alignments_equal, ALN_FIELDS_OURS
    sc = abi.Scoring.from_values()
    arena, anchors, hits = synthetic_anchor_set(31, 16, 2500)
    p = gpu(len(arena), sc)
    p.InitializeReferenceMemory(0, arena)
    whole = p.extender_body(anchors, hits, 320, 128, 0)
    parts_res, parts_ops, base = [], [], 0
    for rank in range(2):                                       # emulate two ranks on the one GPU, then concatenate
        a, hp, _ = shard.local_shard(anchors, hits, rank, 2)
        r, o = p.extender_body(a, hp, 320, 128, 0)
        used = int((r["ops_offset"] + r["n_ops"]).max())
        r = r.copy()
        r["ops_offset"] += base
        base += used
        parts_res.append(r)
        parts_ops.append(o[:used])
    assert alignments_equal(whole[0], whole[1], np.concatenate(parts_res), np.concatenate(parts_ops), ALN_FIELDS_OURS) == []
    p.close()


@pytest.mark.parametrize("vals", [(2, -6, -1, -4, -2, -25, -1), (3, -2, -1, -5, -1, -30, 0), (1, -1, 0, -1, -1, -1, -1)])
def test_wide_score_multistrip_T1024(gpu, vals):
    """T = 1024: match * 1024 + bias needs 12 score bits -> the multi-strip fast path runs in its 4-tag-bit layout.
    The batch mixes ragged tiles, low-error tiles and PERFECT 1024-base matches (score 2048 and 3072: the top score bit
    of the low half is set, nothing may leak into the neighbouring half's markers)."""
    sc = abi.Scoring.from_values(*vals)
    arena, req = _ragged_batch(31 + vals[0], 40, 1024)
    rng = np.random.default_rng(2)
    extra, pos = [], len(arena)
    req2 = np.zeros(6, abi.TILE_REQ)
    for k in range(6):
        r = synth.random_seq(rng, 1024)
        q = r.copy() if k < 3 else synth.mutate(rng, r, 0.01, 0.005, 0.005)[:1024]
        req2[k]["ref_bases_start_addr"], req2[k]["ref_size"] = pos, len(r)
        req2[k]["query_bases_start_addr"], req2[k]["query_size"] = pos + len(r), len(q)
        extra += [r, q]
        pos += len(r) + len(q)
        req2[k]["max_tb_steps"], req2[k]["align_fields"], req2[k]["index"] = 2048, (1 if k % 2 else 21), k
    arena = np.concatenate([arena] + extra + [np.full(64, ord("N"), np.uint8)])
    req = np.concatenate([req, req2])
    fast, exact, rerun = _run(gpu, sc, arena, req)
    assert fast + exact == len(req)
    if vals[0] * 1024 > 2047:
        assert fast >= 20                    # beyond the 11-bit layout, yet mostly on the fast path


def test_failed_create_leaves_no_handle_and_names_the_reason(gpu):
    import darwin_b200
    with pytest.raises(darwin_b200.DarwinGpuError) as e:
        darwin_b200.Processor(1 << 46)                       # 32 TiB of packed arena: cudaMalloc must refuse
    assert e.value.code == abi.ERR_CUDA and "cudaMalloc" in str(e.value)
    p = gpu(1 << 16, abi.Scoring.from_values())   # the device is still usable afterwards
    p.InitializeReferenceMemory(0, b"ACGT" * 64)
    req = np.zeros(1, abi.TILE_REQ)
    req["ref_size"] = req["query_size"] = 64
    req["query_bases_start_addr"] = 64
    req["max_tb_steps"] = 128
    res, _ = p.BatchAlignmentSIMD(req)
    assert int(res["total_TB_pointers"][0]) > 0


def test_oversized_tb_row_with_pageable_buffers_is_rejected(gpu):
    import darwin_b200
    p = gpu(1 << 16, abi.Scoring.from_values())
    p.InitializeReferenceMemory(0, b"ACGT" * 64)
    req = np.zeros(2, abi.TILE_REQ)
    req["ref_size"] = req["query_size"] = 64
    req["max_tb_steps"] = 128
    with pytest.raises(darwin_b200.DarwinGpuError) as e:
        p.BatchAlignmentSIMD(req, tb_words_per_req=(40 << 20) // 8)   # one TB row larger than the pinned staging buffer
    assert e.value.code == abi.ERR_INVALID


def _large_tiles(seed, n, related_every=3):
    """1984x960 / 960x1984 corner-traceback tiles as the extender requests them after a stall (extender.cpp:61-78): mostly
    UNRELATED sequences (spurious anchors: the corner is ZERO, no pointers), some true continuations, some with the corner
    sitting right after a long gap, some clipped at a sequence end."""
    rng = np.random.default_rng(seed)
    parts, pos = [np.full(64, ord("N"), np.uint8)], 64
    req = np.zeros(n, abi.TILE_REQ)
    for k in range(n):
        R, Q = (1984, 960) if k % 2 else (960, 1984)
        if k % 7 == 5:
            R -= int(rng.integers(1, 300))                       # clipped at a chromosome / read end
        r = synth.random_seq(rng, R)
        if k % related_every == 0:
            q = synth.mutate(rng, r, 0.05, 0.05, 0.05, indel_run=(2, 60))
            q = np.concatenate([q, synth.random_seq(rng, max(0, Q - len(q)))])[:Q]
        elif k % related_every == 1 and k % 4 == 1:
            # related, but the end of the query is a 30-base insertion: the corner sits inside / right after a long gap
            q = synth.mutate(rng, r[:Q - 30] if Q - 30 <= R else r, 0.03, 0.02, 0.02)
            q = np.concatenate([q, synth.random_seq(rng, max(0, Q - len(q)))])[:Q]
        else:
            q = synth.random_seq(rng, Q)
        fl = [1, 21, 7, 19][k % 4]
        req[k]["ref_bases_start_addr"], req[k]["ref_size"] = pos, len(r)
        pos += len(r)
        req[k]["query_bases_start_addr"], req[k]["query_size"] = pos, len(q)
        pos += len(q)
        parts += [r, q]
        req[k]["max_tb_steps"], req[k]["align_fields"], req[k]["index"] = 768, fl, k % 250
    return np.concatenate(parts + [np.full(64, ord("N"), np.uint8)]), req


@pytest.mark.parametrize("vals", [(2, -6, -1, -4, -2, -25, -1), (2, -3, -1, -3, -2, -8, -1), (1, -1, 0, -2, -1, -4, 0)])
def test_large_tiles_score_only_prepass(gpu, vals):
    """The score-only pre-pass (gact_score.cuh) settles large tiles whose corner is provably ZERO and hands every other one
    to the traced paths; results are the reference's either way, and the pre-pass really is what ran for the zero ones."""
    sc = abi.Scoring.from_values(*vals)
    arena, req = _large_tiles(900 + vals[1], 48)
    p = gpu(len(arena), sc)
    p.InitializeReferenceMemory(0, arena)
    st0 = p.stats()
    res, tb = p.BatchAlignmentSIMD(req, 1, tb_words_per_req=50)
    st1 = p.stats()
    pres, ptb, _ = oracle.port(sc).tiles(arena, req, 1, oracle.Port.STREAM, tb_words_per_req=50)
    assert tiles_equal(pres, ptb, res, tb) == []
    zero = int((pres["total_TB_pointers"] == 0).sum())
    settled = st1.tiles_scoreonly - st0.tiles_scoreonly
    assert 0 <= settled <= zero                                  # never claims a tile that has pointers ...
    assert settled >= zero - 3                                   # ... and misses at most the rare "long chain exactly 0" corners
    if vals[0] == 2:
        assert zero > 10                                         # (+1/-1 scoring never reaches a zero corner on random sequences)
    assert int((pres["total_TB_pointers"] > 0).sum()) > 10
    p.close()


def test_extend_slots_reports_whole_waves(gpu):
    """darwin_gpu_extend_slots: anchors one launch keeps in flight -- a multiple of the SM count for every tile size, refused
    before the scoring is set (the geometry depends on it) and for tile sizes beyond the largest tile."""
    import torch
    import darwin_b200
    from darwin_b200.gact import DarwinGpuError
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    raw = darwin_b200.Processor(1 << 20)
    with pytest.raises(DarwinGpuError):
        raw.extend_slots(384)
    raw.close()
    p = gpu(1 << 20, abi.Scoring.from_values())
    for T in (128, 256, 320, 384, 512, 1024):
        s = p.extend_slots(T)
        assert s >= sms and s % sms == 0
    assert p.extend_slots(384) >= p.extend_slots(512)          # K = 8 keeps two band words per lane-step: fewer resident warps
    with pytest.raises(DarwinGpuError):
        p.extend_slots(4096)
    p.close()


def test_mixed_shapes_keep_their_own_geometry(gpu):
    """Geometry is chosen per tile: a batch of 320 x 320 tiles with a handful of 400-, 512- and 700-wide ones in between gives
    the oracle's answers, and the 320 x 320 tiles do not fall off a cliff (before: one wide tile put the whole call on the
    two-strip variant, ~1.7x slower)."""
    sc = abi.Scoring.from_values()
    n = 60000
    arena, req = synth.tile_batch_fast(91, n, 320)
    p = gpu(len(arena) + (1 << 16), sc)
    p.InitializeReferenceMemory(0, arena)
    p.BatchAlignmentSIMD(req[:2000], 1, 46)
    p.BatchAlignmentSIMD(req, 1, 46)
    t_uniform = p.stats().last_kernel_ms
    # wide tiles cut from the same arena (reference of tile k followed by its query and the next tiles: related prefixes)
    mixed = req.copy()
    for k, (R, Q) in zip(range(100, n, 9973), [(400, 400), (512, 512), (700, 384), (400, 320), (320, 450), (512, 300)]):
        mixed[k]["ref_size"], mixed[k]["query_size"] = R, Q
        mixed[k]["max_tb_steps"] = 1400
    res, tb = p.BatchAlignmentSIMD(mixed, 1, 90)
    t_mixed = p.stats().last_kernel_ms
    idx = np.concatenate([np.arange(0, n, 37), np.arange(100, n, 9973)])
    pres, ptb, _ = oracle.port(sc).tiles(arena, mixed[idx], 1, oracle.Port.STREAM, tb_words_per_req=90)
    assert tiles_equal(pres, ptb, res[idx], tb[idx]) == []
    assert t_mixed < 1.25 * t_uniform, (t_mixed, t_uniform)
    p.close()

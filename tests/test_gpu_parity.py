"""GPU parity tests: the CUDA path (through the C-ABI, via darwin_b200.Processor) against the golden
vectors of the compiled reference and against the oracle on fresh seeded inputs.  Bit-exact: integer scores,
coordinates and op strings -- no tolerance anywhere."""
import numpy as np
import pytest

import oracle
from darwin_b200 import abi, synth
from conftest import tiles_equal, alignments_equal, ALN_FIELDS, ALN_FIELDS_OURS

pytestmark = pytest.mark.gpu
SCHEMES = ("stock", "tie", "s2", "s3")


@pytest.mark.parametrize("scheme", SCHEMES)
def test_tiles_golden(gpu, golden_tiles, scheme):
    g = golden_tiles
    arena = g[scheme + "_arena"]
    p = gpu(len(arena), abi.Scoring.from_values(*g[scheme + "_scoring"].tolist()))
    p.InitializeReferenceMemory(0, arena)
    res, tb = p.BatchAlignmentSIMD(g[scheme + "_req"], 1, tb_words_per_req=260)
    assert tiles_equal(g[scheme + "_res"], g[scheme + "_tb"], res, tb) == []
    res0, _ = p.BatchAlignmentSIMD(g[scheme + "_req"], 0)
    assert np.array_equal(res0, g[scheme + "_res_notb"])
    p.close()


@pytest.mark.parametrize("tag", ["T384_O64_ovl0", "T320_O128_ovl0", "T256_O64_ovl1"])
def test_extend_golden(gpu, golden_extend, tag):
    g = golden_extend
    arena = g["arena"]
    p = gpu(len(arena), abi.Scoring.from_values(*g["scoring"].tolist()))
    p.InitializeReferenceMemory(0, arena)
    T, O, ovl = [int(x.lstrip("TOovl")) for x in tag.split("_")]
    res, ops = p.extender_body(g[tag + "_anchors"], g[tag + "_hits"], T, O, ovl)
    assert alignments_equal(g[tag + "_res"], g[tag + "_ops"], res, ops, ALN_FIELDS) == []
    p.close()


@pytest.mark.parametrize("tile,overlap", [(320, 128), (384, 64), (100, 20)])
def test_tiles_config2_shape_vs_oracle(gpu, tile, overlap):
    """BASELINE.json configs[1] shape (independent T x T tiles, 15 % error, left and right extension flags)."""
    sc = abi.Scoring.from_values()
    arena, req = synth.tile_batch_fast(11 + tile, 3000, tile)
    p = gpu(len(arena), sc)
    p.InitializeReferenceMemory(0, arena)
    res, tb = p.BatchAlignmentSIMD(req, 1)
    pres, ptb, _ = oracle.port(sc).tiles(arena, req, 1, oracle.Port.STREAM, tb_words_per_req=tb.shape[1])
    assert tiles_equal(pres, ptb, res, tb) == []
    # consumption rule (T - O steps + word quirk) on both sides gives the same op prefix
    S = tile - overlap
    for k in range(0, len(req), 97):
        assert synth.consumed_ops(tb[k], int(res[k]["total_TB_pointers"]), S) == \
            synth.consumed_ops(ptb[k], int(pres[k]["total_TB_pointers"]), S)
    p.close()


def test_tiles_edge_cases(gpu):
    """Empty tiles, 1x1, ragged shapes, N runs, lower case, max_tb_steps truncation, unaligned arena addresses."""
    sc = abi.Scoring.from_values()
    rng = np.random.default_rng(3)
    seq = synth.random_seq(rng, 5000)
    seq[100:140] = ord("N")
    seq[1000:1200] = np.char.lower(seq[1000:1200].view("S1")).view(np.uint8)
    arena = np.concatenate([np.full(37, ord("N"), np.uint8), seq, np.full(64, ord("N"), np.uint8)])
    shapes = [(0, 10), (10, 0), (1, 1), (1, 300), (300, 1), (17, 333), (333, 17), (255, 257), (256, 256), (257, 255),
              (512, 512), (1, 2), (64, 640), (31, 33), (400, 399)]
    req = np.zeros(len(shapes) * 4, abi.TILE_REQ)
    k = 0
    for (R, Q) in shapes:
        for fl in (1, 21, 0, 7):
            req[k]["ref_bases_start_addr"] = 37 + (k * 13) % 900
            req[k]["query_bases_start_addr"] = 37 + 60 + (k * 29) % 1500
            req[k]["ref_size"], req[k]["query_size"] = R, Q
            req[k]["max_tb_steps"] = 40 if k % 5 == 0 else 1400
            req[k]["align_fields"] = fl
            req[k]["index"] = k
            k += 1
    p = gpu(len(arena), sc)
    p.InitializeReferenceMemory(0, arena)
    res, tb = p.BatchAlignmentSIMD(req, 1, tb_words_per_req=100)
    pres, ptb, _ = oracle.port(sc).tiles(arena, req, 1, oracle.Port.STREAM, tb_words_per_req=100)
    assert tiles_equal(pres, ptb, res, tb) == []
    p.close()


def test_upload_chunks_and_odd_addresses(gpu):
    """InitializeReferenceMemory at arbitrary (odd) arena offsets and in pieces gives the same packed arena."""
    sc = abi.Scoring.from_values()
    arena, req = synth.tile_batch_fast(5, 64, 96)
    p1 = gpu(len(arena), sc)
    p1.InitializeReferenceMemory(0, arena)
    p2 = gpu(len(arena), sc)
    cuts = [0, 1, 2, 7, 333, 334, 4097, len(arena)]
    for a, b in reversed(list(zip(cuts[:-1], cuts[1:]))):
        p2.InitializeReadMemory(a, arena[a:b])
    r1, t1 = p1.BatchAlignmentSIMD(req, 1)
    r2, t2 = p2.BatchAlignmentSIMD(req, 1)
    assert tiles_equal(r1, t1, r2, t2) == []      # TB words beyond total_TB_pointers are unspecified
    p1.close()
    p2.close()


def test_extend_random_vs_oracle(gpu):
    """Synthetic anchors (true positions + regular chained hits) on a random reference, both strands,
    including reads that run off the reference ends (clipped tiles over 'N' padding)."""
    from test_host_logic import synthetic_anchor_set
    sc = abi.Scoring.from_values()
    arena, anchors, hits = synthetic_anchor_set(seed=21, n_reads=24, read_len=3000)
    p = gpu(len(arena), sc)
    p.InitializeReferenceMemory(0, arena)
    for (T, O) in ((384, 64), (320, 128), (128, 32), (512, 64), (1024, 64)):     # 512: multi-strip fast; 1024: exact path
        res, ops = p.extender_body(anchors, hits, T, O, 0)
        pres, pops = oracle.port(sc).extend(arena, abi.ExtendParams(T, O, 0, 0), anchors, hits, oracle.Port.STREAM)
        assert alignments_equal(pres, pops, res, ops, ALN_FIELDS_OURS) == []
        assert (res["flags"] & 1).sum() > 0
    p.close()


@pytest.mark.parametrize("err", [(0.015, 0.09, 0.045), (0.015, 0.045, 0.09)], ids=["insertion_rich", "deletion_rich"])
def test_extend_drifting_reads_vs_oracle(gpu, err):
    """Reads whose paths drift off the tile diagonal (PacBio-like 9 % insertions, and the mirror image): the anchor walker centres
    the band of the next tile on the drift of the previous ones (TileJob::band_shift) -- a scheduling hint that must never show
    in the result, and that keeps these tiles on the packed path."""
    from test_host_logic import synthetic_anchor_set
    sc = abi.Scoring.from_values()
    arena, anchors, hits = synthetic_anchor_set(seed=33, n_reads=32, read_len=6000, ref_len=120000, err=err)
    p = gpu(len(arena), sc)
    p.InitializeReferenceMemory(0, arena)
    for (T, O) in ((384, 64), (320, 128), (256, 64)):
        st0 = p.stats()
        res, ops = p.extender_body(anchors, hits, T, O, 0)
        st = p.stats()
        pres, pops = oracle.port(sc).extend(arena, abi.ExtendParams(T, O, 0, 0), anchors, hits, oracle.Port.STREAM)
        assert alignments_equal(pres, pops, res, ops, ALN_FIELDS_OURS) == []
        assert (res["flags"] & 1).sum() > 0
        tiles = int(res["n_tiles"].sum())
        assert (st.tiles_rerun - st0.tiles_rerun) < 0.25 * tiles                 # (half of the reads carry indel runs of up to 50 bases)
    p.close()


def test_upload_spans_equals_single_upload(gpu):
    """darwin_gpu_upload_spans (many spans, shared staging buffers, odd boundaries, one span larger than a staging buffer)
    leaves the arena exactly as one darwin_gpu_upload of the whole range does."""
    sc = abi.Scoring.from_values()
    arena, req = synth.tile_batch_fast(23, 3000, 320)
    big = np.concatenate([arena, synth.random_seq(np.random.default_rng(1), 40 << 20)])      # + a 40 MB tail (> 32 MB staging)
    p1 = gpu(len(big), sc)
    p1.InitializeReferenceMemory(0, big)
    p2 = gpu(len(big), sc)
    rng = np.random.default_rng(2)
    cuts = sorted(set([0, len(arena)] + [int(x) for x in rng.integers(1, len(arena), 200)]))
    spans = [(a, big[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    rng.shuffle(spans)
    spans.append((len(arena), big[len(arena):]))                                             # the oversized span
    p2.upload_spans(spans)
    r1, t1 = p1.BatchAlignmentSIMD(req, 1)
    r2, t2 = p2.BatchAlignmentSIMD(req, 1)
    assert tiles_equal(r1, t1, r2, t2) == []
    tail = np.zeros(64, abi.TILE_REQ)
    tail["ref_bases_start_addr"] = len(arena) + np.arange(64) * 600000 + 1
    tail["query_bases_start_addr"] = tail["ref_bases_start_addr"] + 7
    tail["ref_size"], tail["query_size"], tail["max_tb_steps"], tail["align_fields"] = 300, 300, 600, 1
    a1, b1 = p1.BatchAlignmentSIMD(tail, 1)
    a2, b2 = p2.BatchAlignmentSIMD(tail, 1)
    assert tiles_equal(a1, b1, a2, b2) == []
    p1.close()
    p2.close()


def test_upload_mixed_case_and_non_acgt_at_every_alignment(gpu):
    """Nt2Int semantics of the packing kernel (Processor.cpp:21-46): lower case counts, everything else is N -- on the
    vector path, on the byte-wise edges and for spans shorter than one 16-base group, at all 16 address alignments."""
    sc = abi.Scoring.from_values()
    arena, req = synth.tile_batch_fast(31, 400, 96)
    rng = np.random.default_rng(32)
    lower = rng.random(len(arena)) < 0.3
    arena[lower] = np.frombuffer(arena[lower].tobytes().lower(), np.uint8)
    odd = rng.random(len(arena)) < 0.03
    arena[odd] = rng.choice(np.frombuffer(b"NnRYKMSWxX-*.\x00\xff@[`{", np.uint8), int(odd.sum()))
    want = oracle.port(sc).tiles(arena, req, 1, oracle.Port.STREAM, tb_words_per_req=14)[:2]
    p1 = gpu(len(arena) + 64, sc)
    p1.InitializeReferenceMemory(0, arena)
    r1, t1 = p1.BatchAlignmentSIMD(req, 1, 14)
    assert tiles_equal(want[0], want[1], r1, t1) == []
    # the same bytes shifted to every alignment, uploaded as ragged spans (1..40 bases) and through the chunked call
    for shift in range(1, 17):
        p2 = gpu(len(arena) + 64, sc)
        cuts = [0]
        while cuts[-1] < len(arena):
            cuts.append(min(len(arena), cuts[-1] + int(rng.integers(1, 41 if shift % 2 else 5000))))
        spans = [(shift + a, arena[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
        if shift % 4 == 0:
            p2.InitializeReadMemory(shift, arena)
        else:
            p2.upload_spans(spans)
        rq = req.copy()
        rq["ref_bases_start_addr"] += shift
        rq["query_bases_start_addr"] += shift
        r2, t2 = p2.BatchAlignmentSIMD(rq, 1, 14)
        assert tiles_equal(want[0], want[1], r2, t2) == [], shift
        p2.close()
    p1.close()

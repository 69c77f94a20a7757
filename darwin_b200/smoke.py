"""__graft_entry__.smoke(): one small invocation of the hot path on cuda:0, checked against the oracle
(the only place in this package that touches oracle/ -- as the checker, never as the thing shipped)."""
import numpy as np


def run():
    import oracle
    from . import abi, synth, Processor
    sc = abi.Scoring.from_values()
    arena, req = synth.tile_batch_fast(1, 512, 320)
    p = Processor(len(arena), 0)
    p.InitializeScoringParameters(sc)
    p.InitializeReferenceMemory(0, arena)
    res, tb = p.BatchAlignmentSIMD(req, 1)
    pres, ptb, _ = oracle.port(sc).tiles(arena, req, 1, oracle.Port.STREAM, tb_words_per_req=tb.shape[1])
    chk = res.copy()
    chk["status"] &= 0x0F                      # bit 4 (DARWIN_TILE_LONG_INS_PATH) is information the reference does not return
    assert np.array_equal(chk, pres), "tile results differ from the oracle"
    for k in range(len(req)):
        nw = (int(res[k]["total_TB_pointers"]) + 31) // 32
        assert np.array_equal(tb[k, :nw], ptb[k, :nw]), "TB words differ from the oracle (tile %d)" % k
    ms_tiles = p.stats().last_kernel_ms
    p.close()
    # first-tile filter (packed score-only kernel) and whole anchors (in-kernel tile walking) on a small read set
    arena, anchors, hits = synth.anchor_batch(2, 16, 2500, 100000)
    p = Processor(len(arena), 0)
    p.InitializeScoringParameters(sc)
    p.InitializeReferenceMemory(0, arena)
    cands = np.zeros(len(anchors), abi.FILTER_CAND)
    for f in ("read_addr", "read_len", "chr_start", "strand"):
        cands[f] = anchors[f]
    cands["chr_len"] = anchors["ref_len"]
    cands["hit"] = np.maximum(anchors["reference_pos"].astype(np.int64) - 50, anchors["chr_start"])
    cands["offset"] = np.maximum(anchors["query_pos"].astype(np.int64) - 50, 0)
    fres = p.filter_body(cands)
    assert np.array_equal(fres, oracle.port(sc).filter(arena, cands)), "first-tile filter results differ from the oracle"
    res, ops = p.extender_body(anchors, hits, 384, 64, 0)
    pres, pops = oracle.port(sc).extend(arena, abi.ExtendParams(384, 64, 0, 0), anchors, hits, oracle.Port.STREAM)
    for k in range(len(anchors)):
        assert all(res[k][f] == pres[k][f] for f in ("n_ops", "reference_start_offset", "reference_end_offset",
                                                      "query_start_offset", "query_end_offset", "score", "n_tiles")), "alignment %d differs" % k
        assert np.array_equal(ops[int(res[k]["ops_offset"]):int(res[k]["ops_offset"]) + int(res[k]["n_ops"])],
                              pops[int(pres[k]["ops_offset"]):int(pres[k]["ops_offset"]) + int(pres[k]["n_ops"])]), "ops %d differ" % k
    st = p.stats()
    assert st.kernel_launches > 0 and st.tiles_filter == len(cands)
    print("smoke OK: %d tiles, %d first tiles, %d anchors (%d tiles) bit-exact; tiles kernel %.3f ms" % (
        len(req), len(cands), len(anchors), int(res["n_tiles"].sum()), ms_tiles))
    p.close()

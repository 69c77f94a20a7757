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
    assert np.array_equal(res, pres), "tile results differ from the oracle"
    for k in range(len(req)):
        nw = (int(res[k]["total_TB_pointers"]) + 31) // 32
        assert np.array_equal(tb[k, :nw], ptb[k, :nw]), "TB words differ from the oracle (tile %d)" % k
    st = p.stats()
    assert st.kernel_launches > 0
    print("smoke OK: %d tiles bit-exact, %.3f ms kernel, %d launches" % (len(req), st.last_kernel_ms, st.kernel_launches))
    p.close()

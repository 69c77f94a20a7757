"""First-tile filter shape (filter.cpp:61-71): 128x128 score-only tiles in max-cell mode through darwin_gpu_tiles."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import darwin_b200, oracle
from darwin_b200 import abi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
arena, req = synth.tile_batch_fast(3, n, 128, mode="filter")
p = darwin_b200.Processor(len(arena)); p.InitializeScoringParameters(abi.Scoring.from_values()); p.InitializeReferenceMemory(0, arena)
p.BatchAlignmentSIMD(req[:1000], 0)
t0 = time.time(); res, _ = p.BatchAlignmentSIMD(req, 0); wall = time.time() - t0
ms = p.stats().last_kernel_ms
print("filter tiles: %d x 128x128 score-only: kernel %.2f ms -> %.1f GCUPS (%.2f M tiles/s), wall %.1f ms" % (n, ms, n * 16384 / ms / 1e6, n / ms / 1e3, wall * 1e3))
pres, _, _ = oracle.port(abi.Scoring.from_values()).tiles(arena, req[:2000], 0, oracle.Port.STREAM, tb_words_per_req=1)
print("parity (2000 tiles):", np.array_equal(pres, res[:2000]), "pass rate (score>=60): %.2f" % (res["score"] >= 60).mean())
if oracle.have_reference():
    ref = oracle.reference("as-is"); ref.set_scoring(abi.Scoring.from_values())
    m = min(n, 40000 * (os.cpu_count() or 1) // 16)
    _, _, secs = ref.tiles(arena, req[:m], 0, tb_words_per_req=1, threads=os.cpu_count())
    print("reference BatchAlignmentSIMD score-only x %d threads: %.2f GCUPS" % (os.cpu_count(), m * 16384 / secs / 1e9))

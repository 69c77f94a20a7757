"""Analysis script for SURVEY 8(d) config 5: ONT-like 50 kbp reads at 12 % error (sub 4 / ins 3 / del 5), tile_size sweep
T = 256 / 384 / 512 / 1024 (tile_overlap 64) through darwin_gpu_extend, plus do_overlap = 1 at T = 256.
Usage: python scripts/tile_sweep.py [n_reads] [read_len]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import darwin_b200
from darwin_b200 import abi, synth
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
read_len = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
arena, anchors, hits = synth.anchor_batch(11, n_reads, read_len, 20_000_000, err=(0.04, 0.03, 0.05))
p = darwin_b200.Processor(len(arena))
p.InitializeScoringParameters(abi.Scoring.from_values())
p.InitializeReferenceMemory(0, arena)
for T, O, ovl in ((256, 64, 0), (384, 64, 0), (512, 64, 0), (1024, 64, 0), (256, 64, 1)):
    p.extender_body(anchors[:32], hits, T, O, ovl)
    st0 = p.stats()
    t0 = time.time()
    res, ops = p.extender_body(anchors, hits, T, O, ovl)
    wall = time.time() - t0
    st = p.stats()
    cells = float(res["cells"].sum())
    print("T=%4d O=%d overlap=%d: %d reads x %d bp: kernel %.1f ms, %.3g cells -> %.0f GCUPS, %.0f reads/s (kernel), %.0f reads/s (wall) | tiles %d: fast %d xfast %d exact %d rerun %d | aligned %d, mean aligned len %.0f" % (
        T, O, ovl, n_reads, read_len, st.last_kernel_ms, cells, cells / st.last_kernel_ms / 1e6, n_reads / st.last_kernel_ms * 1e3, n_reads / wall,
        int(res["n_tiles"].sum()), st.tiles_fast - st0.tiles_fast, st.tiles_xfast - st0.tiles_xfast, st.tiles_exact - st0.tiles_exact,
        st.tiles_rerun - st0.tiles_rerun, int((res["flags"] & 1).sum()),
        float((res["query_end_offset"].astype(np.int64) - res["query_start_offset"])[(res["flags"] & 1) == 1].mean())))

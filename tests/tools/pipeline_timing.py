"""Analysis script (not part of bench.py): realistic anchors from the reference's D-SOFT + filter (oracle/_ref), extended
by (a) the reference's extender_body on all host cores and (b) darwin_gpu_extend.  Prints the tile census and both
throughputs.  Usage: python tests/tools/pipeline_timing.py [n_reads]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import oracle  # noqa: E402
import darwin_b200  # noqa: E402
from darwin_b200 import abi  # noqa: E402
from test_gpu_pipeline_scale import build_case  # noqa: E402
from conftest import alignments_equal, ALN_FIELDS  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 400
t0 = time.time()
ref, anchors, hits = build_case(5, n_reads, genome_len=2000000)
print("reads %d anchors %d hits %d (D-SOFT+filter %.1f s on 1 core)" % (n_reads, len(anchors), len(hits), time.time() - t0))
ref.set_extend(384, 64, 2, 0)
cores = os.cpu_count()
secs, cells, alns = ref.extend_mt(anchors, hits, cores)
print("reference extender_body x %d threads: %.2f s, %.3g cells, %d alignments -> %.2f GCUPS, %.0f reads/s" % (
    cores, secs, cells, alns, cells / secs / 1e9, n_reads / secs))
arena = ref.arena().copy()
p = darwin_b200.Processor(len(arena))
p.InitializeScoringParameters(abi.Scoring.from_values())
p.InitializeReferenceMemory(0, arena)
p.extender_body(anchors[:8], hits, 384, 64, 0)
st0 = p.stats()
t0 = time.time()
res, ops = p.extender_body(anchors, hits, 384, 64, 0)
wall = time.time() - t0
st = p.stats()
gc = float(res["cells"].sum())
large_cells = 0
print("GPU darwin_gpu_extend: kernel %.2f ms (wall %.1f ms), %.3g cells, %d alignments -> %.1f GCUPS, %.0f reads/s (kernel)" % (
    st.last_kernel_ms, wall * 1e3, gc, int((res["flags"] & 1).sum()), gc / st.last_kernel_ms / 1e6, n_reads / st.last_kernel_ms * 1e3))
print("tiles %d (large %d), fast %d exact %d rerun %d" % (int(res["n_tiles"].sum()), int(res["n_large_tiles"].sum()),
      st.tiles_fast - st0.tiles_fast, st.tiles_exact - st0.tiles_exact, st.tiles_rerun - st0.tiles_rerun), "xfast", st.tiles_xfast - st0.tiles_xfast)
print("cells on the exact path: %.3g of %.3g (%.1f %%)" % (st.cells_exact - st0.cells_exact, gc, 100.0 * (st.cells_exact - st0.cells_exact) / gc))
want_res, want_ops = ref.extend(anchors[:100], hits)
print("parity on first 100 anchors:", alignments_equal(want_res, want_ops, res[:100], ops, ALN_FIELDS) == [])

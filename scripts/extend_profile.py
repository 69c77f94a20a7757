"""Where the extension time goes on realistic anchors (D-SOFT + filter locations, spurious ones and large tiles included):
one resident darwin_gpu_align_reads call per configuration, single lane; prints the per-stage kernel times, the tile-path
counters (fast / score-only / packed exact / unpacked exact / reruns) and the extension rate.
Usage: python scripts/extend_profile.py [config3|config4|config5] [n_reads] [tile_size]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import darwin_b200
from darwin_b200 import abi, workloads

which = sys.argv[1] if len(sys.argv) > 1 else "config3"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
T = int(sys.argv[3]) if len(sys.argv) > 3 else 384
cfgs = {"config3": (10_000_000, 10000, (0.015, 0.09, 0.045), 31), "config4": (250_000_000, 10000, (0.05, 0.05, 0.05), 41),
        "config5": (20_000_000, 50000, (0.04, 0.03, 0.05), 51)}
G, L, err, seed = cfgs[which]
dev = torch.device("cuda", 0)
case = workloads.ReadSetCase(dev, G, n_reads, 0, n_reads, L, err, seed)
p = darwin_b200.Processor(case.arena_bytes, 0)
p.InitializeScoringParameters(abi.Scoring.from_values())
p.InitializeReferenceMemory(0, case.ref_numpy())
p.InitializeReadMemory(case.read_addr(0), case.reads_numpy(0, case.n))
p.build_seed_index(abi.SeedParams.stock(), case.chroms, case.ref_end)
prm = abi.AlignParams.stock(T, 64, 0)
p.align_reads(case.seed_reads[:min(2000, n_reads)], prm)
for rep in range(2):
    st0 = p.stats()
    t0 = time.time()
    an, res, ops = p.align_reads(case.seed_reads, prm)
    wall = time.time() - t0
    st = p.stats()
    cells = float(res["cells"].sum())
    em = (res["flags"] & 1) != 0
    large_cells = 0.0
    print("%s T=%d: %d reads -> %d locations, %d alignments, %d tiles (%d large) | seed %.1f ms, filter %.1f ms, extend %.1f ms, wall %.1f ms"
          % (which, T, n_reads, len(res), int(em.sum()), int(res["n_tiles"].sum()), int(res["n_large_tiles"].sum()),
             st.last_seed_ms, st.last_filter_ms, st.last_extend_ms, wall * 1e3))
    print("   extension: %.3g cells -> %.0f GCUPS (kernel), %.0f reads/s (kernel), %.0f reads/s (call) | tiles fast %d, score-only %d, "
          "packed-exact %d, unpacked-exact %d, reruns %d, exact cells %.3g" % (
              cells, cells / st.last_extend_ms / 1e6, n_reads / (st.last_kernel_ms * 1e-3), n_reads / wall,
              st.tiles_fast - st0.tiles_fast, st.tiles_scoreonly - st0.tiles_scoreonly, st.tiles_xfast - st0.tiles_xfast,
              st.tiles_exact - st0.tiles_exact, st.tiles_rerun - st0.tiles_rerun, float(st.cells_exact - st0.cells_exact)))
    nt = res["n_tiles"].astype(np.int64)
    print("   per location: tiles p50 %d p90 %d max %d; locations with <= 4 tiles: %d; long-ins flagged %d, exact-rerun flagged %d" % (
        np.percentile(nt, 50), np.percentile(nt, 90), nt.max(), int((nt <= 4).sum()), int(((res["flags"] & 8) != 0).sum()), int(((res["flags"] & 4) != 0).sum())))

"""D-SOFT phase profile: resident reads of a ReadSetCase through seeder_body (one call), meant to be run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel launch list.
Usage: python scripts/seed_phases.py [genome_bp] [n_reads] [read_len] [sub ins del]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import darwin_b200
from darwin_b200 import abi, workloads
G = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000_000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 12000
read_len = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
err = tuple(float(x) for x in sys.argv[4:7]) if len(sys.argv) > 6 else (0.05, 0.05, 0.05)
c = workloads.ReadSetCase(torch.device("cuda", 0), G, n_reads, 0, n_reads, read_len, err, 11)
p = darwin_b200.Processor(c.arena_bytes)
p.InitializeScoringParameters(abi.Scoring.from_values())
p.InitializeReferenceMemory(0, c.ref_numpy())
p.InitializeReadMemory(int(c.seed_reads["read_addr"][0]), c.reads_numpy(0, n_reads))
p.build_seed_index(abi.SeedParams.stock(), c.chroms, c.ref_end)
p.seeder_body(c.seed_reads[:256])
for rep in range(2):
    t0 = time.time()
    b, a, pool = p.seeder_body(c.seed_reads)
    print("seed %d reads vs %d bp: device %.2f ms, wall %.1f ms, %d anchors, %d pool" % (n_reads, G, p.stats().last_kernel_ms, (time.time() - t0) * 1e3, len(a), len(pool)))

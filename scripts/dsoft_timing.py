"""D-SOFT on the GPU: table build and seeding throughput (no reference code involved).
Usage: python scripts/dsoft_timing.py [genome_bp] [n_reads] [read_len]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import darwin_b200
from darwin_b200 import abi, synth
G = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
read_len = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
t0 = time.time()
arena, anchors, hits = synth.anchor_batch(3, n_reads, read_len, G, err=(0.015, 0.09, 0.045))
print("workload built in %.1f s: arena %d MB" % (time.time() - t0, len(arena) >> 20))
p = darwin_b200.Processor(len(arena))
p.InitializeScoringParameters(abi.Scoring.from_values())
p.InitializeReferenceMemory(0, arena)
chroms = np.zeros(1, abi.CHROM); chroms["start"] = 128; chroms["len_unpadded"] = G
ref_size = 128 + G + ((-G) % 128)
t0 = time.time(); p.build_seed_index(abi.SeedParams.stock(), chroms, ref_size); t_ix = time.time() - t0
print("seed position table: %.2f s" % t_ix)
reads = np.zeros(n_reads, abi.SEED_READ); reads["read_addr"] = anchors["read_addr"]; reads["read_len"] = anchors["read_len"]
p.seeder_body(reads[:64])
for chunk in (n_reads, 2048, 512):
    t0 = time.time(); ms = 0.0; na = 0; npool = 0
    for lo in range(0, n_reads, chunk):
        b, a, pool = p.seeder_body(reads[lo:lo + chunk]); ms += p.stats().last_kernel_ms; na += len(a); npool += len(pool)
    wall = time.time() - t0
    print("seeding %d reads in batches of %d: device %.1f ms (%.0f reads/s), wall %.1f ms (%.0f reads/s); %d anchors (%.2f per read), %d chained hits (%.0f per anchor)" % (
        n_reads, chunk, ms, n_reads / ms * 1e3, wall * 1e3, n_reads / wall, na, na / n_reads, npool, npool / max(na, 1)))

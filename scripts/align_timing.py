"""The resident pipeline call (darwin_gpu_align_reads): seeding + first tiles + slope filter + extension of resident reads.
Usage: python scripts/align_timing.py [genome_bp] [n_reads] [read_len]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import darwin_b200
from darwin_b200 import abi, synth
G = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
read_len = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
arena, anchors, hits = synth.anchor_batch(3, n_reads, read_len, G, err=(0.015, 0.09, 0.045))
p = darwin_b200.Processor(len(arena))
p.InitializeScoringParameters(abi.Scoring.from_values())
p.InitializeReferenceMemory(0, arena)
chroms = np.zeros(1, abi.CHROM); chroms["start"] = 128; chroms["len_unpadded"] = G
p.build_seed_index(abi.SeedParams.stock(), chroms, 128 + G + ((-G) % 128))
reads = np.zeros(n_reads, abi.SEED_READ); reads["read_addr"] = anchors["read_addr"]; reads["read_len"] = anchors["read_len"]
p.align_reads(reads)
for chunk in (n_reads, 2048, 512, 128):
    t0 = time.time(); ms = 0.0; na = 0; al = 0; cells = 0
    for lo in range(0, n_reads, chunk):
        a, r, ops = p.align_reads(reads[lo:lo + chunk]); ms += p.stats().last_kernel_ms; na += len(a); al += int((r["flags"] & 1).sum()); cells += int(r["cells"].sum())
    wall = time.time() - t0
    print("align %d reads in batches of %d: kernels %.1f ms (%.0f reads/s), wall %.1f ms (%.0f reads/s); %d locations extended, %d alignments, %.0f GCUPS (wall)" % (
        n_reads, chunk, ms, n_reads / ms * 1e3, wall * 1e3, n_reads / wall, na, al, cells / wall / 1e9))

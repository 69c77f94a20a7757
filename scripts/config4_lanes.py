"""Config 4 (250 Mbp replicated reference, 10 kbp reads at 5/5/5) by host lanes and chunk size, for a whole 200 k-read set and
for the 25 k-read shard one rank of an 8-GPU run owns: where does the end-to-end rate of the read-level leg go when the shard
gets small?  One process, case and seed position table built once; every combination is one warm pass + two timed passes.
Usage: python scripts/config4_lanes.py [genome_bp] [n_reads]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from darwin_b200 import abi
genome = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000_000
n_total = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
sc = abi.Scoring.from_values()
MAXCH = 12500
t0 = time.time()
leg = bench.ReadLeg(0, sc, genome, n_total, 0, n_total, 10000, (0.05, 0.05, 0.05), 41, MAXCH, lanes=4)
leg.build_index(0)
print("case: %d reads against %d bp built in %.1f s (index %.3f s)" % (leg.n, genome, time.time() - t0, leg.index_s), flush=True)
prm = abi.AlignParams.stock(384, 64, 0)
leg.chunk = MAXCH
leg.one_pass(prm, n_reads=min(leg.n, 4 * MAXCH))                      # pools and buffers grown
for n_reads in (leg.n, 25000):
    for lanes in (2, 3, 4):
        for chunk in (12500, 6250, 4000, 3125):
            if chunk * 2 > n_reads:
                continue
            leg.chunk = chunk
            best = None
            for rep in range(2):
                t = leg.one_pass(prm, n_reads=n_reads, lanes=lanes)
                if best is None or t["wall_s"] < best["wall_s"]:
                    best = t
            print("reads %6d lanes %d chunk %5d: %7.1f ms -> %7.0f reads/s, %6.0f GCUPS end to end | summed event times: seed %.0f filter %.0f extend %.0f ms | checksum %d" % (
                n_reads, lanes, chunk, best["wall_s"] * 1e3, n_reads / best["wall_s"], best["cells"] / best["wall_s"] / 1e9,
                best["seed_ms"], best["filter_ms"], best["extend_ms"], best["score_sum"]), flush=True)
leg.close()

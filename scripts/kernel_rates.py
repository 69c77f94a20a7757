"""Device-resident rates of the tile kernels (config-2 style batches at several tile sizes) and of the extension kernel
(single lane, realistic D-SOFT anchors), each with a digest of the results: the A/B harness for kernel changes -- the
digests must not move.  Usage: python scripts/kernel_rates.py [n_tiles] [tiles: 256,320,384,512] [extend: 0|1] [cases]
(DARWIN_GPU_LIB selects another build of the library, DARWIN_GPU_MAX_CTAS_PER_SM caps the resident warps.)"""
import hashlib, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import darwin_b200
from darwin_b200 import abi, synth, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
tiles = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "256,320,384,512").split(",") if x]
do_extend = (sys.argv[3] if len(sys.argv) > 3 else "1") == "1"
only = sys.argv[4].split(",") if len(sys.argv) > 4 else None        # extension cases to run (default: all)
dev = torch.device("cuda", 0)

for tile in tiles:
    parts, reqs, base = [], [], 0
    nt = max(1000, int(n * (320.0 / tile) ** 2))
    for c0 in range(0, nt, 100000):
        a, r = synth.tile_batch_fast(1000 + c0 // 100000, min(100000, nt - c0), tile)
        a = a[:len(r) * 2 * tile]
        r["ref_bases_start_addr"] += base; r["query_bases_start_addr"] += base
        base += len(a); parts.append(a); reqs.append(r)
    arena = np.concatenate(parts + [np.full(128, ord("N"), np.uint8)]); req = np.concatenate(reqs)
    p = darwin_b200.Processor(len(arena), 0)
    p.InitializeScoringParameters(abi.Scoring.from_values())
    p.InitializeReferenceMemory(0, arena)
    tbw = 2 * tile // 16 + 2
    d_req = torch.from_numpy(req.view(np.uint8).reshape(len(req), -1)).to(dev)
    d_res = torch.zeros((len(req), 16), dtype=torch.uint8, device=dev)
    d_tb = torch.zeros((len(req), tbw), dtype=torch.int64, device=dev)
    ms = []
    for it in range(8):
        p.BatchAlignmentSIMD_device(d_req.data_ptr(), len(req), d_res.data_ptr(), d_tb.data_ptr(), tbw, tile, tile)
        if it >= 3:
            ms.append(p.stats().last_kernel_ms)
    st = p.stats()
    res = d_res.cpu().numpy().view(abi.TILE_RES).reshape(-1)
    tb = d_tb.cpu().numpy().view(np.uint64)
    used = (res["total_TB_pointers"].astype(np.int64) + 31) // 32
    mask = np.arange(tbw)[None, :] < used[:, None]
    dig = hashlib.sha1(res.tobytes() + tb[mask].tobytes()).hexdigest()[:16]
    print("tiles T=%d n=%d: %.3f ms -> %.0f GCUPS | fast %d rerun %d exact %d | sha1 %s" % (
        tile, len(req), np.mean(ms), len(req) * tile * tile / np.mean(ms) / 1e6, st.tiles_fast, st.tiles_rerun, st.tiles_exact, dig), flush=True)
    p.close()
    del d_req, d_res, d_tb

if do_extend:
    for which, (G, L, err, seed, nr, T) in {"config3": (10_000_000, 10000, (0.015, 0.09, 0.045), 31, 10000, 384),
                                            "config5_T256": (20_000_000, 50000, (0.04, 0.03, 0.05), 51, 2000, 256),
                                            "config5_T512": (20_000_000, 50000, (0.04, 0.03, 0.05), 51, 2000, 512),
                                            "config5_T1024": (20_000_000, 50000, (0.04, 0.03, 0.05), 51, 2000, 1024)}.items():
        if only and which not in only:
            continue
        case = workloads.ReadSetCase(dev, G, nr, 0, nr, L, err, seed)
        p = darwin_b200.Processor(case.arena_bytes, 0)
        p.InitializeScoringParameters(abi.Scoring.from_values())
        p.InitializeReferenceMemory(0, case.ref_numpy())
        p.InitializeReadMemory(case.read_addr(0), case.reads_numpy(0, case.n))
        p.build_seed_index(abi.SeedParams.stock(), case.chroms, case.ref_end)
        prm = abi.AlignParams.stock(T, 64, 0)
        p.align_reads(case.seed_reads[:min(2000, nr)], prm)
        best = None
        for rep in range(3):
            an, res, ops = p.align_reads(case.seed_reads, prm)
            st = p.stats()
            best = st.last_extend_ms if best is None else min(best, st.last_extend_ms)
        cells = float(res["cells"].sum())
        em = (res["flags"] & 1) != 0
        order = np.lexsort((res["query_start_offset"], res["reference_start_offset"], an["read_num"]))
        h = hashlib.sha1()
        for k in order:
            r = res[k]
            h.update(np.array([an["read_num"][k], r["score"], r["reference_start_offset"], r["reference_end_offset"], r["query_start_offset"],
                               r["query_end_offset"], r["n_ops"], r["flags"] & 1], np.int64).tobytes())
            if r["flags"] & 1:
                h.update(np.asarray(ops[int(r["ops_offset"]):int(r["ops_offset"]) + int(r["n_ops"])]).tobytes())
        print("extend %s T=%d: %d reads, %d alignments, %.3g cells, %.2f ms -> %.0f GCUPS | sha1 %s" % (
            which, T, nr, int(em.sum()), cells, best, cells / best / 1e6, h.hexdigest()[:16]), flush=True)
        p.close()
        del case
        torch.cuda.empty_cache()

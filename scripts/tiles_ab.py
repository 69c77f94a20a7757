"""A/B of the tile kernels on the config-2 workload (device-resident): the pair kernel (two tiles per warp) against the
one-tile-per-warp kernel (DARWIN_GPU_PAIRS=0), identical results required.  Usage: python scripts/tiles_ab.py [n_tiles] [tile]"""
import os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 320
if os.environ.get("TILES_AB_CHILD"):
    import torch
    import darwin_b200
    from darwin_b200 import abi, synth
    parts, reqs, base = [], [], 0
    for c0 in range(0, n, 100000):
        a, r = synth.tile_batch_fast(1000 + c0 // 100000, min(100000, n - c0), tile)
        a = a[:len(r) * 2 * tile]
        r["ref_bases_start_addr"] += base; r["query_bases_start_addr"] += base
        base += len(a); parts.append(a); reqs.append(r)
    arena = np.concatenate(parts + [np.full(128, ord("N"), np.uint8)]); req = np.concatenate(reqs)
    p = darwin_b200.Processor(len(arena), 0)
    p.InitializeScoringParameters(abi.Scoring.from_values())
    p.InitializeReferenceMemory(0, arena)
    tbw = 2 * tile // 16 + 2
    dev = torch.device("cuda", 0)
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
    import hashlib
    dig = hashlib.sha1(res.tobytes() + tb[mask].tobytes()).hexdigest()[:16]
    print("pairs=%s T=%d n=%d: %.3f ms -> %.0f GCUPS | fast %d paired %d rerun %d | sha1 %s" % (
        os.environ.get("DARWIN_GPU_PAIRS", "1"), tile, len(req), np.mean(ms), len(req) * tile * tile / np.mean(ms) / 1e6,
        st.tiles_fast, st.tiles_paired, st.tiles_rerun, dig))
else:
    for pairs in ("0", "1"):
        env = dict(os.environ, TILES_AB_CHILD="1", DARWIN_GPU_PAIRS=pairs)
        subprocess.run([sys.executable, os.path.abspath(__file__), str(n), str(tile)], env=env, check=True)

"""Seeded synthetic inputs for the GACT path: random references, mutated reads,
independent-tile batches (BASELINE.json configs[1]).  numpy only; shared by tests and bench.py."""
import numpy as np

from . import abi

_ACGT = np.frombuffer(b"ACGT", np.uint8)


def random_seq(rng, n):
    return _ACGT[rng.integers(0, 4, n)]


def mutate(rng, seq, sub=0.05, ins=0.05, dele=0.05, n_rate=0.0, indel_run=None):
    """Mutated copy of `seq` (uint8 ASCII): per-base substitution / insertion / deletion rates;
    indel_run=(count, max_len) additionally injects long indels."""
    n = len(seq)
    r = rng.random(n)
    keep = r >= dele
    out = seq.copy()
    s = (r >= dele) & (r < dele + sub)
    out[s] = _ACGT[(np.searchsorted(_ACGT, out[s]) + rng.integers(1, 4, int(s.sum()))) % 4]
    pieces = []
    insert_at = np.flatnonzero(rng.random(n) < ins)
    last = 0
    for p in insert_at:
        pieces.append(out[last:p][keep[last:p]])
        pieces.append(random_seq(rng, 1))
        last = p
    pieces.append(out[last:][keep[last:]])
    res = np.concatenate(pieces) if pieces else out
    if indel_run:
        cnt, mx = indel_run
        for _ in range(cnt):
            if len(res) < 8:
                break
            p = int(rng.integers(1, len(res) - 1))
            L = int(rng.integers(1, mx + 1))
            if rng.random() < 0.5:
                res = np.concatenate([res[:p], random_seq(rng, L), res[p:]])
            else:
                res = np.concatenate([res[:p], res[p + L:]])
    if n_rate > 0 and len(res):
        res = res.copy()
        res[rng.random(len(res)) < n_rate] = ord("N")
    return res


def mutate_fast(rng, seq, sub=0.05, ins=0.05, dele=0.05):
    """Vectorised mutate(): same error model (per-base substitution / insertion-before / deletion), no Python loop."""
    n = len(seq)
    r = rng.random(n)
    is_del = r < dele
    is_sub = (r >= dele) & (r < dele + sub)
    is_ins = rng.random(n) < ins
    codes = np.searchsorted(_ACGT, seq)
    codes = np.where(is_sub, (codes + rng.integers(1, 4, n)) % 4, codes)
    contrib = (~is_del).astype(np.int64) + is_ins.astype(np.int64)
    end = np.cumsum(contrib)
    start = end - contrib
    out = np.empty(int(end[-1]) if n else 0, np.uint8)
    out[start[is_ins]] = _ACGT[rng.integers(0, 4, int(is_ins.sum()))]
    keep = ~is_del
    out[(start + is_ins)[keep]] = _ACGT[codes[keep]]
    return out


def anchor_batch(seed, n_reads, read_len=10000, ref_len=1000000, err=(0.05, 0.05, 0.05), hit_spacing=60):
    """Synthetic reference-guided extension workload (SURVEY 8(d).3 without the host D-SOFT stage): an arena laid
    out like the reference's (Index.cpp:10-17, main.cpp:430-456, :645-686: 128 'N', one 'N'-padded chromosome,
    128-aligned 'N'-padded reads), one anchor per read at its true position (both strands) with chained hits every
    `hit_spacing` bases along the true diagonal (left list ascending, right list descending, seed_pos_table.cpp:432-490).
    Returns (arena uint8, anchors, hit_pool)."""
    rng = np.random.default_rng(seed)
    genome = random_seq(rng, ref_len)
    pad = (-ref_len) % 128
    parts = [np.full(128, ord("N"), np.uint8), genome, np.full(pad, ord("N"), np.uint8)]
    chr_start, chr_len = 128, ref_len + pad
    pos = 128 + chr_len
    anchors = np.zeros(n_reads, abi.ANCHOR)
    hit_parts, nh = [], 0
    for k in range(n_reads):
        L = read_len + int(rng.integers(0, read_len // 50 + 1))
        g0 = int(rng.integers(0, ref_len - L))
        src = genome[g0:g0 + L]
        read = mutate_fast(rng, src, *err)
        strand = k & 1
        fwd = revcomp(read) if strand else read            # the arena holds the forward read (main.cpp:662-670)
        rl = len(fwd)
        rpad = (-rl) % 128
        parts += [fwd, np.full(rpad, ord("N"), np.uint8)]
        qa = rl // 2
        ra = min(ref_len - 1, g0 + int(qa * len(src) / max(rl, 1)))
        a = anchors[k]
        a["read_addr"], a["read_len"], a["read_num"] = pos, rl, k
        a["reference_pos"], a["query_pos"] = chr_start + ra, qa
        a["chr_start"], a["ref_len"], a["chr_id"], a["score"], a["strand"] = chr_start, chr_len, 0, 100, strand
        dl = np.arange(0, min(ra, qa), hit_spacing, dtype=np.uint64)[::-1]
        dr = np.arange(0, min(ref_len - ra, rl - qa), hit_spacing - 1, dtype=np.uint64)[::-1]
        lh = ((np.uint64(chr_start + ra) - dl) << np.uint64(32)) | (np.uint64(qa) - dl)
        rh = ((np.uint64(chr_start + ra) + dr) << np.uint64(32)) | (np.uint64(qa) + dr)
        a["left_hits_off"], a["left_hits_n"] = nh, len(lh)
        nh += len(lh)
        a["right_hits_off"], a["right_hits_n"] = nh, len(rh)
        nh += len(rh)
        hit_parts += [lh, rh]
        pos += rl + rpad
    arena = np.concatenate(parts + [np.full(128, ord("N"), np.uint8)])
    return arena, anchors, (np.concatenate(hit_parts) if hit_parts else np.zeros(0, np.uint64))


def revcomp(seq):
    """main.cpp:59-121 semantics on uint8 ASCII (case preserved)."""
    lut = np.full(256, ord("N"), np.uint8)
    for a, b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
        lut[a] = b
    return lut[seq[::-1]]


def tile_batch(seed, n, tile=320, err=(0.05, 0.05, 0.05), max_tb_steps=None, mode="extend"):
    """BASELINE.json configs[1] / SURVEY 8(d).2: n independent tiles.  Returns (arena uint8, requests).
    ref = `tile` random bases; query = mutated copy cut/padded to `tile`; half the tiles are left
    extensions (start_end), half right extensions (reverse_ref|reverse_query|start_end).
    Each tile owns 2*tile arena bytes: [ref | query]."""
    rng = np.random.default_rng(seed)
    arena = np.empty(n * 2 * tile + 128, np.uint8)
    arena[:] = ord("N")
    req = np.zeros(n, abi.TILE_REQ)
    # vectorised generation: mutate one long sequence, then cut per-tile windows
    for k in range(n):
        ref = random_seq(rng, tile)
        q = mutate(rng, ref, *err)
        if len(q) < tile:
            q = np.concatenate([q, random_seq(rng, tile - len(q))])
        base = k * 2 * tile
        arena[base:base + tile] = ref
        arena[base + tile:base + 2 * tile] = q[:tile]
    _fill_tile_req(req, n, tile, max_tb_steps, mode)
    return arena, req


def tile_batch_fast(seed, n, tile=320, err=(0.05, 0.05, 0.05), max_tb_steps=None, mode="extend"):
    """Same distribution as tile_batch but fully vectorised (for the 1M-tile bench workload):
    edits are applied with per-tile cumulative offsets instead of per-tile Python loops."""
    rng = np.random.default_rng(seed)
    sub, ins, dele = err
    W = tile + tile // 2                       # source window long enough to survive deletions
    src = rng.integers(0, 4, (n, W), dtype=np.uint8)
    r = rng.random((n, W), dtype=np.float32)
    is_del = r < dele
    is_sub = (r >= dele) & (r < dele + sub)
    is_ins = rng.random((n, W), dtype=np.float32) < ins
    q_src = np.where(is_sub, (src + rng.integers(1, 4, (n, W), dtype=np.uint8)) % 4, src)
    # output length contribution of each source position: kept base (0/1) + inserted base before it (0/1)
    contrib = (~is_del).astype(np.int32) + is_ins.astype(np.int32)
    end = np.cumsum(contrib, axis=1)
    start = end - contrib
    q = rng.integers(0, 4, (n, tile), dtype=np.uint8)          # default = random padding
    rows = np.arange(n)[:, None].repeat(W, 1)
    ins_pos = start                                            # inserted base first, then the kept base
    m = is_ins & (ins_pos < tile)
    q[rows[m], ins_pos[m]] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
    keep_pos = start + is_ins.astype(np.int32)
    m = (~is_del) & (keep_pos < tile)
    q[rows[m], keep_pos[m]] = q_src[m]
    arena = np.full(n * 2 * tile + 128, ord("N"), np.uint8)
    body = arena[:n * 2 * tile].reshape(n, 2 * tile)
    body[:, :tile] = _ACGT[src[:, :tile]]
    body[:, tile:] = _ACGT[q]
    req = np.zeros(n, abi.TILE_REQ)
    _fill_tile_req(req, n, tile, max_tb_steps, mode)
    return arena, req


def _fill_tile_req(req, n, tile, max_tb_steps, mode):
    k = np.arange(n, dtype=np.uint64)
    req["ref_bases_start_addr"] = k * np.uint64(2 * tile)
    req["query_bases_start_addr"] = k * np.uint64(2 * tile) + np.uint64(tile)
    req["ref_size"] = tile
    req["query_size"] = tile
    req["max_tb_steps"] = max_tb_steps if max_tb_steps is not None else 2 * tile
    req["index"] = (k % 65536).astype(np.uint16)
    if mode == "extend":
        left = abi.START_END
        right = abi.REVERSE_REF | abi.REVERSE_QUERY | abi.START_END
        req["align_fields"] = np.where(k % 2 == 0, left, right).astype(np.uint8)
    elif mode == "filter":
        req["align_fields"] = 0
    else:
        raise ValueError(mode)


def consumed_ops(tb_words, total, S):
    """The extender's consumption rule on one tile's TB words (extender.cpp:280-331): ops are taken
    word by word; once `steps >= S` an M ends the CURRENT 32-op word only (the quirk)."""
    out = []
    steps = 0
    for w in range(0, total, 32):
        word = int(tb_words[w // 32])
        for p in range(min(32, total - w)):
            d = (word >> (2 * p)) & 3
            out.append(d)
            steps += 1
            if steps >= S and d == abi.OP_M:
                break
    return out

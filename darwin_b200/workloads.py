"""Synthetic read-level workloads of BASELINE.json configs[2..4] for bench.py and the timing scripts.

The sequences are made on the GPU with torch (plumbing: a 250 Mbp reference and 200 000 reads of 10 kbp are 2.5 GB of
ASCII; numpy needs minutes for that, the GPU a second) and are then moved to PAGE-LOCKED HOST memory, because the
end-to-end legs start from host buffers: every timed step uploads its reads.  Same error model as synth.mutate_fast
(per source base: deletion, substitution, insertion before the base), reads of exactly `read_len` bases, odd reads
stored reverse-complemented (the arena holds the forward read of a '-' strand alignment, main.cpp:662-670).

The read set is a pure function of (seed, global read index / BLOCK): a rank that owns reads [lo, hi) of the set generates
exactly those, whatever the number of ranks -- the strong-scaling leg (configs[3]) shards ONE fixed set.
"""
import numpy as np
import torch

from . import abi

BLOCK = 1000                       # reads generated per RNG block (shard boundaries must be multiples of it)
WORD = 128                         # DRAM.h:4 WORD_SIZE: sequences are padded with 'N' to a multiple of it


def _ascii_lut(device):
    return torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=device)          # A C G T


def genome_codes(n, seed, device):
    """n uniform bases as codes 0..3 (uint8) on `device`; identical on every rank for the same seed."""
    g = torch.Generator(device=device)
    g.manual_seed(0x9E3779B1 * 7 + seed)
    return torch.randint(0, 4, (n,), dtype=torch.uint8, device=device, generator=g)


def read_stride(read_len):
    return read_len + ((-read_len) % WORD)


def simulate_block(genome, block_index, seed, read_len, err, n=BLOCK):
    """Reads [block_index * BLOCK, +n) of the set: (n, stride) ASCII uint8 tensor on genome.device, 'N'-padded rows."""
    dev = genome.device
    sub, ins, dele = err
    W = int(read_len * (1.0 + dele + 0.06)) + 64                                    # source window that survives the deletions
    G = genome.numel()
    g = torch.Generator(device=dev)
    g.manual_seed((seed * 1000003 + block_index) & 0x7FFFFFFFFFFF)
    g0 = torch.randint(0, G - W, (n,), device=dev, generator=g)
    src = genome[(g0[:, None] + torch.arange(W, device=dev)[None, :])]               # n x W codes
    r = torch.rand((n, W), device=dev, generator=g)
    is_del = r < dele
    is_sub = (r >= dele) & (r < dele + sub)
    is_ins = torch.rand((n, W), device=dev, generator=g) < ins
    shift = torch.randint(1, 4, (n, W), dtype=torch.uint8, device=dev, generator=g)
    q = torch.where(is_sub, (src + shift) & 3, src)
    ins_base = torch.randint(0, 4, (n, W), dtype=torch.uint8, device=dev, generator=g)
    contrib = (~is_del).to(torch.int32) + is_ins.to(torch.int32)
    end = torch.cumsum(contrib, 1)
    start = end - contrib
    if int(end[:, -1].min()) < read_len:
        raise RuntimeError("workloads: source window too short for this error profile")
    stride = read_stride(read_len)
    codes = torch.zeros((n, read_len), dtype=torch.uint8, device=dev)
    rows = torch.arange(n, device=dev)[:, None].expand(n, W)
    m = is_ins & (start < read_len)
    codes[rows[m], start[m]] = ins_base[m]
    kp = start + is_ins.to(torch.int32)
    m = (~is_del) & (kp < read_len)
    codes[rows[m], kp[m]] = q[m]
    # odd reads of the SET (global index) are stored as their reverse complement
    odd = ((torch.arange(n, device=dev) + block_index * BLOCK) & 1).bool()
    rc = (3 - codes).flip(1)
    codes = torch.where(odd[:, None], rc, codes)
    out = torch.full((n, stride), 78, dtype=torch.uint8, device=dev)                 # 'N'
    out[:, :read_len] = _ascii_lut(dev)[codes.long()]
    return out


class ReadSetCase:
    """One reference + the shard [lo, hi) of a fixed simulated read set, laid out like the reference's arena
    (Index.cpp:10-17, main.cpp:430-456, :645-686): [128 'N'][chromosome, 'N'-padded][reads, 128-aligned, 'N'-padded].

    host_ref   : page-locked ASCII of the reference region (arena offset 0 .. ref_end)
    host_reads : page-locked (n, stride) ASCII of this shard's reads; read k lives at arena offset ref_end + k * stride
    """

    def __init__(self, device, genome_len, n_total, lo, hi, read_len, err, seed):
        assert lo % BLOCK == 0 and (hi % BLOCK == 0 or hi == n_total), "shard boundaries must be multiples of workloads.BLOCK"
        self.genome_len, self.read_len, self.n = genome_len, read_len, hi - lo
        self.stride = read_stride(read_len)
        pad = (-genome_len) % WORD
        self.chr_start, self.chr_len = WORD, genome_len + pad
        self.ref_end = WORD + self.chr_len
        self.arena_bytes = self.ref_end + self.n * self.stride + WORD
        genome = genome_codes(genome_len, seed, device)
        self.host_ref = torch.full((self.ref_end,), 78, dtype=torch.uint8).pin_memory()
        lut = _ascii_lut(device)
        step = 32 << 20
        for a in range(0, genome_len, step):
            b = min(genome_len, a + step)
            self.host_ref[WORD + a:WORD + b].copy_(lut[genome[a:b].long()])
        self.host_reads = torch.empty((self.n, self.stride), dtype=torch.uint8).pin_memory()
        for blk in range(lo // BLOCK, (hi + BLOCK - 1) // BLOCK):
            k0 = blk * BLOCK
            cnt = min(BLOCK, hi - k0)
            self.host_reads[k0 - lo:k0 - lo + cnt].copy_(simulate_block(genome, blk, seed, read_len, err)[:cnt])
        torch.cuda.synchronize(device)
        del genome
        self.chroms = np.zeros(1, abi.CHROM)
        self.chroms["start"], self.chroms["len_unpadded"] = self.chr_start, genome_len
        self.seed_reads = np.zeros(self.n, abi.SEED_READ)
        self.seed_reads["read_addr"] = self.ref_end + np.arange(self.n, dtype=np.uint64) * np.uint64(self.stride)
        self.seed_reads["read_len"] = read_len

    def ref_numpy(self):
        return self.host_ref.numpy()

    def reads_numpy(self, a, b):
        """ASCII of reads [a, b) of the shard as ONE contiguous page-locked span (rows are back to back)."""
        return self.host_reads[a:b].numpy().reshape(-1)

    def read_addr(self, k):
        return self.ref_end + k * self.stride

import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import darwin_b200
from darwin_b200 import abi, synth
def pinned(shape, dtype):
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    return t.numpy().view(dtype).reshape(shape)
arena, anchors, hits = synth.anchor_batch(7, 8000, 10000, 4000000)
p = darwin_b200.Processor(len(arena)); p.InitializeScoringParameters(abi.Scoring.from_values()); p.InitializeReferenceMemory(0, arena)
p.extender_body(anchors[:64], hits, 384, 64, 0)
for name in ("pageable", "pageable2", "pinned", "pinned2"):
    if name.startswith("pinned"):
        t0=time.perf_counter(); out = (pinned((len(anchors),), abi.ALN_RES), pinned((int(anchors["read_len"].sum())*2,), np.uint8)); ta=time.perf_counter()-t0
    else:
        out = None; ta = 0
    t0=time.perf_counter(); res, ops = p.extender_body(anchors, hits, 384, 64, 0, out=out); dt=time.perf_counter()-t0
    print(name, "alloc %.3f s, call %.3f s, kernel %.1f ms" % (ta, dt, p.stats().last_kernel_ms), flush=True)

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
os.environ['DARWIN_GPU_DEBUG'] = '/tmp/dbg.bin'
import darwin_b200, oracle
from darwin_b200 import abi
oracle.build('port')
g = np.load('tests/golden/extend_v1.npz')
tag = 'T384_O64_ovl0'
arena = g['arena']; sc = abi.Scoring.from_values(*g['scoring'].tolist())
port = oracle.port(sc)
anchors = g[tag+'_anchors']; hits = g[tag+'_hits']
p = darwin_b200.Processor(len(arena)); p.InitializeScoringParameters(sc); p.InitializeReferenceMemory(0, arena)
res, ops = p.extender_body(anchors, hits, 384, 64, 0)
d = np.fromfile('/tmp/dbg.bin', np.uint32).reshape(-1, 128, 8)
for a in range(len(anchors)):
    log = np.zeros(200, abi.TILE_REQ)
    port.lib.gact_set_tile_log(abi.ptr(log), len(log))
    pres, pops = port.extend(arena, abi.ExtendParams(384,64,0,0), anchors[a:a+1], hits, 1)
    n = port.lib.gact_tile_log_count(); port.lib.gact_set_tile_log(None, 0)
    ok = True
    for k in range(n):
        L = log[k]
        if (int(L['ref_size']), int(L['query_size']), int(L['score_threshold']), int(L['ref_bases_start_addr']), int(L['query_bases_start_addr'])) != tuple(int(x) for x in d[a,k,:5]):
            print('anchor', a, 'tile', k, 'of', n, 'port R,Q,len,cr,cq', int(L['ref_size']), int(L['query_size']), int(L['score_threshold']), int(L['ref_bases_start_addr']), int(L['query_bases_start_addr']), ' gpu', d[a,k].tolist(), 'prev gpu', d[a,k-1].tolist() if k else None, 'flags', int(L['align_fields']))
            ok = False; break
    if ok: print('anchor', a, 'OK', n, 'tiles')

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import darwin_b200
from darwin_b200 import abi
g = np.load('tests/golden/extend_v1.npz')
tag = 'T384_O64_ovl0'
arena = g['arena']
p = darwin_b200.Processor(len(arena)); p.InitializeScoringParameters(abi.Scoring.from_values(*g['scoring'].tolist()))
p.InitializeReferenceMemory(0, arena)
res, ops = p.extender_body(g[tag+'_anchors'], g[tag+'_hits'], 384, 64, 0)
gr, go = g[tag+'_res'], g[tag+'_ops']
for k in range(len(res)):
    a, b = gr[k], res[k]
    diff = [f for f in a.dtype.names if a[f] != b[f] and f not in ('ops_offset','n_left_ops')]
    oa = go[int(a['ops_offset']):int(a['ops_offset'])+int(a['n_ops'])]; ob = ops[int(b['ops_offset']):int(b['ops_offset'])+int(b['n_ops'])]
    same_ops = np.array_equal(oa, ob)
    first = -1
    if not same_ops:
        m = min(len(oa), len(ob)); d = np.flatnonzero(oa[:m] != ob[:m]); first = int(d[0]) if len(d) else m
    print(k, 'strand', g[tag+'_anchors'][k]['strand'], 'diff', {f: (int(a[f]), int(b[f])) for f in diff}, 'ops_same', same_ops, 'first_diff', first, 'nleft', int(b['n_left_ops']))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import darwin_b200, oracle
from darwin_b200 import abi, gact
if len(sys.argv) > 1:
    gact._LIB_PATH = sys.argv[1]
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import alignments_equal, ALN_FIELDS_OURS
g = np.load('tests/golden/extend_v1.npz')
tag = 'T384_O64_ovl0'
arena = g['arena']; sc = abi.Scoring.from_values(*g['scoring'].tolist())
port = oracle.port(sc)
anchors = g[tag+'_anchors']; hits = g[tag+'_hits']
pres, pops = port.extend(arena, abi.ExtendParams(384,64,0,0), anchors, hits, 1)
p = darwin_b200.Processor(len(arena)); p.InitializeScoringParameters(sc); p.InitializeReferenceMemory(0, arena)
for rep in range(3):
    res, ops = p.extender_body(anchors, hits, 384, 64, 0)
    print('batch run', rep, 'bad', alignments_equal(pres, pops, res, ops, ALN_FIELDS_OURS))
bad1 = []
for k in range(len(anchors)):
    res, ops = p.extender_body(anchors[k:k+1], hits, 384, 64, 0)
    if alignments_equal(pres[k:k+1], pops, res, ops, ALN_FIELDS_OURS): bad1.append(k)
print('one-at-a-time bad', bad1)

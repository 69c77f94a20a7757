import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import darwin_b200, oracle
from darwin_b200 import abi
oracle.build('port')
g = np.load('tests/golden/extend_v1.npz')
tag = 'T384_O64_ovl0'
arena = g['arena']; sc = abi.Scoring.from_values(*g['scoring'].tolist())
port = oracle.port(sc)
log = np.zeros(2000, abi.TILE_REQ)
port.lib.gact_set_tile_log(abi.ptr(log), len(log))
pres, pops = port.extend(arena, abi.ExtendParams(384,64,0,0), g[tag+'_anchors'], g[tag+'_hits'], 1)
n = port.lib.gact_tile_log_count(); port.lib.gact_set_tile_log(None, 0)
log = log[:n]
print('tiles logged', n)
p = darwin_b200.Processor(len(arena)); p.InitializeScoringParameters(sc); p.InitializeReferenceMemory(0, arena)
res, tb = p.BatchAlignmentSIMD(log, 1, 100)
ores, otb, _ = port.tiles(arena, log, 1, 1, tb_words_per_req=100)
bad = [k for k in range(n) if res[k] != ores[k] or not np.array_equal(tb[k], otb[k])]
print('bad tiles', len(bad), bad[:20])
for k in bad[:5]:
    print(log[k], res[k], ores[k])

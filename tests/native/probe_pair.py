"""Bring-up tool (GPU): prints where a cta_group::2 MMA of M rows puts D in each CTA's TMEM."""
import sys
import numpy as np
sys.path.insert(0, ".")
from stt_b200 import api

for M in (256, 128):
    out = np.zeros((2, 128, 128), np.float32)
    rc = api.dev_lib().STTX_DebugPairLayout(M, out.ctypes.data)
    print("M=%d rc=%d" % (M, rc))
    v = out.astype(np.int64)
    row = v // 1024 - 1
    col = v % 1024 - 1
    for cta in range(2):
        for lane in (0, 1, 31, 32, 63, 64, 65, 96, 127):
            r, c = row[cta, lane], col[cta, lane]
            print("  cta %d lane %3d: cols 0..3 -> (row,col) %s | col 63,64,127 -> %s" % (
                cta, lane, [(int(r[q]), int(c[q])) for q in range(4)], [(int(r[q]), int(c[q])) for q in (63, 64, 127)]))

"""One Scan-Context query batch against an n-row database (for ncu captures): python tools/sc_one.py [n] [nq] [stream|tile|auto] [reps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from direct_stereo_slam_b200 import api, synthetic as syn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 32
flavour = sys.argv[3] if len(sys.argv) > 3 else "auto"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
s = api.Session(0)
sig, key = syn.make_sc_database(n, 2024)
db = api.ScanContextDB(s, n)
db.add(key, sig)
db.set_scan_kernel(flavour)
qs, qk, truth = syn.make_sc_queries(sig, key, nq, 77)
ms = []
for _ in range(reps):
    idx, diff = db.query(qs)
    ms.append(db.last_scan_ms())
known = truth >= 0
print("n=%d Q=%d %s: scan %.3f ms (median of %d), %d/%d revisits found" % (n, nq, flavour, float(np.median(ms)), reps, int((idx[known] == truth[known]).sum()), int(known.sum())))

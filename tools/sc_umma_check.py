"""tcgen05 (3xTF32) Scan-Context scan vs the FFMA tile scan on the same database: same argmin and distance bits after the exact
re-score, kernel time of both.  python tools/sc_umma_check.py [rows] [nq] [umma|umma_masked]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from direct_stereo_slam_b200 import api, synthetic as syn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 32
UMMA = sys.argv[3] if len(sys.argv) > 3 else "umma"
s = api.Session(0)
sig, key = syn.make_sc_database(n, 2024)
db = api.ScanContextDB(s, n)
db.add(key, sig)
qs, qk, truth = syn.make_sc_queries(sig, key, nq, 77)
res = {}
for flavour in ("tile", UMMA):
    db.set_scan_kernel(flavour)
    ms = []
    for _ in range(6):
        idx, diff = db.query(qs)
        ms.append(db.last_scan_ms())
    res[flavour] = (idx.copy(), diff.copy(), float(np.median(ms)))
    known = truth >= 0
    print("%-5s n=%d Q=%d: scan %.3f ms, %d/%d revisits found" % (flavour, n, nq, res[flavour][2], int((idx[known] == truth[known]).sum()), int(known.sum())), flush=True)
db.set_scan_kernel(UMMA)
i_g, d_g = db.query(qs, ringkeys=qk, ringkey_thres=0.5, max_id=n // 2)
db.set_scan_kernel("tile")
i_t, d_t = db.query(qs, ringkeys=qk, ringkey_thres=0.5, max_id=n // 2)
same = np.array_equal(res["tile"][0], res[UMMA][0]) and np.array_equal(res["tile"][1].view(np.uint32), res[UMMA][1].view(np.uint32))
same_gate = np.array_equal(i_g, i_t) and np.array_equal(d_g.view(np.uint32), d_t.view(np.uint32))
print("umma == tile: %s ; with ring-key gate + max_id: %s ; speed-up %.2fx" % (same, same_gate, res["tile"][2] / res[UMMA][2]))
sys.exit(0 if same and same_gate else 1)

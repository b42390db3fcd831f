"""Scan-Context scan kernel timing vs query batch size on one GPU (100k descriptors)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from direct_stereo_slam_b200 import api, synthetic as syn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
s = api.Session(0)
sig, key = syn.make_sc_database(n, 2024)
db = api.ScanContextDB(s, n)
db.add(key, sig)
flavours = sys.argv[2].split(",") if len(sys.argv) > 2 else ["stream", "tile", "auto"]
for flavour in flavours:
  db.set_scan_kernel(flavour)
  print("--- scan kernel:", flavour)
  for nq in (1, 2, 4, 8, 16, 32, 64, 256):
      qs, qk, truth = syn.make_sc_queries(sig, key, nq, 77)
      for _ in range(3):
          db.query(qs)
      scan, lat = [], []
      for _ in range(20):
          t0 = time.perf_counter()
          idx, diff = db.query(qs)
          lat.append((time.perf_counter() - t0) * 1e3)
          scan.append(db.last_scan_ms())
      sm, lm = np.median(scan), np.median(lat)
      known = truth >= 0
      print("Q=%3d scan %.3f ms (%.0f GB/s of 4880 B/row/batch-of-32) e2e latency %.3f ms  found %d/%d" % (
          nq, sm, n * 4880 * ((nq + 31) // 32) / sm / 1e6, lm, int((idx[known] == truth[known]).sum()), int(known.sum())))

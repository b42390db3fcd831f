"""Pins the oracle's restatements against the REFERENCE'S OWN CODE where that compiles in place (oracle/ref_build.py ->
oracle/_ref/libdslam_ref.so): Accumulator9 (deps:dso MatrixAccumulators.h:982-1345), ScaleAccumulator
(src/scale_optimization/ScaleAccumulator.h) and search_ringkey / search_sc (src/loop_closure/loop_detection/search_place.h)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as orc
from helpers import OracleCase

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def ref():
    if not orc.ReferencePieces.available():
        if os.path.isdir("/root/reference"):
            subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_build.py")], check=True)
        else:
            pytest.skip("oracle/_ref not built and /root/reference not present")
    return orc.ReferencePieces()


def test_accumulator9_bit_exact_on_tracker_rows(oracle, ref):
    oc = OracleCase(oracle, "tiny", 3)
    for lvl in range(oc.levels):
        for pose, aff in ((oc.case["pose7_true"], (0.01, 1.0)), (np.array([0, 0, 0, 1, 0, 0, 0.0]), (0.0, 0.0))):
            res, n = oc.trk.calc_res_pose(lvl, pose, aff)
            J, w = oc.trk.pose_rows(lvl, aff)
            a_ref = ref.accumulator9(J, w)
            assert np.array_equal(oracle.accumulator9(J, w), a_ref)
            H, b, acc = oc.trk.calc_gs_pose(lvl, 0, aff)  # the oracle's calcGSSSEPose goes through the same accumulator
            assert np.array_equal(acc, a_ref.astype(np.float64))


def test_accumulators_tier_shifts(oracle, ref):
    """> 1000 and > 1e6 updates exercise shiftUp's 1k and 1m tiers (MatrixAccumulators.h:1325-1344)."""
    rng = np.random.default_rng(0)
    for n in (4, 4000, 4004, 1_001_000 * 4 + 8):
        J = rng.normal(size=(9, n)).astype(np.float32)
        w = rng.uniform(0, 1, n).astype(np.float32)
        assert np.array_equal(oracle.accumulator9(J, w), ref.accumulator9(J, w))
        assert np.array_equal(oracle.scale_accumulator(J[0], J[1], w), ref.scale_accumulator(J[0], J[1], w))


def test_scale_accumulator_on_tracker_rows(oracle, ref):
    oc = OracleCase(oracle, "tiny", 4, scale_error=2.5)
    for lvl in range(oc.levels):
        for s in (0.5, 1.0, 2.5):
            oc.trk.calc_res_scale(lvl, s)
            J, r, w = oc.trk.scale_rows(lvl, s)
            a_ref = ref.scale_accumulator(J, r, w)
            H, b, acc = oc.trk.calc_gs_scale(lvl, 0, s)
            assert np.array_equal(acc, a_ref.astype(np.float64))
            n = len(J)
            assert H == np.float32(a_ref[0]) * (np.float32(1.0) / np.float32(n)) and b == np.float32(a_ref[1]) * (np.float32(1.0) / np.float32(n))


def test_search_sc_matches_reference(oracle, ref):
    rng = np.random.default_rng(1)
    ptr, idxs, vals = [0], [], []
    for r in range(40):
        rk, si, sv, _ = oracle.sc_generate(rng.normal(0, 12, (2000, 3)))
        idxs.append(si)
        vals.append(sv)
        ptr.append(ptr[-1] + len(si))
    sidx, sval = np.concatenate(idxs), np.concatenate(vals)
    for q in range(20):
        rk, qi, qv, _ = oracle.sc_generate(rng.normal(0, 12, (2000, 3)))
        cands = rng.choice(40, int(rng.integers(1, 6)), replace=False).astype(np.int32)
        assert oracle.search_sc(qi, qv, ptr, sidx, sval, cands) == ref.search_sc(qi, qv, ptr, sidx, sval, cands)
    # identical signature: distance (1 - 60/60)/2 up to rounding, first of equal candidates wins (strict '>')
    i0, d0 = ref.search_sc(idxs[7], vals[7], ptr, sidx, sval, np.array([3, 7, 7], np.int32))
    assert i0 == 7 and abs(d0) < 1e-6
    assert oracle.search_sc(idxs[7], vals[7], ptr, sidx, sval, np.array([3, 7, 7], np.int32)) == (i0, d0)


def test_search_ringkey_queue_semantics(oracle, ref):
    """The reference's search_ringkey (static LOOP_MARGIN queue + dummy row 0) run in its own process, against the
    contract of dslam_sc_search_ringkey: exact 3-NN among ids < call_index - 100, kept when dist < 0.1."""
    code = """
import sys, numpy as np
sys.path.insert(0, %r)
import oracle as orc
rng = np.random.default_rng(5)
base = rng.uniform(0, 1, (12, 20))
keys = (base[rng.integers(0, 12, 260)] + rng.normal(0, 0.03, (260, 20))).astype(np.float32)
cand, ncand = orc.ReferencePieces().search_ringkey_sequence(keys)
np.save(sys.argv[1], cand)
np.save(sys.argv[2], keys)
""" % ROOT
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        a, b = os.path.join(tmp, "cand.npy"), os.path.join(tmp, "keys.npy")
        subprocess.run([sys.executable, "-c", code, a, b], check=True)
        cand, keys = np.load(a), np.load(b)
    some = 0
    for c in range(len(keys)):
        max_id = c - 100
        expect = np.full(3, -1, np.int32)
        if max_id >= 3:  # "ringkeys->size() > FLANN_NN": dummy row + at least 3 real keys
            # the dummy all-zero row competes for the 3 nearest but is never returned (idx > 0)
            rows = np.concatenate([np.zeros((1, 20), np.float32), keys[:max_id]])
            ci, di = oracle.search_ringkey(keys[c], rows, k=3, thres=0.1)
            ci = ci[ci > 0] - 1
            expect[:len(ci)] = ci
        assert np.array_equal(cand[c], expect), c
        some += int((expect >= 0).sum())
    assert some > 50

"""Scan-Context database scan vs the oracle's restatement of search_ringkey / search_sc: argmin index bit-exact,
distance bit-exact (the survivors are re-scored on the device in the reference's float += double*double order)."""
import numpy as np
import pytest

from direct_stereo_slam_b200 import api
from direct_stereo_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _oracle_query(oracle, qs, db, candidates=None):
    idx, diff = [], []
    for q in qs:
        i, d = oracle.search_sc_dense(q, db, candidates)
        idx.append(i)
        diff.append(d)
    return np.array(idx, np.int32), np.array(diff, np.float32)


@pytest.fixture(scope="module")
def small_db(session):
    sig, key = syn.make_sc_database(3000, 7)
    db = api.ScanContextDB(session, 4096)
    db.add(key[:1000], sig[:1000])
    db.add(key[1000:], sig[1000:], global_ids=np.arange(1000, 3000))
    yield db, sig, key
    db.close()


@pytest.mark.parametrize("nq", [1, 5, 32, 33, 70])
def test_query_argmin_bit_exact(small_db, oracle, nq):
    db, sig, key = small_db
    qs, qk, truth = syn.make_sc_queries(sig, key, nq, 100 + nq)
    idx, diff = db.query(qs)
    idx_o, diff_o = _oracle_query(oracle, qs, sig)
    assert np.array_equal(idx, idx_o)
    assert np.array_equal(diff.view(np.uint32), diff_o.view(np.uint32))
    known = truth >= 0
    assert np.array_equal(idx[known], truth[known])  # perturbed revisits find their row


@pytest.fixture(params=["stream", "tile"])
def scan_kernel(request, small_db):
    """Force one of the two scan kernels (dslam_sc_set_scan_kernel) for the test, then restore the automatic choice."""
    small_db[0].set_scan_kernel(request.param)
    yield request.param
    small_db[0].set_scan_kernel("auto")


@pytest.mark.parametrize("nq", [1, 7, 32, 33, 70])
def test_both_scan_kernels_bit_exact(small_db, oracle, scan_kernel, nq):
    """The HBM-streaming kernel and the register-blocked TMA tile kernel feed the same exact re-score: identical answers
    on a database whose row count (3000) is not a multiple of the 256-row tile and for ragged query batches."""
    db, sig, key = small_db
    qs, qk, truth = syn.make_sc_queries(sig, key, nq, 300 + nq)
    idx, diff = db.query(qs)
    idx_o, diff_o = _oracle_query(oracle, qs, sig)
    assert np.array_equal(idx, idx_o)
    assert np.array_equal(diff.view(np.uint32), diff_o.view(np.uint32))


def test_query_max_id_and_ringkey_gate(small_db, oracle, scan_kernel):
    db, sig, key = small_db
    qs, qk, truth = syn.make_sc_queries(sig, key, 16, 5)
    # only ids < max_id compete (the LOOP_MARGIN delay of search_ringkey, search_place.h:42-56)
    idx, diff = db.query(qs, max_id=1500)
    idx_o, diff_o = _oracle_query(oracle, qs, sig[:1500])
    assert np.array_equal(idx, idx_o) and np.array_equal(diff.view(np.uint32), diff_o.view(np.uint32))
    # ring-key gate: rows whose squared-L2 ring-key distance is >= thres are skipped
    thres = 0.02
    idx, diff = db.query(qs, ringkeys=qk, ringkey_thres=thres)
    for q in range(3):
        d2 = np.array([oracle.search_ringkey(qk[q], key[r:r + 1], k=1, thres=np.inf)[1][0] for r in range(len(key))])
        cand = np.nonzero(d2 < thres)[0].astype(np.int32)
        if len(cand) == 0:
            assert idx[q] == -1
        else:
            i_o, d_o = oracle.search_sc_dense(qs[q], sig, cand)
            assert idx[q] == i_o and diff[q] == np.float32(d_o)


def test_query_ties_go_to_lowest_id(session, oracle):
    sig, key = syn.make_sc_database(64, 3)
    sig = np.concatenate([sig, sig[10:11], sig[10:11], sig[10:11]])  # rows 64, 65, 66 duplicate row 10
    key = np.concatenate([key, key[10:11], key[10:11], key[10:11]])
    db = api.ScanContextDB(session, 128)
    db.add(key, sig)
    idx, diff = db.query(sig[10:11])
    assert idx[0] == 10
    i_o, d_o = oracle.search_sc_dense(sig[10], sig)
    assert i_o == 10 and diff[0] == np.float32(d_o)
    # more duplicates than the device keeps per query (K = 8): still the lowest id
    sig2 = np.repeat(sig[3:4], 40, 0)
    db2 = api.ScanContextDB(session, 64)
    db2.add(np.repeat(key[3:4], 40, 0), sig2, global_ids=np.arange(100, 140))
    idx, _ = db2.query(sig[3:4])
    assert idx[0] == 100
    db.close()
    db2.close()


def test_empty_and_tiny_database(session):
    db = api.ScanContextDB(session, 16)
    sig, key = syn.make_sc_database(3, 1)
    idx, diff = db.query(sig[:2])
    assert np.array_equal(idx, [-1, -1]) and np.allclose(diff, 1.1)
    db.add(key[:1], sig[:1])
    idx, diff = db.query(sig[:2])
    assert np.array_equal(idx, [0, 0])
    cand, dist = db.search_ringkey(key[:2], k=3, thres=0.1)
    assert cand[0, 0] == 0 and dist[0, 0] == 0 and np.all(cand[:, 1:] == -1)
    db.close()


def test_search_ringkey_exact_knn(small_db, oracle):
    db, sig, key = small_db
    rng = np.random.default_rng(2)
    qk = key[rng.integers(0, len(key), 20)] + rng.normal(0, 0.03, (20, 20)).astype(np.float32)
    cand, dist = db.search_ringkey(qk, k=3, thres=0.1)
    for q in range(20):
        c_o, d_o = oracle.search_ringkey(qk[q], key, k=3, thres=0.1)
        n = len(c_o)
        assert np.array_equal(cand[q, :n], c_o) and np.all(cand[q, n:] == -1)
        assert np.array_equal(dist[q, :n].view(np.uint32), d_o.view(np.uint32))
    cand, dist = db.search_ringkey(qk, k=3, thres=0.1, max_id=500)
    for q in range(20):
        c_o, d_o = oracle.search_ringkey(qk[q], key[:500], k=3, thres=0.1)
        assert np.array_equal(cand[q, :len(c_o)], c_o)


def test_search_sc_reference_semantics(session, oracle):
    """search_sc over explicit candidates, fed in the reference's sparse (index, double) form."""
    rng = np.random.default_rng(9)
    db = api.ScanContextDB(session, 256)
    ptr, idxs, vals, dense = [0], [], [], []
    for r in range(200):
        pts = rng.normal(0, 12, (3000, 3))
        rk, si, sv, _ = oracle.sc_generate(pts)
        sv32 = sv.astype(np.float32).astype(np.float64)  # the database stores fp32
        db.add_sparse(rk, si, sv)
        idxs.append(si)
        vals.append(sv32)
        ptr.append(ptr[-1] + len(si))
        d = np.zeros(1200, np.float32)
        d[si] = sv.astype(np.float32)
        dense.append(d)
    idxs, vals, dense = np.concatenate(idxs), np.concatenate(vals), np.stack(dense)
    for q in range(10):
        cands = rng.choice(200, 3, replace=False).astype(np.int32)
        qd = dense[rng.integers(0, 200)] * 1.0
        qd[rng.integers(0, 1200, 50)] = 0
        qi = np.nonzero(qd)[0].astype(np.int32)
        i_o, d_o = oracle.search_sc(qi, qd[qi].astype(np.float64), ptr, idxs, vals, cands)
        i_g, d_g = db.search_sc(qd, cands)
        assert i_g[0] == i_o and d_g[0] == np.float32(d_o)
    # skipped candidates (-1) and the "first candidate, 1.1" default
    i_g, d_g = db.search_sc(dense[0], np.array([-1, 5, -1], np.int32))
    i_o, d_o = oracle.search_sc_dense(dense[0], dense, np.array([5], np.int32))
    assert i_g[0] == 5 and d_g[0] == np.float32(d_o)
    db.close()


def test_sharded_query_equals_single(session, oracle):
    """Two shards (row i on shard i % 2) combined by min over packed keys == the unsharded answer."""
    sig, key = syn.make_sc_database(2001, 21)
    qs, qk, _ = syn.make_sc_queries(sig, key, 24, 8)
    shards = []
    for r in range(2):
        rows = api.shard_rows(len(sig), 2, r)
        db = api.ScanContextDB(session, 1024)
        db.add(key[rows], sig[rows], global_ids=rows)
        shards.append(db)
    keys = np.minimum(shards[0].query_keys(qs), shards[1].query_keys(qs))
    diff, idx = api.unpack_key(keys)
    idx_o, diff_o = _oracle_query(oracle, qs, sig)
    assert np.array_equal(idx, idx_o) and np.array_equal(diff.view(np.uint32), diff_o.view(np.uint32))
    assert np.array_equal(api.pack_key(diff, idx), keys)
    for db in shards:
        db.close()


def test_full_size_database_properties(session):
    """BASELINE config 4 size (100k descriptors): self-queries return themselves with distance ~0, perturbed revisits
    return their source row, and the result does not depend on the query batch size."""
    n = 100_000
    sig, key = syn.make_sc_database(n, 2024)
    db = api.ScanContextDB(session, n)
    db.add(key, sig)
    rows = np.array([0, 1, 4999, 50_000, 99_999])
    idx, diff = db.query(sig[rows])
    assert np.array_equal(idx, rows) and np.all(np.abs(diff) < 1e-6)
    qs, qk, truth = syn.make_sc_queries(sig, key, 64, 77)
    idx64, diff64 = db.query(qs)
    known = truth >= 0
    assert np.array_equal(idx64[known], truth[known])
    idx1 = np.array([db.query(qs[i:i + 1])[0][0] for i in range(0, 64, 9)])
    assert np.array_equal(idx1, idx64[::9])
    assert 0.0 < db.last_scan_ms() < 1000.0
    # both scan kernels agree bit for bit on the final (index, distance) at full size
    out = {}
    for flavour in ("stream", "tile"):
        db.set_scan_kernel(flavour)
        out[flavour] = db.query(qs)
    db.set_scan_kernel("auto")
    assert np.array_equal(out["stream"][0], out["tile"][0]) and np.array_equal(out["stream"][1].view(np.uint32), out["tile"][1].view(np.uint32))
    assert np.array_equal(out["tile"][0], idx64)
    db.close()


def test_generate_matches_oracle(session, oracle):
    """dslam_sc_generate (ScanContext::generate on the device) vs the oracle restatement.  The device sums the mean / scatter
    matrix in a tree order, so heights agree to ~1e-12 relative; ring keys and the occupancy pattern are identical."""
    rng = np.random.default_rng(11)
    db = api.ScanContextDB(session, 64)
    for trial in range(6):
        n = int(rng.integers(500, 20000))
        pts = rng.normal(0, [4.0, 14.0, 9.0], (n, 3)) @ np.linalg.qr(rng.normal(size=(3, 3)))[0] + rng.normal(0, 3, 3)
        rk_o, si, sv, tfm_o = oracle.sc_generate(pts)
        rk, sig, sig64, tfm = db.generate(pts, append=True)
        dense_o = np.zeros(1200)
        dense_o[si] = sv
        assert np.array_equal(rk, rk_o)
        assert np.array_equal(sig64 != 0, dense_o != 0)
        assert np.allclose(sig64, dense_o, rtol=1e-10, atol=1e-13)
        assert np.array_equal(sig, sig64.astype(np.float32))
        assert np.allclose(tfm, tfm_o, rtol=1e-10, atol=1e-12)
    assert len(db) == 6
    # the appended descriptors are searchable: a jittered revisit of cloud 3 finds row 3
    rng = np.random.default_rng(11)
    clouds = []
    for trial in range(6):
        n = int(rng.integers(500, 20000))
        clouds.append(rng.normal(0, [4.0, 14.0, 9.0], (n, 3)) @ np.linalg.qr(rng.normal(size=(3, 3)))[0] + rng.normal(0, 3, 3))
    rk, sig, _, _ = db.generate(clouds[3] + np.random.default_rng(1).normal(0, 0.05, clouds[3].shape))
    idx, diff = db.query(sig)
    assert idx[0] == 3 and diff[0] < 0.1
    db.close()


def _lidar_cloud(rng, n=5000):
    """'Imitated LiDAR' cloud of SURVEY.md §8d config 2: n points inside a 40 m sphere; every place has its own anisotropy and a few
    dense structures so that ring keys and signatures differ between places."""
    sig = rng.uniform(4.0, 22.0, 3)
    pts = rng.normal(0, sig, (n, 3))
    for _ in range(int(rng.integers(2, 6))):
        k = int(rng.integers(100, 600))
        pts[rng.integers(0, n, k)] = rng.normal(rng.uniform(-25, 25, 3), rng.uniform(0.5, 3.0, 3), (k, 3))
    return pts


def test_loop_closure_chain_malaga_config(session, oracle):
    """BASELINE configs[2], loop-closure half: 2000 keyframes' clouds -> ScanContext::generate on the device (appended device to
    device) -> 50 revisits (database clouds + N(0, 0.2 m) jitter) -> the reference's two stages, search_ringkey (k = 3, squared L2 <
    0.1) then search_sc over its candidates.  Every stage equals the oracle run on the same descriptors: candidates and ring-key
    distances, winner and sector-cosine distance, bit for bit; and the brute-force scan (no ring-key stage) finds the revisits."""
    rng = np.random.default_rng(7)
    n_db, n_q = 2000, 50
    db = api.ScanContextDB(session, n_db)
    clouds, keys, sigs = [], [], []
    for i in range(n_db):
        pts = _lidar_cloud(rng)
        rk, sig, _, _ = db.generate(pts, append=True)
        clouds.append(pts if i < n_q * 7 else None)
        keys.append(rk)
        sigs.append(sig)
    keys, sigs = np.stack(keys), np.stack(sigs)
    assert len(db) == n_db
    rows = rng.choice(n_q * 7, n_q, replace=False)
    q_keys, q_sigs = [], []
    for r in rows:
        rk, sig, _, _ = db.generate(clouds[r] + rng.normal(0, 0.2, clouds[r].shape))
        q_keys.append(rk)
        q_sigs.append(sig)
    q_keys, q_sigs = np.stack(q_keys), np.stack(q_sigs)
    assert len(db) == n_db  # generate without append leaves the database alone
    cand, cdist = db.search_ringkey(q_keys, k=3, thres=0.1)
    found_two_stage = 0
    for q in range(n_q):
        c_o, d_o = oracle.search_ringkey(q_keys[q], keys, k=3, thres=0.1)
        got = cand[q][cand[q] >= 0]
        assert np.array_equal(got, c_o) and np.array_equal(cdist[q, :len(c_o)].view(np.uint32), d_o.view(np.uint32))
        if len(c_o) == 0:
            continue
        i_g, d_g = db.search_sc(q_sigs[q], cand[q])
        i_o, dd_o = oracle.search_sc_dense(q_sigs[q], sigs, c_o)
        assert i_g[0] == i_o and d_g[0] == np.float32(dd_o)
        found_two_stage += int(i_g[0] == rows[q])
    # exhaustive scan of the same queries (both scan kernels): bit-exact against the oracle, and it finds the places
    idx_o, diff_o = _oracle_query(oracle, q_sigs, sigs)
    for flavour in ("stream", "tile"):
        db.set_scan_kernel(flavour)
        idx, diff = db.query(q_sigs)
        assert np.array_equal(idx, idx_o) and np.array_equal(diff.view(np.uint32), diff_o.view(np.uint32))
    db.set_scan_kernel("auto")
    found_scan = int(np.sum(idx_o == rows))
    print("revisits found: two-stage (FLANN-style k=3 ring-key gate) %d / %d, exhaustive scan %d / %d" % (found_two_stage, n_q, found_scan, n_q))
    assert found_scan >= int(0.8 * n_q)
    db.close()


def test_fp64_database_matches_reference_on_unrounded_signatures(session, oracle, tmp_path):
    """DSLAM_SC_FP64: the signatures keep the reference's double values (SigType = vector<pair<int, double>>, ScanContext.h:24),
    the exact re-score multiplies doubles like search_place.h:71-77, so res_idx / res_diff equal the oracle — and, when
    oracle/_ref is built, the reference's own search_sc compiled in place — on UNROUNDED inputs.  Also: capacity growth past the
    initial allocation and the .scdb save / load round trip."""
    rng = np.random.default_rng(19)
    db = api.ScanContextDB(session, 64, fp64=True)  # grows: 260 rows appended below
    ptr, idxs, vals, dense64, keys = [0], [], [], [], []
    for r in range(260):
        pts = rng.normal(0, 12, (2500, 3))
        rk, si, sv, _ = oracle.sc_generate(pts)
        db.add_sparse(rk, si, sv)
        idxs.append(si)
        vals.append(sv)
        ptr.append(ptr[-1] + len(si))
        d = np.zeros(1200, np.float64)
        d[si] = sv
        dense64.append(d)
        keys.append(rk)
    assert len(db) == 260
    idxs, vals, dense64 = np.concatenate(idxs), np.concatenate(vals), np.stack(dense64)
    ref = None
    try:
        import oracle as orc
        ref = orc.ReferencePieces() if orc.ReferencePieces.available() else None
    except Exception:
        ref = None
    for q in range(12):
        cands = rng.choice(260, 3, replace=False).astype(np.int32)
        qd = dense64[rng.integers(0, 260)] * (1.0 + rng.normal(0, 1e-3, 1200))  # doubles that are NOT fp32-representable
        qd[rng.integers(0, 1200, 50)] = 0
        qi = np.nonzero(qd)[0].astype(np.int32)
        i_o, d_o = oracle.search_sc(qi, qd[qi], ptr, idxs, vals, cands)
        i_g, d_g = db.search_sc64(qd, cands)
        assert i_g[0] == i_o and d_g[0] == np.float32(d_o)
        if ref is not None:
            i_r, d_r = ref.search_sc(qi, qd[qi], ptr, idxs, vals, cands)
            assert i_g[0] == i_r and d_g[0] == np.float32(d_r)
        # full scan: fp32 ranking, fp64 exact re-score of the survivors == brute force over all rows in the reference's arithmetic
        i_all, d_all = oracle.search_sc(qi, qd[qi], ptr, idxs, vals, np.arange(260, dtype=np.int32))
        i_q, d_q = db.query64(qd)
        assert i_q[0] == i_all and d_q[0] == np.float32(d_all)
    # save / load round trip (fp64 table included)
    path = tmp_path / "shard.scdb"
    db.save(path)
    db2 = api.ScanContextDB(session, 16, fp64=True)
    assert db2.load(path) == 260 and len(db2) == 260
    qd = dense64[7] * (1.0 + rng.normal(0, 1e-3, 1200))
    assert db2.query64(qd)[0][0] == db.query64(qd)[0][0] == 7
    assert np.array_equal(db2.query64(qd)[1].view(np.uint32), db.query64(qd)[1].view(np.uint32))
    raw = np.fromfile(path, np.uint8)
    assert raw[:4].tobytes() == b"SCDB" and len(raw) == 64 + 260 * (4 + 20 * 4 + 1200 * 4 + 1200 * 8)
    # an fp32 database reads the same file (ignores the fp64 table) and finds the same row
    db3 = api.ScanContextDB(session, 16)
    assert db3.load(path) == 260
    assert db3.query(qd.astype(np.float32))[0][0] == 7
    for d in (db, db2, db3):
        d.close()


def test_ringkey_width_not_multiple_of_four(session, oracle):
    """A 60x10 descriptor: ring-key rows are only 8-byte aligned — the kNN kernel must not use 16-byte loads there."""
    rng = np.random.default_rng(4)
    db = api.ScanContextDB(session, 128, n_sectors=60, n_rings=10)
    key = rng.uniform(0, 1, (100, 10)).astype(np.float32)
    sig = rng.normal(0, 1, (100, 600)).astype(np.float32)
    db.add(key, sig)
    qk = key[:5] + rng.normal(0, 0.01, (5, 10)).astype(np.float32)
    cand, dist = db.search_ringkey(qk, k=3, thres=0.5)
    for q in range(5):
        c_o, d_o = oracle.search_ringkey(qk[q], key, k=3, thres=0.5)
        assert np.array_equal(cand[q, :len(c_o)], c_o)
        assert np.array_equal(dist[q, :len(c_o)].view(np.uint32), d_o.view(np.uint32))
    idx, _ = db.query(sig[:4])
    assert np.array_equal(idx, [0, 1, 2, 3])
    db.close()


def test_three_scan_kernels_agree_at_full_size(session, oracle):
    """100k descriptors (BASELINE config 4): the HBM-streaming kernel, the FFMA tile kernel and the tcgen05 tensor-core kernel
    (3xTF32 split products, TMEM accumulators) hand the same survivors to the exact re-score: identical argmin and distance bits,
    with and without the ring-key gate / id limit, for a full chunk, a ragged batch and several chunks — the tensor-core kernel in
    all three of its queries-per-pass variants (32: nq 11 / 32, 64: nq 50, 128: nq 70 and the first pass of 150) — spot-checked
    against the oracle's brute force."""
    n = 100_000
    sig, key = syn.make_sc_database(n, 2024)
    db = api.ScanContextDB(session, n)
    db.add(key, sig)
    try:
        for nq in (32, 11, 50, 70, 150):
            qs, qk, truth = syn.make_sc_queries(sig, key, nq, 70 + nq)
            out = {}
            for flavour in ("stream", "tile", "umma"):
                db.set_scan_kernel(flavour)
                out[flavour] = db.query(qs) + db.query(qs, ringkeys=qk, ringkey_thres=0.5, max_id=n // 2)
            for flavour in ("tile", "umma"):
                for a, b in zip(out["stream"], out[flavour]):
                    assert np.array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32)), (flavour, nq)
            known = truth >= 0
            assert np.array_equal(out["umma"][0][known], truth[known])
            for q in range(0, nq, 9):
                i_o, d_o = oracle.search_sc_dense(qs[q], sig)
                assert out["umma"][0][q] == i_o and out["umma"][1][q] == np.float32(d_o)
    finally:
        db.set_scan_kernel("auto")
        db.close()


def test_tensor_core_split_is_exact(session, oracle):
    """3xTF32 on values chosen against the split.  Nine decoy rows hold 1 + 2^-11 + 2^-12 in every cell, the true best row holds
    1 + 2^-10 (a tf32 number), the query is all ones.  With an exact hi / lo split the decoys score 1200.88 and the best row
    1201.17.  If the tensor cores ROUNDED the raw fp32 operand to tf32 instead of dropping its low 13 mantissa bits, the "umma"
    flavour (which feeds the landed fp32 box as the hi operand) would see hi = 1 + 2^-10 for the decoys on top of their lo part:
    all nine would outrank the best row, push it out of the 8 survivors and the exact re-score could not bring it back.  The
    "umma_masked" flavour stores the masked hi explicitly and is the reference for the other."""
    n, cells = 8192, 1200
    sig, key = syn.make_sc_database(n, 99)
    sig = np.ascontiguousarray(sig, np.float32)
    assert sig.shape[1] == cells
    sig[0:9, :] = np.float32(1 + 2.0 ** -11 + 2.0 ** -12)
    sig[9, :] = np.float32(1 + 2.0 ** -10)
    db = api.ScanContextDB(session, n)
    db.add(key, sig)
    qs = np.ascontiguousarray(sig[100:109] * np.float32(0.999), np.float32)
    qs[0, :] = 1.0
    try:
        i_o, d_o = oracle.search_sc_dense(qs[0], sig)
        assert i_o == 9
        for flavour in ("umma_masked", "umma", "tile"):
            db.set_scan_kernel(flavour)
            idx, diff = db.query(qs)
            assert idx[0] == 9 and diff[0] == np.float32(d_o), (flavour, idx[0], diff[0], d_o)
            assert np.array_equal(idx[1:], np.arange(101, 109)), (flavour, idx)
    finally:
        db.set_scan_kernel("auto")
        db.close()

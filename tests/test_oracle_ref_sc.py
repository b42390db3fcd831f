"""Row f-2 pin: the oracle's restatement of ScanContext::generate (oracle/sc_generate.cpp) against the REFERENCE'S OWN
src/loop_closure/loop_detection/ScanContext.cpp compiled in place (oracle/ref_build.py -> oracle/_ref/libdslam_ref_sc.so).
Both sides use the same 3x3 eigen-solver (oracle/jacobi_eig3.h — the reference's Eigen::SelfAdjointEigenSolver is not in this
image), so everything else — centring, covariance, rotation into the PCA frame, polar binning, max-height, ring key, per-sector
L2 normalisation, tfm_pca_rig — must agree BIT FOR BIT."""
import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.skipif(not orc.ReferenceScanContext.available(), reason="oracle/_ref not built (needs /root/reference)")


def _clouds():
    rng = np.random.default_rng(11)
    yield "gaussian", rng.normal(0, 12, (4000, 3)), 40.0
    yield "imitated lidar", rng.normal(0, 1, (5000, 3)) * np.array([25.0, 18.0, 3.0]) + np.array([3.0, -2.0, 1.0]), 40.0
    yield "many out of range", rng.normal(0, 40, (3000, 3)), 30.0
    yield "sparse", rng.uniform(-30, 30, (60, 3)), 40.0
    yield "planar", np.concatenate([rng.uniform(-20, 20, (2000, 2)), rng.normal(0, 0.01, (2000, 1))], 1), 25.0
    yield "tiny sectors", rng.normal(0, 10, (3000, 3)), 40.0


@pytest.mark.parametrize("shape", [(60, 20), (12, 8), (90, 10)])
def test_generate_equals_reference_bit_for_bit(shape):
    o = orc.Oracle()
    ref = orc.ReferenceScanContext()
    ns, nr = shape
    for name, pts, rng_ in _clouds():
        rk_o, i_o, v_o, t_o = o.sc_generate(pts, rng_, ns, nr)
        rk_r, i_r, v_r, t_r = ref.generate(pts, rng_, ns, nr)
        assert np.array_equal(i_o, i_r), name
        assert np.array_equal(v_o.view(np.uint64), v_r.view(np.uint64)), name
        assert np.array_equal(rk_o.view(np.uint32), rk_r.view(np.uint32)), name
        assert np.array_equal(t_o.view(np.uint64), t_r.view(np.uint64)), name
        assert len(i_o) > 0


def test_generated_columns_are_unit_norm():
    """What search_sc relies on (search_place.h:79): every occupied sector column has unit L2 norm, so the sum of products over a
    descriptor with itself is the number of occupied sectors."""
    ref = orc.ReferenceScanContext()
    rng = np.random.default_rng(5)
    rk, idx, val, _ = ref.generate(rng.normal(0, 12, (5000, 3)))
    dense = np.zeros(1200)
    dense[idx] = val
    norms = np.sqrt((dense.reshape(60, 20) ** 2).sum(1))
    occupied = norms > 0
    assert np.allclose(norms[occupied], 1.0, rtol=1e-12)
    assert np.allclose(rk, (dense.reshape(60, 20) != 0).sum(0) / 60.0)

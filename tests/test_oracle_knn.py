"""The nearest-neighbour oracle (oracle/knn_oracle.py, brute force) against scipy's exact k-d tree, and its sentinels."""
import numpy as np

from oracle import knn_oracle as ko


def test_brute_force_matches_kdtree():
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(0)
    pts = rng.uniform(-1, 1, (4000, 3)).astype(np.float32)
    pts[100:140] = pts[:40]  # duplicates
    mean, idx = ko.dist_cuda2(pts)
    dd, _ = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k=4)
    assert np.allclose(mean, (dd[:, 1:] ** 2).mean(1), rtol=1e-4, atol=1e-10)
    assert not np.any(idx == np.arange(len(pts))[:, None])  # a point is never its own neighbour (simple_knn.cu:185-186)
    q = rng.uniform(-3, 3, (500, 3)).astype(np.float32)
    d2, ix = ko.knn(q, pts, 5)
    dd, ii = cKDTree(pts.astype(np.float64)).query(q.astype(np.float64), k=5)
    assert np.allclose(d2, dd ** 2, rtol=1e-4, atol=1e-10)
    assert np.all(np.diff(d2, axis=1) >= 0)


def test_sentinels_and_delete():
    pts = np.zeros((2, 3), np.float32)
    pts[1, 0] = 2.0
    d2, idx = ko.knn(pts, pts, 3, skip_self=True)
    assert d2[0, 0] == 4.0 and idx[0, 0] == 1 and np.all(d2[:, 1:] == ko.FLT_MAX) and np.all(idx[:, 1:] == ko.INT_MAX)
    d2, idx = ko.knn(pts, np.zeros((0, 3), np.float32), 2)
    assert np.all(d2 == ko.FLT_MAX)
    a = np.arange(12).reshape(4, 3)
    (out,) = ko.delete_rows(np.array([True, False, False, True]), [a])
    assert np.array_equal(out, a[1:3])

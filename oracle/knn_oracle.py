"""CPU restatement (numpy) of the nearest-neighbour queries and the mask compaction of SURVEY.md section 8(f) #4.

TEST INFRASTRUCTURE: imported only by tests/ (never by the product path).

* `knn(query, ref, K, skip_self)`: brute-force exact K nearest neighbours with the reference's fp32 distance expression
  d.x*d.x + d.y*d.y + d.z*d.z (submodules/simple-knn/simple_knn.cu:132-145); `skip_self` ignores ref[i] for query i
  (simple_knn.cu:185-186,203-204). Slots without a candidate keep simple-knn's sentinels FLT_MAX / INT_MAX.
* `dist_cuda2(points)`: distCUDA2 = mean of the three smallest squared distances + their indices (simple_knn.cu:213-216).
* `delete_rows(mask, arrays)`: GaussianPointCloud.delete (SLAM/gaussian_pointcloud.py:195-206): rows with ~mask.
The reference search (Morton boxes) and this one are both exact: distances agree to fp32 rounding of the identical
expression; indices agree wherever the K-th and (K+1)-th distances differ.
"""
import numpy as np

FLT_MAX = np.float32(3.402823466e+38)
INT_MAX = np.int32(2**31 - 1)


def knn(query, ref, K, skip_self=False, chunk=2048):
    q = np.asarray(query, np.float32)
    r = np.asarray(ref, np.float32)
    nq, nr = len(q), len(r)
    d2 = np.full((nq, K), FLT_MAX, np.float32)
    idx = np.full((nq, K), INT_MAX, np.int32)
    if nr == 0:
        return d2, idx
    for a in range(0, nq, chunk):
        qq = q[a:a + chunk]
        dx = r[None, :, 0] - qq[:, None, 0]
        dy = r[None, :, 1] - qq[:, None, 1]
        dz = r[None, :, 2] - qq[:, None, 2]
        d = (dx * dx + dy * dy + dz * dz).astype(np.float32)
        if skip_self:
            rows = np.arange(len(qq))
            d[rows, a + rows] = np.inf
        k = min(K, nr - (1 if skip_self else 0))
        if k <= 0:
            continue
        part = np.argpartition(d, k - 1, axis=1)[:, :k]
        pd = np.take_along_axis(d, part, 1)
        order = np.argsort(pd, axis=1, kind="stable")
        d2[a:a + chunk, :k] = np.take_along_axis(pd, order, 1)
        idx[a:a + chunk, :k] = np.take_along_axis(part, order, 1)
    return d2, idx


def dist_cuda2(points):
    d2, idx = knn(points, points, 3, skip_self=True)
    return ((d2[:, 0] + d2[:, 1] + d2[:, 2]) / np.float32(3.0)).astype(np.float32), idx


def delete_rows(mask, arrays):
    keep = ~np.asarray(mask, bool)
    return [np.asarray(a)[keep] for a in arrays]

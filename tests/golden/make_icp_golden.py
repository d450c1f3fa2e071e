"""Generates tests/golden/icp_*.npz by running the UNMODIFIED reference SLAM/icp.py (+ SLAM/utils.py pyramid
helpers) on CPU tensors. Run in the build container, where /root/reference exists:

    python tests/golden/make_icp_golden.py

The reference file imports packages that are not installed here (open3d, plyfile, pytorch3d, skimage) and
`utils.general_utils`, which allocates CUDA tensors at import; they are replaced by inert stub modules so that
the reference *source* is executed as is.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"


class _Stub(types.ModuleType):
    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []
        self.__spec__ = None

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Stub(self.__name__ + "." + k)

    def __call__(self, *a, **k):
        return None


def import_reference_icp():
    sys.path.insert(0, REF)
    for m in ("open3d", "plyfile", "pytorch3d", "pytorch3d.loss", "pytorch3d.ops", "skimage", "skimage.color", "skimage.filters",
              "cv2"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = _Stub(m)
    gu = types.ModuleType("utils.general_utils")
    gu.devF = lambda t: t.float()
    gu.devI = lambda t: t.int()
    gu.devB = lambda t: t.bool()
    gu.inverse_sigmoid = lambda x: torch.log(x / (1 - x))
    gu.build_rotation = None
    gu.quaternion_from_axis_angle = None
    gu.build_covariance_from_scaling_rotation = None
    import utils  # the reference's utils package (namespace)
    sys.modules["utils.general_utils"] = gu
    utils.general_utils = gu
    import SLAM.icp as ricp
    import SLAM.utils as rutils
    return ricp, rutils


def main():
    from rtg_slam_b200 import scene
    ricp, rutils = import_reference_icp()
    torch.manual_seed(2024)
    out_dir = os.path.dirname(os.path.abspath(__file__))
    cases = {
        "icp_small": dict(cam="small", noise=0.002, pose=scene.small_pose()),
        "icp_ragged": dict(cam="ragged", noise=0.0, pose=scene.small_pose((0.6, 0.9, -0.5), (-0.015, 0.01, 0.02))),
    }
    for name, c in cases.items():
        cam0 = scene.make_camera(c["cam"])
        cam1 = scene.make_camera(c["cam"], c2w=c["pose"])
        d0 = scene.raycast_room_depth(cam0, noise_sigma=c["noise"], seed=3)
        d1 = scene.raycast_room_depth(cam1, noise_sigma=c["noise"], seed=4)
        K = torch.tensor(cam0.K)
        builder = ricp.ImagePyramids([2, 1, 0], "max")
        td0, td1 = torch.from_numpy(d0), torch.from_numpy(d1)
        v0 = rutils.build_vertex_pyramid(td0, builder, K)
        n0 = rutils.build_normal_pyramid(v0)
        v1 = rutils.build_vertex_pyramid(td1, builder, K)
        n1 = rutils.build_normal_pyramid(v1)
        pose = torch.eye(4)
        poses, ratios = [], []
        for lvl, s in enumerate([0.25, 0.5, 1.0]):
            Kl = K * s
            Kl[2, 2] = 1.0
            tr = ricp.ICP(5, damping=1e-4, distance_threshold=0.1, normal_threshold=20)
            # predict_pose's call: (pose, vertex_t1, vertex_t0, normal_t1, normal_t0, K) with t1 = current frame
            pose, vr = tr.icp(pose, v1[lvl], v0[lvl], n1[lvl], n0[lvl], Kl)
            poses.append(pose.numpy().copy())
            ratios.append(float(vr))
        p2p = ricp.point2plane_loss(v0[-1], v1[-1] @ pose[:3, :3].T + pose[:3, 3], n0[-1])
        # one residual/Jacobian evaluation at a non-trivial pose on the finest level
        res, J, valid = ricp.ICP.compute_residuals_jacobian(v1[-1], v0[-1], n1[-1], n0[-1], v1[-1][..., -1] > 0, pose, K, 0.1,
                                                           float(np.cos(np.deg2rad(20))))
        # update_last_status depth filling on synthetic "rendered" maps
        g = torch.Generator().manual_seed(7)
        rd = td0.clone()[..., None]
        rd[torch.rand(rd.shape, generator=g) < 0.2] = 0
        rd = rd + 0.02 * (torch.rand(rd.shape, generator=g) < 0.2)
        trk = types.SimpleNamespace(icp_sample_normal_threshold=0.01, icp_sample_distance_threshold=0.01)
        rn = n0[-1] + 0.2 * torch.randn(n0[-1].shape, generator=g) * (torch.rand(n0[-1].shape[:2], generator=g) < 0.3)[..., None]
        filled = rd.clone()
        ricp.IcpTracker.update_last_status(trk, types.SimpleNamespace(get_intrinsic=None), filled, td0[..., None], rn, n0[-1])
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"), depth0=d0, depth1=d1, K=np.array([cam0.fx, cam0.fy, cam0.cx, cam0.cy], np.float32),
            **{f"v0_{i}": v0[i].numpy() for i in range(3)}, **{f"n0_{i}": n0[i].numpy() for i in range(3)},
            **{f"v1_{i}": v1[i].numpy() for i in range(3)}, **{f"n1_{i}": n1[i].numpy() for i in range(3)},
            poses=np.stack(poses), valid_ratios=np.array(ratios, np.float32), p2ploss=np.float32(p2p),
            res=res.numpy().reshape(-1), J=J.numpy().reshape(-1, 6), valid=valid.numpy().reshape(-1),
            fill_render_depth=rd.numpy()[..., 0], fill_render_normal=rn.numpy(), fill_out=filled.numpy()[..., 0])
        print(name, "final pose\n", poses[-1], "\nvalid", ratios, "p2p", float(p2p))


if __name__ == "__main__":
    main()

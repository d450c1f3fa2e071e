"""Generates tests/golden/icp_sequence_*.npz: the trajectory of the UNMODIFIED reference tracker front-end
(SLAM/icp.py `IcpTracker`: update_curr_status -> predict_pose -> update_last_status / move_last_status, driven the way
SLAM/multiprocess/tracker.py:265-290 drives it) over a synthetic TUM-like RGB-D sequence, on CPU tensors. Run in the
build container, where /root/reference exists:

    python tests/golden/make_icp_sequence_golden.py [name ...]

The reference file is executed as is (stub modules for the packages that are not installed, see make_icp_golden.py).
`IcpTracker.predict_pose` hard-codes `.cuda()` on the pose (icp.py:443); on this GPU-less host `torch.Tensor.cuda` is
patched to the identity for the duration of the run -- no reference source is modified.

Sequence (`sequence_inputs`, shared with the tests so that only poses are stored): the camera moves through the box
room of `scene.raycast_room_depth` along a smooth path (about 2 cm and 1 degree per frame); the measured depth of frame k
is the ray cast from the ground-truth pose plus 2 mm noise; the "model" depth the tracker is given for frame k
(`update_last_status`, icp_use_model_depth = True as in configs/{tum,replica,ours}_base.yaml) is the noise-free ray cast
from the same pose with holes and a band of 3 cm outliers, i.e. what a rendered map with missing and wrong surfels
looks like; its normals come from the reference's own normal-map builder.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

SEQUENCES = {
    # thresholds of configs/base.yaml:92-102; TUM fr1 intrinsics at 640x480 / a ragged size for the edge handling
    "icp_sequence_tum": dict(cam="tum", frames=50, use_model_depth=True, warmup=0),
    "icp_sequence_ragged": dict(cam="ragged", frames=12, use_model_depth=True, warmup=3),
}


def tracker_args(use_model_depth=True, warmup=0):
    return types.SimpleNamespace(icp_downscales=[0.25, 0.5, 1.0], icp_warmup_frames=warmup, icp_use_model_depth=use_model_depth,
                                 icp_downscale_iters=[5, 5, 5], icp_distance_threshold=0.1, icp_normal_threshold=20, icp_damping=1e-4,
                                 verbose=False, icp_sample_distance_threshold=0.01, icp_sample_normal_threshold=0.01,
                                 icp_fail_threshold=0.02)


def gt_pose(k):
    """Camera-to-world pose of frame k: a smooth path, ~2 cm and ~1 degree between consecutive frames."""
    from rtg_slam_b200 import scene
    a = 0.13 * k
    rot = (4.0 * np.sin(a), 6.0 * np.sin(0.7 * a + 0.3), 3.0 * np.sin(1.3 * a))
    trans = (0.25 * np.sin(0.5 * a), 0.10 * np.sin(0.9 * a + 1.0), 0.20 * (1 - np.cos(0.6 * a)))
    return scene.small_pose(rot, trans)


def sequence_inputs(name, k):
    """(measured depth, model depth) of frame k, both (H,W) float32; seeded, so the fixture stores only poses."""
    from rtg_slam_b200 import scene
    cfg = SEQUENCES[name]
    cam = scene.make_camera(cfg["cam"], c2w=gt_pose(k))
    clean = scene.raycast_room_depth(cam)
    depth = scene.raycast_room_depth(cam, noise_sigma=0.002, seed=100 + k)
    rng = np.random.default_rng(5000 + k)
    model = clean.copy()
    model[rng.uniform(size=model.shape) < 0.05] = 0.0                      # unreconstructed pixels
    H, W = model.shape
    x0 = int(rng.integers(0, max(1, W - W // 6)))
    model[:, x0:x0 + W // 8] += np.float32(0.03)                            # a band of wrong surfels
    return depth, model


def main(names):
    from make_icp_golden import import_reference_icp
    from rtg_slam_b200 import scene
    ricp, rutils = import_reference_icp()
    torch.Tensor.cuda = lambda self, *a, **k: self  # icp.py:443 on a GPU-less host
    for name in names:
        cfg = SEQUENCES[name]
        cam = scene.make_camera(cfg["cam"])
        K = torch.tensor(cam.K)
        trk = ricp.IcpTracker(tracker_args(cfg["use_model_depth"], cfg["warmup"]))
        builder = ricp.ImagePyramids([0], "max")
        rel, ok, traj = [], [], [np.eye(4)]
        for k in range(cfg["frames"]):
            depth, model = sequence_inputs(name, k)
            td = torch.from_numpy(depth)
            trk.update_curr_status(td, K)
            if k == 0:
                # tracker.py: the first frame only initialises the state
                trk.move_last_status()
            else:
                pose, success = trk.predict_pose({"K": K, "frame_id": k})
                pose = np.asarray(pose, dtype=np.float64)
                rel.append(pose.astype(np.float32))
                ok.append(bool(success))
                traj.append(traj[-1] @ pose)  # tracker.py:282: pose_es[-1] @ pose
                trk.move_last_status()
            # the mapper's render of this frame replaces the measured depth as the next reference (tracker.py:284-290)
            tm = torch.from_numpy(model)
            rn = rutils.build_normal_pyramid(rutils.build_vertex_pyramid(tm, builder, K))[0]
            fn = trk.normal_pyramid_t1[-1]
            render_depth = tm[..., None].clone()
            trk.update_last_status(types.SimpleNamespace(get_intrinsic=None), render_depth, td[..., None], rn, fn)
            print(name, k, "ok" if (k == 0 or ok[-1]) else "FAIL", flush=True)
        gt = np.stack([np.linalg.inv(gt_pose(0)) @ gt_pose(k) for k in range(cfg["frames"])])
        traj = np.stack(traj)
        ate = np.sqrt(((traj[:, :3, 3] - gt[:, :3, 3]) ** 2).sum(-1).mean())
        print(name, "ATE of the reference tracker vs ground truth: %.2f mm" % (1e3 * ate))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rel_poses=np.stack(rel), success=np.array(ok),
                            trajectory=traj.astype(np.float64), gt=gt, ate_vs_gt=np.float64(ate))


if __name__ == "__main__":
    main(sys.argv[1:] or sorted(SEQUENCES))

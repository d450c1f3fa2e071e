"""The reference's own Python files for the ICP path (SLAM/icp.py + the helpers of SLAM/utils.py it imports), executed
UNMODIFIED as the reference-owned baseline of bench.py (on CUDA tensors, B5: on CPU tensors).

TEST / MEASUREMENT INFRASTRUCTURE, never imported by the product path.

`install()` copies the two files (and the reference's `utils/` package they import) from /root/reference into the
git-ignored `baseline/_ref/` -- no reference source enters the repository's history; the directory travels to the GPU
box with the snapshot, where /root/reference does not exist. `load()` imports `SLAM.icp` / `SLAM.utils` from there with
inert stub modules for the packages the files import at module scope but that this path never calls (open3d, plyfile,
pytorch3d, skimage, cv2), and -- on a host without CUDA -- a four-line stand-in for
`utils.general_utils` (the real one allocates CUDA tensors at import).
"""
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ("SLAM/icp.py", "SLAM/utils.py", "utils/general_utils.py", "utils/graphics_utils.py", "utils/sh_utils.py", "utils/system_utils.py")


def install(force=False):
    """Copy the reference files into baseline/_ref/ (no-op without /root/reference). Returns True if they are in place."""
    if os.path.isdir(REF):
        for f in FILES:
            src, dst = os.path.join(REF, f), os.path.join(DST, f)
            if os.path.exists(src) and (force or not os.path.exists(dst)):
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
    return available()


def available():
    return os.path.exists(os.path.join(DST, "SLAM", "icp.py")) and os.path.exists(os.path.join(DST, "SLAM", "utils.py"))


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


def load():
    """(SLAM.icp module, SLAM.utils module) of the reference, or None if baseline/_ref is not installed."""
    if not available():
        return None
    import torch
    if DST not in sys.path:
        sys.path.insert(0, DST)
    for m in ("open3d", "plyfile", "pytorch3d", "pytorch3d.loss", "pytorch3d.ops", "skimage", "skimage.color", "skimage.filters", "cv2"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = _Stub(m)
    for pkg in ("SLAM", "utils"):  # namespace packages rooted at baseline/_ref
        if pkg in sys.modules and not str(getattr(sys.modules[pkg], "__path__", [""])[0]).startswith(DST):
            del sys.modules[pkg]
    if not torch.cuda.is_available():
        gu = types.ModuleType("utils.general_utils")
        gu.devF = lambda t: t.float()
        gu.devI = lambda t: t.int()
        gu.devB = lambda t: t.bool()
        gu.inverse_sigmoid = lambda x: torch.log(x / (1 - x))
        gu.build_rotation = gu.quaternion_from_axis_angle = gu.build_covariance_from_scaling_rotation = None
        import utils
        sys.modules["utils.general_utils"] = gu
        utils.general_utils = gu
    import SLAM.icp as ricp
    import SLAM.utils as rutils
    return ricp, rutils


if __name__ == "__main__":
    print("installed" if install(force="--force" in sys.argv) else "reference not available")

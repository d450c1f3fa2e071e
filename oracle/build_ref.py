"""Build the UNMODIFIED reference rasterizer (RAST = submodules/diff-gaussian-rasterizer-depth)
for sm_100 and drop only the resulting extension module into oracle/_ref/.

TEST INFRASTRUCTURE. Reference sources are compiled where they lie (a scratch copy under /tmp is
needed because setuptools writes build/ next to setup.py and /root/reference is read-only); no
reference source enters this repository. oracle/_ref/ is git-ignored but travels to the GPU box.

The only deviation from `python setup.py build_ext` is `-include cstdint` (gcc 13 no longer pulls
<cstdint> in transitively; rasterizer_impl.h:24 uses std::uintptr_t) and the explicit arch list.
"""
import glob
import os
import shutil
import subprocess
import sys

REF = "/root/reference/submodules/diff-gaussian-rasterizer-depth"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def main():
    if not os.path.isdir(REF):
        print("reference not present; keeping whatever is in oracle/_ref/")
        return 0
    os.makedirs(OUT, exist_ok=True)
    if glob.glob(os.path.join(OUT, "_C_depth*.so")) and "--force" not in sys.argv:
        print("oracle/_ref already built")
        return 0
    tmp = "/tmp/rtg_refbuild"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(REF, tmp)
    env = dict(os.environ, NVCC_APPEND_FLAGS="-include cstdint", TORCH_CUDA_ARCH_LIST="10.0", MAX_JOBS="8")
    subprocess.check_call([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, env=env)
    so = glob.glob(os.path.join(tmp, "diff_gaussian_rasterization_depth", "_C_depth*.so"))
    assert so, "reference build produced no extension"
    shutil.copy(so[0], OUT)
    print("built", os.path.join(OUT, os.path.basename(so[0])))
    return 0


def build_cuda_utils():
    """submodules/cuda_utils (accumulate_gaussian_error and friends): same recipe, module `_C`
    kept under oracle/_ref/cuda_utils/."""
    ref = "/root/reference/submodules/cuda_utils"
    out = os.path.join(OUT, "cuda_utils")
    if not os.path.isdir(ref):
        return 0
    if glob.glob(os.path.join(out, "_C*.so")) and "--force" not in sys.argv:
        print("oracle/_ref/cuda_utils already built")
        return 0
    os.makedirs(out, exist_ok=True)
    tmp = "/tmp/rtg_refbuild_cuda_utils"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(ref, tmp)
    env = dict(os.environ, NVCC_APPEND_FLAGS="-include cstdint", TORCH_CUDA_ARCH_LIST="10.0", MAX_JOBS="8")
    subprocess.check_call([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, env=env)
    so = glob.glob(os.path.join(tmp, "cuda_utils", "_C*.so"))
    assert so, "cuda_utils build produced no extension"
    shutil.copy(so[0], out)
    print("built", os.path.join(out, os.path.basename(so[0])))
    return 0


def build_simple_knn():
    """submodules/simple-knn (distCUDA2), compiled from its three sources where they lie with the flags of
    its own setup.py, plus <cfloat> / <climits> (gcc 13: FLT_MAX / INT_MAX, simple_knn.cu:90,176-177) and a distinct module
    name: cuda_utils and simple-knn both call their extension `_C`, and two single-phase extension modules of the same
    name cannot be loaded side by side in one test process. Output: oracle/_ref/simple_knn/_C_simple_knn*.so."""
    import sysconfig
    from torch.utils import cpp_extension as ce
    ref = "/root/reference/submodules/simple-knn"
    out = os.path.join(OUT, "simple_knn")
    if not os.path.isdir(ref):
        return 0
    if glob.glob(os.path.join(out, "_C_simple_knn*.so")) and "--force" not in sys.argv:
        print("oracle/_ref/simple_knn already built")
        return 0
    os.makedirs(out, exist_ok=True)
    tmp = "/tmp/rtg_refbuild_simple_knn"
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    name = "_C_simple_knn"
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    defs = [f"-DTORCH_EXTENSION_NAME={name}", "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=1"]
    objs = []
    for src in ("spatial.cu", "simple_knn.cu"):
        o = os.path.join(tmp, src + ".o")
        subprocess.check_call(["nvcc", "-c", os.path.join(ref, src), "-o", o, "-O3", "-std=c++17", "--expt-relaxed-constexpr",
                               "-gencode", "arch=compute_100,code=sm_100", "-Xcompiler", "-fPIC", "-include", "cstdint", "-include",
                               "cfloat", "-include", "climits", *defs, *inc])
        objs.append(o)
    o = os.path.join(tmp, "ext.o")
    subprocess.check_call(["g++", "-c", os.path.join(ref, "ext.cpp"), "-o", o, "-O2", "-std=c++17", "-fPIC", *defs, *inc])
    objs.append(o)
    so = os.path.join(out, name + sysconfig.get_config_var("EXT_SUFFIX"))
    libdirs = [f"-L{p}" for p in ce.library_paths("cuda")]
    subprocess.check_call(["g++", "-shared", "-o", so, *objs, *libdirs, "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda",
                           "-ltorch_cuda", "-lcudart"])
    for old in glob.glob(os.path.join(out, "_C.*.so")):
        os.remove(old)
    print("built", so)
    return 0


if __name__ == "__main__":
    sys.exit(main() or build_cuda_utils() or build_simple_knn())

"""Generates tests/golden/api_surface.json: the public surface of the reference interfaces this repository mirrors, read
from the reference's own source files with `ast` (nothing is imported or executed). Run in the build container, where
/root/reference exists:

    python tests/golden/make_api_surface_golden.py

Recorded per interface: NamedTuple fields (name, annotation, default), method signatures (parameter names and defaults), the
messages of the exceptions `GaussianRasterizer.forward` raises, and the keys of the dictionary `Renderer.render` returns.
tests/test_api_surface_cpu.py compares the mirrors in rtg_slam_b200 / diff_gaussian_rasterization_depth with it."""
import ast
import json
import os

REF = "/root/reference"
RAST = "submodules/diff-gaussian-rasterizer-depth/diff_gaussian_rasterization_depth/__init__.py"
FILES = {"rasterizer": RAST, "render": "SLAM/render.py", "icp": "SLAM/icp.py", "utils": "SLAM/utils.py"}
METHODS = {
    "rasterizer": {"GaussianRasterizer": ["__init__", "markVisible", "forward"]},
    "render": {"Renderer": ["__init__", "render"]},
    "icp": {"ICP": ["__init__", "icp"], "IcpTracker": ["__init__", "update_curr_status", "move_last_status", "update_last_status",
                                                       "predict_pose"]},
}
FUNCTIONS = {"rasterizer": ["rasterize_gaussians"], "icp": ["point2plane_loss"],
             "utils": ["pixelmask2tilemask", "transmission2tilemask", "colorerror2tilemask", "bilateralFilter_torch"]}
CUDA_UTILS_H = "submodules/cuda_utils/cuda_utils.h"   # C++ declaration behind `cuda_utils._C.accumulate_gaussian_error`


def signature(fn):
    a = fn.args
    pos = [x.arg for x in a.posonlyargs + a.args]
    defaults = [None] * (len(pos) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
    return [{"name": n, "default": d} for n, d in zip(pos, defaults)]


def main():
    out = {}
    for key, rel in FILES.items():
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        classes = {n.name: n for n in tree.body if isinstance(n, ast.ClassDef)}
        funcs = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}
        entry = {"file": rel, "methods": {}, "functions": {}}
        for cls, names in METHODS.get(key, {}).items():
            body = {n.name: n for n in classes[cls].body if isinstance(n, ast.FunctionDef)}
            entry["methods"][cls] = {m: signature(body[m]) for m in names}
        for f in FUNCTIONS.get(key, []):
            entry["functions"][f] = signature(funcs[f])
        if key == "rasterizer":
            nt = classes["GaussianRasterizationSettings"]
            entry["settings_fields"] = [{"name": s.target.id, "annotation": ast.unparse(s.annotation),
                                         "default": None if s.value is None else ast.unparse(s.value)}
                                        for s in nt.body if isinstance(s, ast.AnnAssign)]
            fwd = next(n for n in classes["GaussianRasterizer"].body if isinstance(n, ast.FunctionDef) and n.name == "forward")
            entry["forward_exceptions"] = [r.exc.args[0].value for r in ast.walk(fwd)
                                           if isinstance(r, ast.Raise) and isinstance(r.exc, ast.Call) and r.exc.args
                                           and isinstance(r.exc.args[0], ast.Constant)]
            apply_fn = next(n for n in classes["_RasterizeGaussians"].body if isinstance(n, ast.FunctionDef) and n.name == "forward")
            ret = [r for r in ast.walk(apply_fn) if isinstance(r, ast.Return)][-1]
            entry["rasterize_outputs"] = [ast.unparse(e) for e in ret.value.elts]
        if key == "render":
            rend = next(n for n in classes["Renderer"].body if isinstance(n, ast.FunctionDef) and n.name == "render")
            res = next(n for n in ast.walk(rend) if isinstance(n, ast.Assign) and isinstance(n.value, ast.Dict)
                       and any(isinstance(t, ast.Name) and t.id == "results" for t in n.targets))   # `results = {...}; return results`
            entry["render_result_keys"] = [k.value for k in res.value.keys]
            init = next(n for n in classes["Renderer"].body if isinstance(n, ast.FunctionDef) and n.name == "__init__")
            entry["renderer_args_read"] = sorted({n.attr for n in ast.walk(init) if isinstance(n, ast.Attribute)
                                                  and isinstance(n.value, ast.Name) and n.value.id == "args"})
        if key == "icp":
            init = next(n for n in classes["IcpTracker"].body if isinstance(n, ast.FunctionDef) and n.name == "__init__")
            entry["tracker_args_read"] = sorted({n.attr for n in ast.walk(init) if isinstance(n, ast.Attribute)
                                                 and isinstance(n.value, ast.Name) and n.value.id == "args"})
        out[key] = entry
    import re
    decl = re.search(r"accumulate_gaussian_error\s*\((.*?)\)\s*;", open(os.path.join(REF, CUDA_UTILS_H)).read(), re.S).group(1)
    out["cuda_utils"] = {"file": CUDA_UTILS_H, "functions": {"accumulate_gaussian_error": [
        {"name": re.sub(r"[&*]", "", a.strip().split()[-1]), "default": None} for a in decl.split(",")]}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_surface.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)
    print(json.dumps(out, indent=1)[:3000])


if __name__ == "__main__":
    main()

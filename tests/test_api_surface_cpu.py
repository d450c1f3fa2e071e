"""The operator API this repository mirrors, checked against the surface of the reference's own source files
(tests/golden/api_surface.json: read with `ast` from RAST/.../__init__.py, SLAM/render.py and SLAM/icp.py by
tests/golden/make_api_surface_golden.py): field names, annotations and defaults of GaussianRasterizationSettings, the
parameter names and defaults of every mirrored method, the exception messages of GaussianRasterizer.forward, the keys of
Renderer.render's result and the `args` attributes the constructors read. No CUDA call is made."""
import ast
import inspect
import json
import os
import types

import pytest
import torch

SURFACE = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "api_surface.json")))


def _params(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        assert p.kind in (p.POSITIONAL_OR_KEYWORD,), (fn, p)
        out.append((p.name, None if p.default is p.empty else p.default))
    return out


def _expect(sig):
    return [(a["name"], None if a["default"] is None else ast.literal_eval(a["default"])) for a in sig]


def test_settings_fields_match_the_reference():
    from diff_gaussian_rasterization_depth import GaussianRasterizationSettings as S
    want = SURFACE["rasterizer"]["settings_fields"]
    assert list(S._fields) == [f["name"] for f in want]
    assert {k: v for k, v in S._field_defaults.items()} == {f["name"]: ast.literal_eval(f["default"]) for f in want if f["default"] is not None}
    def ann_name(v):   # this module uses postponed annotations: NamedTuple keeps them as ForwardRef('int') / ForwardRef('torch.Tensor')
        return getattr(v, "__forward_arg__", None) or (v if isinstance(v, str) else getattr(v, "__name__", str(v)))
    ann = {k: ann_name(v) for k, v in S.__annotations__.items()}
    for f in want:
        assert ann[f["name"]].replace("torch.", "") == f["annotation"].replace("torch.", ""), f


def test_rasterizer_signatures_and_messages_match_the_reference():
    import diff_gaussian_rasterization_depth as pkg
    from rtg_slam_b200 import rasterizer
    want = SURFACE["rasterizer"]
    for name, sig in want["methods"]["GaussianRasterizer"].items():
        assert _params(getattr(pkg.GaussianRasterizer, name)) == _expect(sig), name
    assert _params(pkg.rasterize_gaussians) == _expect(want["functions"]["rasterize_gaussians"])
    assert len(want["rasterize_outputs"]) == 8      # colour, depth, two index maps, two weights, T map, radii
    # the two exceptions of forward(): same type, same text (callers may match on the reference's typo), raised before any CUDA work
    src = inspect.getsource(rasterizer.GaussianRasterizer.forward)
    for msg in want["forward_exceptions"]:
        assert msg in src, msg
    r = pkg.GaussianRasterizer(raster_settings=None)
    x = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs"):
        r(means3D=x, opacities=x[:, :1], scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="excatly one of either SHs"):
        r(means3D=x, opacities=x[:, :1], shs=torch.zeros(4, 16, 3), colors_precomp=x, scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        r(means3D=x, opacities=x[:, :1], colors_precomp=x, scales=x)
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        r(means3D=x, opacities=x[:, :1], colors_precomp=x, scales=x, rotations=torch.zeros(4, 4), cov3D_precomp=torch.zeros(4, 6))


def test_renderer_surface_matches_the_reference():
    from rtg_slam_b200.render import Renderer
    want = SURFACE["render"]
    for name, sig in want["methods"]["Renderer"].items():
        assert _params(getattr(Renderer, name)) == _expect(sig), name
    src = inspect.getsource(Renderer.render)
    tree = ast.parse("class _:\n" + src if src.startswith("    ") else src)
    keys = [k.value for n in ast.walk(tree) if isinstance(n, ast.Return) and isinstance(n.value, ast.Dict) for k in n.value.keys]
    assert keys[:len(want["render_result_keys"])] == want["render_result_keys"]      # same keys, same order ...
    assert keys[len(want["render_result_keys"]):] == ["radii"]                        # ... plus the documented extra
    init_src = ast.parse("class _:\n" + inspect.getsource(Renderer.__init__))
    read = sorted({n.attr for n in ast.walk(init_src) if isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name) and n.value.id == "args"})
    assert read == want["renderer_args_read"]


def test_icp_surface_matches_the_reference():
    from rtg_slam_b200 import icp
    want = SURFACE["icp"]
    for cls, methods in want["methods"].items():
        for name, sig in methods.items():
            got, exp = _params(getattr(getattr(icp, cls), name)), _expect(sig)
            if (cls, name) == ("ICP", "__init__"):
                # the reference declares further keyword arguments that its own callers never pass; ours must accept the
                # ones it implements with the same names, order and defaults
                assert got == exp[:len(got)] and len(got) >= 5, (got, exp)
            else:
                assert got == exp, (cls, name)
    assert _params(icp.point2plane_loss) == _expect(want["functions"]["point2plane_loss"])
    init_src = ast.parse("class _:\n" + inspect.getsource(icp.IcpTracker.__init__))
    read = {n.attr for n in ast.walk(init_src) if isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name) and n.value.id == "args"}
    assert read == set(want["tracker_args_read"])
    # a tracker is constructible from the reference's argument bag without touching CUDA
    bag = types.SimpleNamespace(**{k: v for k, v in dict(
        icp_downscales=[0.25, 0.5, 1.0], icp_warmup_frames=0, icp_use_model_depth=True, icp_downscale_iters=[5, 5, 5],
        icp_distance_threshold=0.1, icp_normal_threshold=20, icp_damping=1e-4, verbose=False, icp_sample_distance_threshold=0.01,
        icp_sample_normal_threshold=0.01, icp_fail_threshold=0.02).items()})
    assert set(vars(bag)) >= set(want["tracker_args_read"])
    icp.IcpTracker(bag)


# ----------------------------------------------------------------------------- host logic of the rasterizer shim (no CUDA)
class _FakeEvent:
    def __init__(self, done=False):
        self.done, self.waited = done, 0

    def query(self):
        return self.done

    def synchronize(self):
        self.waited += 1
        self.done = True


def test_capacity_bookkeeping_of_the_rasterizer_shim():
    """rasterizer._DeviceState: counters of a forward are read when its scan kernel has finished; the binning capacity grows to
    1.5x the largest (padded) instance count seen; an overflow that nobody waited for raises at the next reap."""
    from rtg_slam_b200 import rasterizer as rz
    st = rz._DeviceState(torch.device("cpu"))
    assert st.mode == "auto" and st.r_hint == 1 << 16
    # counters layout (include/rtg_splat_b200.h): [num_rendered, active tiles, overflow, longest list, padded entries, ...]
    ev = _FakeEvent(done=False)
    p = rz._Pending([100000, 3000, 0, 900, 104000, 0, 0, 0], ev, r_cap=1 << 16)
    st.pending.append(p)
    assert st.read(p, block=False) is False and not p.done and st.pending == [p]        # not arrived yet: nothing is read
    ev.done = True
    assert st.read(p, block=False) is True and p.done and p.num_rendered == 100000 and not p.overflow
    assert st.r_hint == int(104000 * 1.5) + 4096 and st.last == (100000, 3000, 0, 900)
    assert st.pending == [] and len(st.free) == 1 and p.pinned is None                   # buffer recycled only after its event completed
    assert st.read(p, block=False) is True                                              # idempotent
    # a waiting read blocks on the event
    ev2 = _FakeEvent(done=False)
    p2 = rz._Pending([50, 1, 0, 50, 52, 0, 0, 0], ev2, r_cap=st.r_hint)
    assert st.read(p2, block=True) and ev2.waited == 1
    assert st.r_hint == int(104000 * 1.5) + 4096                                         # the hint never shrinks
    # an overflow that passed unnoticed (deferred mode) raises at the next reap, after the capacity has been raised
    ev3 = _FakeEvent(done=True)
    p3 = rz._Pending([900000, 3225, 1, 4105, 905000, 0, 0, 0], ev3, r_cap=st.r_hint)
    st.pending.append(p3)
    with pytest.raises(RuntimeError, match="rendered empty"):
        st.reap()
    assert st.r_hint == int(905000 * 1.5) + 4096 and st.pending == []
    st.reap()                                                                            # nothing outstanding: no error
    with pytest.raises(ValueError):
        rz.set_capacity_checks("sometimes")


def test_grad_buffer_and_visible_rows_context_managers_nest_and_restore():
    from rtg_slam_b200 import rasterizer as rz
    assert rz._GRAD_BUFFERS[0] is None and rz._VISIBLE_ROWS_ONLY[0] is False
    a, b = {"means3D": torch.zeros(4, 3)}, {"means3D": torch.zeros(4, 3)}
    with rz.grad_buffers(a):
        assert rz._GRAD_BUFFERS[0] is a
        with rz.grad_buffers(b), rz.visible_rows_only():
            assert rz._GRAD_BUFFERS[0] is b and rz._VISIBLE_ROWS_ONLY[0] is True
            with rz.visible_rows_only(False):
                assert rz._VISIBLE_ROWS_ONLY[0] is False
            assert rz._VISIBLE_ROWS_ONLY[0] is True
        assert rz._GRAD_BUFFERS[0] is a and rz._VISIBLE_ROWS_ONLY[0] is False
    assert rz._GRAD_BUFFERS[0] is None
    with pytest.raises(RuntimeError):
        with rz.grad_buffers(a):
            raise RuntimeError("inside")
    assert rz._GRAD_BUFFERS[0] is None                                                   # restored on an exception too
    prev = rz.set_grad_record_hook(print)
    assert prev is None and rz.set_grad_record_hook(None) is print


def test_map_statistics_and_frame_helpers_keep_the_reference_signatures():
    """SLAM/utils.py tile-mask builders and bilateralFilter_torch; the C++ declaration of cuda_utils' accumulate_gaussian_error
    (positional in the reference's only call, mapper.py:546-559: same order, same names here)."""
    from rtg_slam_b200 import frameprep, mapstats
    want = SURFACE["utils"]["functions"]
    for name in ("pixelmask2tilemask", "transmission2tilemask", "colorerror2tilemask"):
        assert _params(getattr(mapstats, name)) == _expect(want[name]), name
    assert _params(frameprep.bilateralFilter_torch) == _expect(want["bilateralFilter_torch"])
    assert [n for n, _ in _params(mapstats.accumulate_gaussian_error)] == \
        [a["name"] for a in SURFACE["cuda_utils"]["functions"]["accumulate_gaussian_error"]]

"""Times rtg_map_adam_step alone: 1 M Gaussians, 40 % of the rows with a gradient, half of the rows attached."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rtg_slam_b200.mapoptim import MapOptimizer

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda", 0)
torch.manual_seed(0)
raw = dict(xyz=torch.randn(P, 3, device=dev), dc=torch.randn(P, 1, 3, device=dev), rest=torch.randn(P, 15, 3, device=dev),
           op=torch.randn(P, 1, device=dev), sc=torch.randn(P, 3, device=dev) - 3, rot=torch.randn(P, 4, device=dev))
lrs = dict(xyz=1e-6, f_dc=1e-6, f_rest=1e-6, opacity=0.0, scaling=1e-6, rotation=1e-6)
mo = MapOptimizer(raw["xyz"], raw["dc"], raw["rest"], raw["op"], raw["sc"], raw["rot"], lrs, confidence=torch.zeros(P, device=dev))
init = {"opacity": raw["op"].clone(), "scaling": raw["sc"].clone(), "xyz": raw["xyz"].clone(), "rotation_raw": raw["rot"].clone()}
init["opacity"][::2] = 0.0
mo.set_attach(init)
radii = (torch.rand(P, device=dev) < 0.4).to(torch.int32)
grads = {k: torch.randn_like(getattr(mo, k)) for k in ("xyz", "shs", "opacity", "scales", "rotations")}


def step():
    for k, g in grads.items():
        getattr(mo, k).grad = g
    mo.step(radii=radii, zero_grad=False)


for _ in range(5):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    step()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 50
from rtg_slam_b200 import _lib
_lib.profile_enable(True)
_lib.profile_read(reset=True)
for _ in range(30):
    step()
prof = _lib.profile_read(reset=True)
_lib.profile_enable(False)
kms = prof["adam"][0] / max(1, prof["adam"][1])
nbytes = P * (59 * 24 + 59 * 4 * 0.4 + 32 + 12 + 4 + 41)
print(f"map_adam_step: call {ms:.4f} ms, kernel {kms:.4f} ms = {nbytes / kms / 1e6:.0f} GB/s algorithmic ({os.environ.get('RTG_SPLAT_LIB', 'in-tree')})")

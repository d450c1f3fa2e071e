# Full single-GPU validation on the B200 box: GPU tests, smoke, default bench line (with extras) into gpurun_out/.
tag=${1:-r02}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warnings.warn" | tail -40
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"])
print({k:round(v["ms"],4) for k,v in d["roofline"]["per_kernel"].items()})
print("parity", d.get("parity"))
print("icp", {k:v for k,v in (d.get("icp") or {}).items() if k in ("value","ms_per_predict_pose","device_ms_per_predict_pose","vs_reference_cuda","vs_reference_cpu","pose_diff_vs_reference","reference_cuda","reference_cpu")})
print("optimize_step", d.get("optimize_step")); print("map_optimize_step", d.get("map_optimize_step"))
ex=d.get("extras",{})
print("ref_cuda", ex.get("reference_cuda")); print("sharded", ex.get("gaussian_sharded_4M")); print("cpu", d.get("cpu_baseline"))
PY

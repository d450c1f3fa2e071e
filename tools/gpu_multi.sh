# multi-GPU bench run: tools/gpu_multi.sh N tag
N=${1:-2}; tag=${2:-r02}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${tag}_${N}gpu.json 2> gpurun_out/bench_${tag}_${N}gpu.err; echo "rc=$?"; tail -5 gpurun_out/bench_${tag}_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}_${N}gpu.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"])
print(d["workload_stats"]["parallelism"][:120])
for k,v in (d.get("extras") or {}).items(): print(k, {a:b for a,b in v.items() if a!="note"})
PY

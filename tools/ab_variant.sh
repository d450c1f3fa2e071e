# A/B helper for gpurun: raster parity tests on each variant, then bench.py alternating base / variant twice.
run() { python bench.py --steps 100 --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))
print({k:round(v['ms'],4) for k,v in d['roofline']['per_kernel'].items()})"; }
for v in "$@"; do RTG_SPLAT_LIB=$PWD/rtg_slam_b200/variants/$v.so timeout 600 python -m pytest tests/test_raster_gpu.py -m gpu -x -q 2>&1 | tail -4; done
for rep in 1 2; do
run base
for v in "$@"; do RTG_SPLAT_LIB=$PWD/rtg_slam_b200/variants/$v.so run $v; done
done

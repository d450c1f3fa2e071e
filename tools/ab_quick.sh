run() { python bench.py --steps 60 --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['ms_per_step'],4), {k:round(v['ms'],4) for k,v in d['roofline']['per_kernel'].items() if 'preprocess' in k or 'zero' in k})"; }
run base
for v in "$@"; do RTG_SPLAT_LIB=$PWD/rtg_slam_b200/variants/$v.so run $v; done

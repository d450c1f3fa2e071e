# ncu captures for profiles/: launch list of the bench command + one full-set capture of each splat kernel.
tag=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 6 --warmup 3 --no-extras > gpurun_out/launches_$tag.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"^(preprocess_fwd|tile_scan|scatter|tile_sort|render_fwd|bwd_zero|render_bwd|preprocess_bwd)_kernel" -s 16 -c 8 -f -o gpurun_out/prof_$tag \
    python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/prof_$tag.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/prof_$tag.ncu-rep

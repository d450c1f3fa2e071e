"""Regenerate profiles/rNN_launches_summary.txt, rNN_ncu_summary.txt and traffic.json.

Usage (after a gpurun capture with tools/gpu_profile.sh):
    cp gpurun_out/launches_r02.csv profiles/r02_launches.csv
    ncu -i gpurun_out/prof_r02.ncu-rep --page raw --csv > /tmp/prof_r02.csv
    python tools/make_profile_summaries.py r02
"""
import csv, json, os, sys
from collections import defaultdict
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
# launch list summary
rows=[r for r in csv.reader(open('/root/repo/profiles/%s_launches.csv' % TAG)) if len(r)>5]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
tot=defaultdict(float); cnt=defaultdict(int)
for r in rows[1:]:
    try: v=float(r[iv].replace(',',''))
    except: continue
    k=r[ik].split('(')[0][:70]; tot[k]+=v; cnt[k]+=1
s=sum(tot.values())
out=["# ncu launch list (gpu__time_duration.sum, --clock-control none), bench.py --steps 6 --warmup 3 --no-extras, 200 launches after skipping 60",
     "# cold-cache, serialised per-launch times: compare SHARES with bench.py's event-timed per_kernel, not absolutes",
     f"{'total us':>12} {'launches':>8} {'share':>7}  kernel"]
for k,v in sorted(tot.items(), key=lambda x:-x[1]):
    out.append(f"{v/1e3:12.1f} {cnt[k]:8d} {100*v/s:6.1f}%  {k}")
open('/root/repo/profiles/%s_launches_summary.txt' % TAG,'w').write("\n".join(out)+"\n")
print("\n".join(out[:14]))
# ncu full summary
rows = list(csv.reader(open('/tmp/prof_%s.csv' % TAG)))
hdr = rows[0]; units = rows[1]
want = ["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__block_size","launch__grid_size","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__thread_inst_executed_per_inst_executed.ratio","smsp__average_warp_latency_per_inst_issued.ratio","l1tex__t_sector_hit_rate.pct","lts__t_sector_hit_rate.pct","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
lines=["# ncu --set full --clock-control none, one launch of each kernel of librtg_splat_b200.so (bench.py workload: 1 M Gaussians, 1200x680)",""]
traffic={}
def num(x):
    return float(x.replace(',',''))
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name=d["Kernel Name"].split('(')[0].replace('rtg::','')
    lines.append(f"== {name}")
    for w in want:
        if w in d: lines.append(f"   {w:75s} {d[w]:>16s} {units[hdr.index(w)]}")
    ur=units[hdr.index("dram__bytes_read.sum")]; uw=units[hdr.index("dram__bytes_write.sum")]
    mult={"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}
    traffic[name.replace('_kernel','')]=int(num(d["dram__bytes_read.sum"])*mult[ur]+num(d["dram__bytes_write.sum"])*mult[uw])
    lines.append("")
open('/root/repo/profiles/%s_ncu_summary.txt' % TAG,'w').write("\n".join(lines))
json.dump(traffic, open('/root/repo/profiles/traffic.json','w'), indent=1)
print(traffic)

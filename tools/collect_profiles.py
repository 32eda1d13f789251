"""Copy / convert the outputs of tools/measure_all.sh <tag> from gpurun_out/ into profiles/ (run in the build
container, where ncu can read the report):  python tools/collect_profiles.py r02"""
import csv
import io
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__bytes_read.sum.per_second,"
           "dram__bytes_write.sum.per_second,launch__block_size,launch__grid_size,launch__registers_per_thread,"
           "l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,"
           "sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,"
           "smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,"
           "smsp__inst_executed.sum")
for name in ("bench_n1.json", "bench_reference_arm.json", "time_detect_cfgs.txt", "sweep_mnn.jsonl", "time_ratio.jsonl",
             "time_corr.jsonl", "time_disk.jsonl", "ncu_launch_list_bench.csv"):
    a = os.path.join(src, f"{tag}_{name}")
    if os.path.exists(a) and os.path.getsize(a):
        shutil.copy(a, os.path.join(dst, f"{tag}_{name}"))
        print("copied", name)
rep = os.path.join(src, f"{tag}_pipe_host_P8.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", METRICS], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    keep = [i for i, h in enumerate(rows[0]) if h == "Kernel Name" or "__" in h]
    out = [[r[i] for i in keep] for r in rows if len(r) == len(rows[0])]
    for r in out[2:]:
        r[0] = r[0].split("(")[0].replace("void ", "").replace("posfeat::", "")
    with open(os.path.join(dst, f"{tag}_ncu_full_pipeline_host_P8.csv"), "w", newline="") as f:
        csv.writer(f).writerows(out)
    print("wrote ncu summary:", len(out) - 2, "kernels")

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
for name in ("bench_n1.json", "bench_reference_arm.json", "bench_n1_table_form.json", "time_detect_cfgs.txt", "time_nms.txt",
             "diag_mnn.txt", "tc_debug_sweep.txt", "verify_debug_sweep.txt", "sweep_mnn.jsonl", "time_ratio.jsonl", "time_corr.jsonl", "time_disk.jsonl",
             "h2d_ceiling_n1.json", "ncu_launch_list_bench.csv", "train_n1.json", "train_n2.json", "train_n8.json",
             "bench_n2.json", "bench_n4.json", "bench_n8.json", "h2d_ceiling_n2.json", "h2d_ceiling_n8.json"):
    a = os.path.join(src, f"{tag}_{name}")
    if os.path.exists(a) and os.path.getsize(a):
        shutil.copy(a, os.path.join(dst, f"{tag}_{name}"))
        print("copied", name)
import json


def summarise(rep_name, out_name):
    rep = os.path.join(src, f"{tag}_{rep_name}.ncu-rep")
    if not os.path.exists(rep):
        return None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", METRICS], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    keep = [i for i, h in enumerate(rows[0]) if h == "Kernel Name" or "__" in h]
    out = [[r[i] for i in keep] for r in rows if len(r) == len(rows[0])]
    for r in out[2:]:
        r[0] = r[0].split("(")[0].replace("void ", "").replace("posfeat::", "")
    with open(os.path.join(dst, f"{tag}_{out_name}.csv"), "w", newline="") as f:
        csv.writer(f).writerows(out)
    print("wrote ncu summary", out_name, ":", len(out) - 2, "kernels")
    return out


pipe = summarise("pipe_P8", "ncu_full_pipeline_P8")
summarise("corr", "ncu_full_corr_kernels")
summarise("window", "ncu_full_window_line_kernels")
if pipe:
    # DRAM read + write of the dominant kernel per pair (the capture ran 8 pairs per launch): bench.py's roofline.traffic
    hdr, units = pipe[0], pipe[1]
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = {}
    for r in pipe[2:]:
        name = r[0].split("<")[0]
        tot = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        traffic[name] = {"dram_bytes_per_pair": tot / 8.0,
                         "source": f"profiles/{tag}_ncu_full_pipeline_P8.csv: dram__bytes_read.sum + dram__bytes_write.sum of one "
                                   f"8-pair launch ({tot / 1e6:.1f} MB)"}
    with open(os.path.join(dst, "ncu_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print("wrote profiles/ncu_traffic.json:", {k: round(v["dram_bytes_per_pair"] / 1e6, 2) for k, v in traffic.items()})

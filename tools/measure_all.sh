#!/usr/bin/env bash
# One GPU-box call that refreshes the measurements DESIGN.md and profiles/ quote (about 6 GPU-minutes on one B200):
#   gpurun --timeout 1800 -- 'bash tools/measure_all.sh r02'
# Outputs land in gpurun_out/<tag>_*; tools/collect_profiles.py <tag> then copies / converts them into profiles/.
# Every command runs plain first; its ncu pass follows only if that exited 0.  Nothing printed under ncu is a bench value.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2> $OUT/${TAG}_bench_reference_arm.err
POSFEAT_MNN_TABLE=1 python bench.py --steps 5 --warmup 3 --no-cpu --no-eager --no-extras > $OUT/${TAG}_bench_n1_table_form.json 2> $OUT/${TAG}_bench_n1_table_form.err
python tools/time_detect_cfgs.py > $OUT/${TAG}_time_detect_cfgs.txt 2>&1
python tools/time_nms.py > $OUT/${TAG}_time_nms.txt 2>&1
python tools/diag_mnn.py > /dev/null 2> $OUT/${TAG}_diag_mnn.txt
python tools/tc_debug_sweep.py 0 64 128 256 384 > $OUT/${TAG}_tc_debug_sweep.txt 2>&1
python tools/verify_debug_sweep.py > $OUT/${TAG}_verify_debug_sweep.txt 2>&1
python tools/time_ratio.py > $OUT/${TAG}_time_ratio.jsonl 2> $OUT/${TAG}_time_ratio.err
python tools/time_corr.py > $OUT/${TAG}_time_corr.jsonl 2> $OUT/${TAG}_time_corr.err
python tools/time_disk.py > $OUT/${TAG}_time_disk.jsonl 2> $OUT/${TAG}_time_disk.err
python tools/h2d_ceiling.py > $OUT/${TAG}_h2d_ceiling_n1.json 2> /dev/null
# launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
python bench.py --steps 2 --warmup 3 --passes 2 --no-cpu --no-eager --no-extras > $OUT/${TAG}_ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_ncu_launch_list_bench.csv \
    python bench.py --steps 2 --warmup 3 --passes 2 --no-cpu --no-eager --no-extras > $OUT/${TAG}_ncu_launch_list.log 2>&1
# full capture of one device-resident pass of the pair pipeline (8 pairs): every kernel of the hot path
python tools/prof_pipeline.py 8 2 > $OUT/${TAG}_ncu_plain_pipe.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:mnn_tc_kernel|tc_rescore|tc_verify|tc_compact|nms_quad|select_kernel|keypoint_outputs|sample_nhwc' \
    --launch-skip 10 --launch-count 10 -f -o $OUT/${TAG}_pipe_P8 \
    python tools/prof_pipeline.py 8 2 > $OUT/${TAG}_ncu_full_pipe.log 2>&1
# the training-side kernels (C4 shapes)
python tools/prof_corr.py > $OUT/${TAG}_ncu_plain_corr.log 2>&1 &&
ncu --set full --clock-control none -k 'regex:corr|normalize_scale' --launch-skip 7 --launch-count 14 -f -o $OUT/${TAG}_corr \
    python tools/prof_corr.py > $OUT/${TAG}_ncu_full_corr.log 2>&1
python tools/prof_window.py > $OUT/${TAG}_ncu_plain_window.log 2>&1 &&
ncu --set full --clock-control none -k 'regex:window|line_search' --launch-skip 3 --launch-count 3 -f -o $OUT/${TAG}_window \
    python tools/prof_window.py > $OUT/${TAG}_ncu_full_window.log 2>&1
tail -2 $OUT/${TAG}_ncu_full_pipe.log $OUT/${TAG}_ncu_full_corr.log $OUT/${TAG}_ncu_full_window.log
ls -la $OUT | grep ${TAG}_ | awk '{print $5, $9}'

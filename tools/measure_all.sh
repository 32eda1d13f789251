#!/usr/bin/env bash
# One GPU-box call that refreshes every measurement DESIGN.md and profiles/ quote (about 5 GPU-minutes on one B200):
#   gpurun --timeout 1500 -- 'bash tools/measure_all.sh r02'
# Outputs land in gpurun_out/<tag>_*; tools/collect_profiles.py <tag> then copies / converts them into profiles/.
# Nothing printed by a run under ncu is a bench value.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -2 $OUT/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
python bench.py --steps 10 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2> $OUT/${TAG}_bench_reference_arm.err
python tools/time_detect_cfgs.py > $OUT/${TAG}_time_detect_cfgs.txt 2>&1
python tools/sweep_mnn.py > $OUT/${TAG}_sweep_mnn.jsonl 2> $OUT/${TAG}_sweep_mnn.err
python tools/time_ratio.py > $OUT/${TAG}_time_ratio.jsonl 2> $OUT/${TAG}_time_ratio.err
python tools/time_corr.py > $OUT/${TAG}_time_corr.jsonl 2> $OUT/${TAG}_time_corr.err
python tools/time_disk.py > $OUT/${TAG}_time_disk.jsonl 2> $OUT/${TAG}_time_disk.err
# launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_ncu_launch_list_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-eager > $OUT/${TAG}_ncu_launch_list.log 2>&1
# full capture of one host-buffer call of the pair pipeline (8 pairs): every kernel of the hot path incl. the staging kernels
ncu --set full --clock-control none --import-source on --launch-skip 13 --launch-count 14 -f -o $OUT/${TAG}_pipe_host_P8 \
    python tools/prof_pipeline.py 8 2 host > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log

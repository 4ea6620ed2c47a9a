#!/bin/bash
# One GPU visit: parity tests, smoke, a short bench, and the ncu launch list of the bench command.
# usage (through gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --layers gpurun_out/${TAG}_layers.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
# ncu launch list of the same bench command (kernel SHARES of the step; absolute times are cold-cache & serialised)
# (bench.py brackets its timed resident steps with cudaProfilerStart/Stop when CCVPE_NCU_RANGE is set)
CCVPE_NCU_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log | cut -c1-300

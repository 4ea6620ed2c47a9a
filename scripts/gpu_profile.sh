#!/bin/bash
# ncu evidence for round 2 (one GPU): launch list of the default bench command + full captures of the dominant / new kernels.
# usage (through gpurun): bash scripts/gpu_profile.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
COMMON="--no-cpu --no-parity --no-torch-gpu-baseline"
# (1) launch list of the timed resident step of the default bench command (bench.py brackets it with cudaProfilerStart/Stop)
CCVPE_NCU_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 $COMMON > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_bench.log | cut -c1-200
# (2) full capture of the dominant kernel (roofline.traffic): one mid-size igemm launch of the default workload
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm_tcgen05 -s 40 -c 1 -o gpurun_out/${TAG}_prof_igemm \
  python bench.py --steps 1 --warmup 3 --no-cuda-graph $COMMON > gpurun_out/${TAG}_ncu_igemm.log 2>&1
# (3) windowed matching kernel (KITTI workload, level-5 launch: 128x128 map)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:match_tcgen05 -s 28 -c 1 -o gpurun_out/${TAG}_prof_match_kitti \
  python bench.py --workload kitti_b32 --steps 1 --warmup 3 --no-cuda-graph $COMMON > gpurun_out/${TAG}_ncu_match.log 2>&1
# (4) tensor-core weight gradient (training workload)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_tcgen05 -s 60 -c 1 -o gpurun_out/${TAG}_prof_wgrad \
  python bench.py --workload train --steps 1 --warmup 3 --no-cuda-graph $COMMON > gpurun_out/${TAG}_ncu_wgrad.log 2>&1
# (5) launch list of one training step
CCVPE_NCU_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches_train.csv \
  python bench.py --workload train --steps 1 --warmup 3 --no-cuda-graph $COMMON > gpurun_out/${TAG}_ncu_train.log 2>&1
ls -la gpurun_out/${TAG}_* | tail -12

#!/bin/bash
# usage: scripts/gpu_multi.sh <tag> <N> [what...]  -- multi-GPU runs on one box (through gpurun --gpus N)
tag=$1; N=$2; shift; shift
what=${*:-"weak strong train"}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N"
for w in $what; do
  case $w in
    devtest) timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -q -k "follows_its_tensors_device" > gpurun_out/${tag}_devtest.log 2>&1; tail -3 gpurun_out/${tag}_devtest.log;;
    weak) timeout 420 $RUN --steps 10 --warmup 3 > gpurun_out/${tag}_n${N}_weak.json 2> gpurun_out/${tag}_n${N}_weak.err; head -c 300 gpurun_out/${tag}_n${N}_weak.json; echo; tail -2 gpurun_out/${tag}_n${N}_weak.err;;
    strong) timeout 420 $RUN --steps 10 --warmup 3 --scaling strong > gpurun_out/${tag}_n${N}_strong.json 2> gpurun_out/${tag}_n${N}_strong.err; head -c 300 gpurun_out/${tag}_n${N}_strong.json; echo; tail -2 gpurun_out/${tag}_n${N}_strong.err;;
    train) timeout 420 $RUN --workload train --steps 5 --warmup 3 > gpurun_out/${tag}_n${N}_train.json 2> gpurun_out/${tag}_n${N}_train.err; head -c 300 gpurun_out/${tag}_n${N}_train.json; echo; tail -4 gpurun_out/${tag}_n${N}_train.err;;
    train_eager) timeout 420 $RUN --workload train --steps 5 --warmup 3 --no-cuda-graph > gpurun_out/${tag}_n${N}_train_eager.json 2> gpurun_out/${tag}_n${N}_train_eager.err; head -c 300 gpurun_out/${tag}_n${N}_train_eager.json; echo; tail -4 gpurun_out/${tag}_n${N}_train_eager.err;;
  esac
done

#!/bin/bash
# BASELINE config 4: every progressive phase, stable and mid-transition, at the per-GPU batches SURVEY.md 8d suggests,
# on N GPUs of this box (weak scaling: the per-GPU batch is fixed).
# usage (on a B200 box): bash scripts/sweep_phases.sh [N] ["alphas"] > gpurun_out/sweep_N.jsonl
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-1}
ALPHAS=${2:-"1.0 0.5"}
PORT=29540
for res in 16 32 64 128 256 512; do
  if [ $res -le 128 ]; then b=64; else b=16; fi
  for alpha in $ALPHAS; do
    if [ $res -eq 16 ] && [ $alpha != 1.0 ]; then continue; fi
    ARGS="--gpus $N --res $res --alpha $alpha --batch $b --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-profile"
    if [ "$N" -eq 1 ]; then
      timeout -s KILL 200 python bench.py $ARGS 2>/dev/null
    else
      PORT=$((PORT + 1))
      timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
        --master-port $PORT bench.py $ARGS 2>/dev/null | grep '^{'
    fi
  done
done

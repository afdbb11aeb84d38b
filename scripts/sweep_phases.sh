#!/bin/bash
# BASELINE config 4: every progressive phase, stable and mid-transition, at the per-GPU batches SURVEY.md 8d suggests.
# usage (on a B200): bash scripts/sweep_phases.sh > gpurun_out/sweep.jsonl
cd "${GRAFT_REPO_ROOT:-.}"
for res in 16 32 64 128 256 512; do
  if [ $res -le 128 ]; then b=64; else b=16; fi
  for alpha in 1.0 0.5; do
    if [ $res -eq 16 ] && [ $alpha != 1.0 ]; then continue; fi
    timeout -s KILL 200 python bench.py --res $res --alpha $alpha --batch $b --steps 10 --warmup 3 --no-cpu-baseline --no-profile 2>/dev/null
  done
done

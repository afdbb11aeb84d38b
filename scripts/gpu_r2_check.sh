#!/bin/bash
# Round-2 GPU check: parity tests, then the default bench line (1 GPU), the eval workload and the reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -n 30 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 5 --dump-kernels gpurun_out/kernels_r2.csv > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
echo "bench rc=$?"
cut -c1-2500 gpurun_out/bench_r2.json
tail -n 5 gpurun_out/bench_r2.err
python bench.py --workload eval --steps 10 --warmup 3 > gpurun_out/bench_eval_r2.json 2> gpurun_out/bench_eval_r2.err
echo "bench eval rc=$?"
cut -c1-2500 gpurun_out/bench_eval_r2.json
tail -n 5 gpurun_out/bench_eval_r2.err

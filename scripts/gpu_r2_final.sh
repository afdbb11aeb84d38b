#!/bin/bash
# Final record of the round: tests, smoke, default bench, eval bench, reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 30 --warmup 5 --dump-kernels gpurun_out/kernels_r2.csv > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"
python bench.py --workload eval --steps 10 --warmup 3 > gpurun_out/bench_eval_r2.json 2> gpurun_out/bench_eval_r2.err; echo "eval rc=$?"
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err; echo "ref rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_r2.json").read().strip().splitlines()[-1])
r=l["roofline"]
print("train", l["value"], l["ms_per_step"], l["rounds"]["ms_per_step"], "e2e", l["e2e"]["value"], "launches", l["gpu_launches"])
print("roofline step", r["frac"], "family", r["family"]["frac"], r["family"]["isolated_ms_per_step"], "wgrad", r["wgrad_family"]["frac"], r["wgrad_family"]["isolated_ms_per_step"], "best", r["best_launch"]["kernel"], r["best_launch"]["dims"], r["best_launch"]["frac"])
print("tensor", r["tensor_TFLOPs"], r["tensor_frac_of_sustained_peak"], "cpu", l["cpu_baseline"]["value"], "eager", l["gpu_eager_baseline"]["value"], l["clocks"])
e=json.loads(open("gpurun_out/bench_eval_r2.json").read().strip().splitlines()[-1])
print("eval", e["value"], e["e2e"]["value"], [(s["n"], s["images_per_s"], s["e2e_images_per_s"]) for s in e["sweep"]], e["roofline"]["frac"])
print(open("gpurun_out/bench_ref_r2.json").read()[:400])
PY

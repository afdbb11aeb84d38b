set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 500 python -m pytest tests -m gpu -q > gpurun_out/t27.log 2>&1; tail -6 gpurun_out/t27.log
timeout -s KILL 200 python scripts/bench_elementwise.py 2>&1 | grep -E "toim_bwd|fromim_bwd|up2_bwd"
for i in 1 2; do timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-profile 2>/dev/null | cut -c1-160; done
timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --res 16 --alpha 1.0 --batch 64 --no-cpu-baseline --no-profile 2>/dev/null | cut -c1-160

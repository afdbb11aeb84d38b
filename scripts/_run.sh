set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 500 python -m pytest tests -m gpu -q > gpurun_out/t30.log 2>&1; tail -6 gpurun_out/t30.log
for i in 1 2; do timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-profile 2>/dev/null | cut -c1-160; done
timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --res 16 --alpha 1.0 --batch 64 --no-cpu-baseline --no-profile 2>/dev/null | cut -c1-160
timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --res 64 --alpha 0.5 --batch 64 --no-cpu-baseline --no-profile 2>/dev/null | cut -c1-160

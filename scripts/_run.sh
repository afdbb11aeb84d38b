set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 500 python -m pytest tests -m gpu -q > gpurun_out/t28.log 2>&1; tail -6 gpurun_out/t28.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-profile 2>/dev/null | cut -c1-160

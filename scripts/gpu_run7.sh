set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 500 python -m pytest tests -m gpu -q > gpurun_out/t13.log 2>&1; tail -8 gpurun_out/t13.log
timeout -s KILL 120 python scripts/bench_conv.py 16,16,512 32,16,256 16,32,128 16,128,16 > gpurun_out/convdbg7.log 2>&1; cat gpurun_out/convdbg7.log
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-kernels gpurun_out/kernels13.csv > gpurun_out/bench13.json 2> gpurun_out/bench13.err; head -c 600 gpurun_out/bench13.json; tail -5 gpurun_out/bench13.err

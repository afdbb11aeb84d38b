set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 500 python -m pytest tests -m gpu -q > gpurun_out/t17.log 2>&1; tail -8 gpurun_out/t17.log
timeout -s KILL 120 python scripts/bench_conv.py 16,16,512 16,16,256 16,32,128 16,32,64 16,64,32 16,128,16 > gpurun_out/convdbg13.log 2>&1; grep -v "^+" gpurun_out/convdbg13.log | sed -E 's/fwd.*(wgrad [^ ]+ [^ ]+).*/\1/'
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --dump-kernels gpurun_out/kernels19.csv > gpurun_out/bench19.json 2> gpurun_out/bench19.err; head -c 1800 gpurun_out/bench19.json; tail -5 gpurun_out/bench19.err

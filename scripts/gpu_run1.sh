set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/t7.log 2>&1; tail -3 gpurun_out/t7.log
python bench.py --steps 10 --warmup 3 --dump-kernels gpurun_out/kernels7.csv > gpurun_out/bench7.json 2> gpurun_out/bench7.err; tail -c 600 gpurun_out/bench7.json
for d in 0 1 2 3 4 6 7; do echo "DEBUG=$d"; NGAN_CONV_DEBUG=$d timeout 120 python scripts/bench_conv.py 16,16,512 16,32,128; done > gpurun_out/convdbg.log 2>&1
NGAN_CONV_TRACE=1 timeout 120 python scripts/trace_conv.py > gpurun_out/trace.log 2>&1
cat gpurun_out/convdbg.log

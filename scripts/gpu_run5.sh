set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 120 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "linear or adam" > gpurun_out/t11a.log 2>&1; tail -8 gpurun_out/t11a.log
timeout -s KILL 400 python -m pytest tests -m gpu -q > gpurun_out/t11.log 2>&1; tail -8 gpurun_out/t11.log
timeout -s KILL 120 python scripts/bench_conv.py 16,16,512 32,16,256 16,32,128 > gpurun_out/convdbg5.log 2>&1; cat gpurun_out/convdbg5.log
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-kernels gpurun_out/kernels11.csv > gpurun_out/bench11.json 2> gpurun_out/bench11.err; head -c 600 gpurun_out/bench11.json; tail -5 gpurun_out/bench11.err
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --res 64 --alpha 0.5 --batch 64 --no-cpu-baseline --dump-kernels gpurun_out/kernels11_64.csv > gpurun_out/bench11_64.json 2> gpurun_out/bench11_64.err; head -c 600 gpurun_out/bench11_64.json; tail -5 gpurun_out/bench11_64.err

set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 500 python -m pytest tests -m gpu -q > gpurun_out/t14.log 2>&1; tail -8 gpurun_out/t14.log
for t in 8 4 2 1; do echo "TPC=$t"; NGAN_WGRAD_TPC=$t timeout -s KILL 120 python scripts/bench_conv.py 16,16,512 32,16,256 16,32,128 16,32,64 16,64,32 16,128,16; done > gpurun_out/convdbg8.log 2>&1; grep -v "^+" gpurun_out/convdbg8.log
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-kernels gpurun_out/kernels14.csv > gpurun_out/bench14.json 2> gpurun_out/bench14.err; head -c 1200 gpurun_out/bench14.json; tail -5 gpurun_out/bench14.err
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --res 64 --alpha 0.5 --batch 64 --no-cpu-baseline > gpurun_out/bench14_64.json 2> gpurun_out/bench14_64.err; head -c 400 gpurun_out/bench14_64.json; tail -5 gpurun_out/bench14_64.err

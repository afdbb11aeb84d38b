set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q > gpurun_out/t9.log 2>&1; tail -15 gpurun_out/t9.log
python bench.py --steps 10 --warmup 3 --dump-kernels gpurun_out/kernels9.csv > gpurun_out/bench9.json 2> gpurun_out/bench9.err; head -c 900 gpurun_out/bench9.json; tail -5 gpurun_out/bench9.err
python bench.py --steps 10 --warmup 3 --res 64 --alpha 0.5 --batch 64 --no-cpu-baseline --dump-kernels gpurun_out/kernels9_64.csv > gpurun_out/bench9_64.json 2> gpurun_out/bench9_64.err; head -c 900 gpurun_out/bench9_64.json; tail -5 gpurun_out/bench9_64.err
timeout 120 python scripts/bench_conv.py 16,16,512 32,16,256 16,32,128 > gpurun_out/convdbg3.log 2>&1; cat gpurun_out/convdbg3.log

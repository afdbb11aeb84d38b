set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 500 python -m pytest tests -m gpu -q > gpurun_out/t15.log 2>&1; tail -8 gpurun_out/t15.log
timeout -s KILL 120 python scripts/bench_conv.py 16,16,512 16,32,128 16,32,64 16,64,32 16,128,16 > gpurun_out/convdbg9.log 2>&1; grep -v "^+" gpurun_out/convdbg9.log
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-kernels gpurun_out/kernels15.csv > gpurun_out/bench15.json 2> gpurun_out/bench15.err; head -c 1100 gpurun_out/bench15.json; tail -5 gpurun_out/bench15.err
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --res 64 --alpha 0.5 --batch 64 --no-cpu-baseline > gpurun_out/bench15_64.json 2> gpurun_out/bench15_64.err; head -c 400 gpurun_out/bench15_64.json; tail -5 gpurun_out/bench15_64.err

set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 500 python -m pytest tests -m gpu -q > gpurun_out/t20.log 2>&1; tail -8 gpurun_out/t20.log
timeout -s KILL 120 python scripts/bench_conv.py 16,16,512 32,16,256 16,32,128 16,64,32 > gpurun_out/convdbg14.log 2>&1; grep -v "^+" gpurun_out/convdbg14.log
for i in 1 2; do timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-kernels gpurun_out/kernels20.csv > gpurun_out/bench20_$i.json 2> gpurun_out/bench20.err; cut -c1-160 gpurun_out/bench20_$i.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench20_$i.json; grep -o '"roofline": {[^}]*}' gpurun_out/bench20_$i.json; done; tail -3 gpurun_out/bench20.err

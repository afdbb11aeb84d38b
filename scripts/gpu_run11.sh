set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 300 python -m pytest tests/test_step_gpu.py -m gpu -q -k "graph or autograd or losses" > gpurun_out/t16.log 2>&1; tail -4 gpurun_out/t16.log
for c in 148 296 444; do echo "CTAS=$c"; NGAN_WGRAD_CTAS=$c timeout -s KILL 120 python scripts/bench_conv.py 16,16,512 16,16,256 32,16,256 16,32,128; done > gpurun_out/convdbg11.log 2>&1; grep -v "^+" gpurun_out/convdbg11.log | sed -E 's/fwd.*(wgrad [^ ]+ [^ ]+).*/\1/'
for i in 1 2; do timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench17_$i.json 2> gpurun_out/bench17.err; cut -c1-160 gpurun_out/bench17_$i.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench17_$i.json; done

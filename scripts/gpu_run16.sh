set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "conv3x3 or linear" 2>&1 | tail -3
timeout -s KILL 120 python scripts/bench_conv.py 16,16,512 32,16,256 16,32,128 16,128,16 2>&1 | grep -v "^+"
for i in 1 2; do timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-profile 2>/dev/null | cut -c1-160; done

set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/t8.log 2>&1; tail -15 gpurun_out/t8.log
python bench.py --steps 10 --warmup 3 --dump-kernels gpurun_out/kernels8.csv > gpurun_out/bench8.json 2> gpurun_out/bench8.err; tail -c 1500 gpurun_out/bench8.json; tail -5 gpurun_out/bench8.err
for k in 3 1; do for d in 0 1 2 8 16; do echo "NKX=$k DEBUG=$d"; NGAN_FOLD_NKX=$k NGAN_CONV_DEBUG=$d timeout 120 python scripts/bench_conv.py 16,16,512 32,16,256 16,32,128; done; done > gpurun_out/convdbg2.log 2>&1
NGAN_CONV_TRACE=1 timeout 120 python scripts/trace_conv.py > gpurun_out/trace2.log 2>&1
NGAN_FOLD_NKX=1 NGAN_CONV_TRACE=1 timeout 120 python scripts/trace_conv.py > gpurun_out/trace2_nkx1.log 2>&1
grep -v "^+" gpurun_out/convdbg2.log | cut -c1-230

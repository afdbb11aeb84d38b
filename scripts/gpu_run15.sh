set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 500 python -m pytest tests -m gpu -q > gpurun_out/t21.log 2>&1; tail -8 gpurun_out/t21.log
for i in 1 2; do
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench21_pdl_$i.json 2> gpurun_out/bench21.err; cut -c1-160 gpurun_out/bench21_pdl_$i.json; tail -2 gpurun_out/bench21.err
NGAN_NO_PDL=1 timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench21_nopdl_$i.json 2> gpurun_out/bench21.err; cut -c1-160 gpurun_out/bench21_nopdl_$i.json
done
timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --res 64 --alpha 0.5 --batch 64 --no-cpu-baseline --no-profile 2>/dev/null | cut -c1-160
NGAN_NO_PDL=1 timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --res 64 --alpha 0.5 --batch 64 --no-cpu-baseline --no-profile 2>/dev/null | cut -c1-160

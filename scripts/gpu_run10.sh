set -x
cd $GRAFT_REPO_ROOT
for i in 1 2; do
timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench16_fused_$i.json 2> gpurun_out/bench16.err; cut -c1-160 gpurun_out/bench16_fused_$i.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench16_fused_$i.json
NGAN_NO_FUSED_TOIM=1 timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench16_unfused_$i.json 2> gpurun_out/bench16.err; cut -c1-160 gpurun_out/bench16_unfused_$i.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench16_unfused_$i.json
done
nproc; uptime

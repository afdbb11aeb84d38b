set -x
cd $GRAFT_REPO_ROOT
for i in 1 2 3; do timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench18_$i.json 2> gpurun_out/bench18.err; cut -c1-160 gpurun_out/bench18_$i.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench18_$i.json; done
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/plain_launches.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches_r01c.csv
python scripts/prof_conv2.py > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv3x3_fold|wgrad" -s 6 -c 3 -o gpurun_out/prof_conv_r01c python scripts/prof_conv2.py > gpurun_out/ncu_prof.log 2>&1
tail -3 gpurun_out/ncu_prof.log

set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q > gpurun_out/t10.log 2>&1; tail -5 gpurun_out/t10.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench10.json 2> gpurun_out/bench10.err; head -c 600 gpurun_out/bench10.json; tail -5 gpurun_out/bench10.err
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/plain_launches.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches_r01b.csv
python scripts/prof_conv2.py > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv3x3_fold -s 4 -c 2 -o gpurun_out/prof_fold_r01b python scripts/prof_conv2.py > gpurun_out/ncu_prof.log 2>&1
tail -3 gpurun_out/ncu_prof.log

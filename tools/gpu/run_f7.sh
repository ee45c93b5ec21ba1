cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r2f7_b.log 2>&1 || { tail -5 gpurun_out/r2f7_b.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f7_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r2f7_ncu.log 2>&1
tail -2 gpurun_out/r2f7_ncu.log | cut -c1-200
wc -l gpurun_out/r2f7_launches.csv

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:lookback_window2 -c 140 --csv --log-file gpurun_out/r2g4_rot_dram.csv python bench.py --steps 60 --warmup 5 --no-cpu --no-extras > gpurun_out/r2g4_ncu.log 2>&1
tail -1 gpurun_out/r2g4_ncu.log | cut -c1-150; wc -l gpurun_out/r2g4_rot_dram.csv

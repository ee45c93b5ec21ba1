cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "tree_tick or c5_full or lookback_c1 or packed_kernel" > gpurun_out/r2t_pytest.log 2>&1; tail -4 gpurun_out/r2t_pytest.log
for n in 65536 131072 1024; do python tools/gpu_launch_timing.py $n 50 1 recompute auto auto 60 >> gpurun_out/r2t_timing.txt 2>&1; done
cat gpurun_out/r2t_timing.txt

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tree_tick" > gpurun_out/r2e1_pytest.log 2>&1; tail -5 gpurun_out/r2e1_pytest.log
for k in k1p k1e; do
python tools/gpu_launch_timing.py 65536 50 1 recompute $k auto 30
python tools/gpu_launch_timing.py 131072 50 1 recompute $k auto 30
python tools/gpu_launch_timing.py 1048576 50 1 recompute $k auto 20
python tools/gpu_launch_timing.py 16384 20 1 recompute $k auto 30
done 2>&1 | tee gpurun_out/r2e1_timing.txt

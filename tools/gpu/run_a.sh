set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -25 gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2a_bench.err
for k in auto k1p k1; do python tools/gpu_launch_timing.py 1024 20 4096 recompute $k >> gpurun_out/r2a_timing.txt 2>&1; done
python tools/gpu_launch_timing.py 1024 20 4096 rolling auto >> gpurun_out/r2a_timing.txt 2>&1
python tools/gpu_launch_timing.py 65536 50 1 recompute auto >> gpurun_out/r2a_timing.txt 2>&1
python tools/gpu_launch_timing.py 2048 20 2048 recompute auto >> gpurun_out/r2a_timing.txt 2>&1
python tools/gpu_launch_timing.py 512 20 4096 recompute auto >> gpurun_out/r2a_timing.txt 2>&1
cat gpurun_out/r2a_timing.txt
python tools/gpu_wide_bank_check.py > gpurun_out/r2a_wide.txt 2>&1; cat gpurun_out/r2a_wide.txt

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "rolling_multi" > gpurun_out/r2v_pytest.log 2>&1; tail -3 gpurun_out/r2v_pytest.log
for cfg in "1024 20 4096" "2048 20 2048" "512 20 4096" "1024 50 2048"; do python tools/gpu_launch_timing.py $cfg rolling auto >> gpurun_out/r2v_timing.txt 2>&1; done
cat gpurun_out/r2v_timing.txt

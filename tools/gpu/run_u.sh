cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "rolling or montecarlo or c4 or replay or set_bank or nan" > gpurun_out/r2u_pytest.log 2>&1; tail -8 gpurun_out/r2u_pytest.log
for cfg in "1024 20 4096" "2048 20 2048" "512 20 4096" "1024 50 2048"; do python tools/gpu_launch_timing.py $cfg rolling auto >> gpurun_out/r2u_timing.txt 2>&1; done
python tools/gpu_launch_timing.py 1024 20 4096 rolling k1r >> gpurun_out/r2u_timing.txt 2>&1
cat gpurun_out/r2u_timing.txt

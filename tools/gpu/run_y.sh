cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "lookahead or montecarlo or c3 or c4 or warm or smoke" > gpurun_out/r2y_pytest.log 2>&1; tail -6 gpurun_out/r2y_pytest.log
python tools/gpu_mc_tick.py rolling > gpurun_out/r2y_mc.txt 2>&1; cat gpurun_out/r2y_mc.txt
python tools/gpu_c3_timing.py 2>&1 | tail -1

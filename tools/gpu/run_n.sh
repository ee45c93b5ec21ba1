cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "planner or montecarlo or c4" > gpurun_out/r2n_pytest.log 2>&1; tail -6 gpurun_out/r2n_pytest.log
python tools/gpu_planner_timing.py > gpurun_out/r2n_planner.txt 2>&1; cat gpurun_out/r2n_planner.txt

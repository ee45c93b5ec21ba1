cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "lookahead or montecarlo or smoke or warm" > gpurun_out/r2f_pytest.log 2>&1; tail -5 gpurun_out/r2f_pytest.log
python tools/gpu_c3_timing.py > gpurun_out/r2f_c3.txt 2>&1; cat gpurun_out/r2f_c3.txt
python tools/gpu_lookahead_timing.py > gpurun_out/r2f_la.txt 2>&1; head -3 gpurun_out/r2f_la.txt

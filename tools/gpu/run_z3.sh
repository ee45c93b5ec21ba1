cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/gpu_mc_lookahead_timing.py > gpurun_out/r2z3.txt 2>&1; cat gpurun_out/r2z3.txt
python tools/gpu_mc_lookahead_timing.py 16384 >> gpurun_out/r2z3.txt 2>&1; tail -3 gpurun_out/r2z3.txt

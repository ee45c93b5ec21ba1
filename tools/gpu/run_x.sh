cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/gpu_mc_tick.py rolling recompute > gpurun_out/r2x_mc.txt 2>&1; cat gpurun_out/r2x_mc.txt
python -m pytest tests -m gpu -q -k "montecarlo or c4" 2>&1 | tail -3

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/gpu_mc_tick.py rolling > gpurun_out/r2z_mc.txt 2>&1; cat gpurun_out/r2z_mc.txt
python tools/gpu_mc_timeline.py 2>&1 | tail -12

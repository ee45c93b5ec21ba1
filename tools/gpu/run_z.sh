cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/gpu_mc_timeline.py > gpurun_out/r2z_timeline.txt 2>&1; tail -16 gpurun_out/r2z_timeline.txt

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/gpu_planner_timing.py > gpurun_out/r2m_planner.txt 2>&1; cat gpurun_out/r2m_planner.txt
python tools/gpu_planner_timing.py 512 >> gpurun_out/r2m_planner.txt 2>&1; tail -4 gpurun_out/r2m_planner.txt

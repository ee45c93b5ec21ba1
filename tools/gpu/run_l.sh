cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/gpu_profile_target.py planner > gpurun_out/r2q_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:planner_kernel -s 3 -c 1 -f -o gpurun_out/r2q_planner python tools/gpu_profile_target.py planner > gpurun_out/r2q_ncu.log 2>&1
tail -n 3 gpurun_out/r2q_plain.log gpurun_out/r2q_ncu.log

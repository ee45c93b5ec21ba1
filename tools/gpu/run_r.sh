cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/gpu_profile_target.py c2 c3 c4 planner > gpurun_out/r2r_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'lookback_window2_kernel' -s 3 -c 1 -f -o gpurun_out/r2r_k1p_c2 python tools/gpu_profile_target.py c2 > gpurun_out/r2r_ncu1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'lookahead2_kernel' -s 3 -c 1 -f -o gpurun_out/r2r_k2p_c3 python tools/gpu_profile_target.py c3 > gpurun_out/r2r_ncu2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'lookback_window2_kernel' -s 2 -c 1 -f -o gpurun_out/r2r_k1p_c4 python tools/gpu_profile_target.py c4 > gpurun_out/r2r_ncu3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'planner_kernel' -s 3 -c 1 -f -o gpurun_out/r2r_planner python tools/gpu_profile_target.py planner > gpurun_out/r2r_ncu4.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep

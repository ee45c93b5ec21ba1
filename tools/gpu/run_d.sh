cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/gpu_profile_target.py c2 c3 > gpurun_out/r2d_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'lookahead2_kernel|lookback_window2_kernel' -s 6 -c 2 -f -o gpurun_out/r2d_prof python tools/gpu_profile_target.py c2 c3 > gpurun_out/r2d_ncu.log 2>&1
tail -5 gpurun_out/r2d_plain.log gpurun_out/r2d_ncu.log
ls -la gpurun_out/

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./tools/ubench/ffma2 > gpurun_out/r2c_ubench.txt 2>&1; cat gpurun_out/r2c_ubench.txt

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./tools/ubench/dfma > gpurun_out/r2o_dfma.txt 2>&1; cat gpurun_out/r2o_dfma.txt

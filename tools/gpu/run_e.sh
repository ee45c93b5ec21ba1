cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in mb3 mb4 mb3_estrin mb4_estrin; do LLAMPC_LIB=$PWD/llampc_b200/libllampc_b200_$v.so python tools/gpu_c3_timing.py >> gpurun_out/r2e_c3.txt 2>&1; done
cat gpurun_out/r2e_c3.txt

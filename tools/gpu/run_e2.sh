cd $GRAFT_REPO_ROOT
export LLAMPC_LIB=$PWD/llampc_b200/libllampc_b200_trace.so
python tools/gpu_k1e_trace.py 65536 50 2>&1 | tee gpurun_out/r2e2_trace.txt
python tools/gpu_k1e_trace.py 131072 50 2>&1 | tee -a gpurun_out/r2e2_trace.txt

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
LLAMPC_LIB=$PWD/llampc_b200/libllampc_b200_trace.so python tools/gpu_k1p_trace.py > gpurun_out/r2q_trace.txt 2>&1
LLAMPC_LIB=$PWD/llampc_b200/libllampc_b200_trace.so python tools/gpu_k1p_trace.py 131072 50 >> gpurun_out/r2q_trace.txt 2>&1
cat gpurun_out/r2q_trace.txt

cd $GRAFT_REPO_ROOT
for c in 200 300 400; do
echo "LLAMPC_EQ_CTAS=$c"
LLAMPC_EQ_CTAS=$c python tools/gpu_launch_timing.py 65536 50 1 recompute k1e auto 30
LLAMPC_EQ_CTAS=$c python tools/gpu_launch_timing.py 131072 50 1 recompute k1e auto 30
LLAMPC_EQ_CTAS=$c python tools/gpu_launch_timing.py 1048576 50 1 recompute k1e auto 20
done 2>&1 | tee gpurun_out/r2e3_timing.txt
export LLAMPC_LIB=$PWD/llampc_b200/libllampc_b200_trace.so
LLAMPC_EQ_CTAS=300 python tools/gpu_k1e_trace.py 65536 50 2>&1 | tee gpurun_out/r2e3_trace.txt

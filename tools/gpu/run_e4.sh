cd $GRAFT_REPO_ROOT
export LLAMPC_LIB=$PWD/llampc_b200/libllampc_b200_trace.so
python tools/gpu_k1e_trace.py 65536 50 2>&1 | tail -6
LLAMPC_EQ_CTAS=200 python tools/gpu_k1e_trace.py 131072 50 2>&1 | tail -6

cd $GRAFT_REPO_ROOT
for s in 0 300 600 900 1500; do
echo "== stagger $s"
LLAMPC_EQ_STAGGER=$s LLAMPC_LIB=$PWD/llampc_b200/libllampc_b200_t512.so python tools/gpu_k1e_trace.py 65536 50 2>&1 | tail -7
done | tee gpurun_out/r2e6_trace.txt

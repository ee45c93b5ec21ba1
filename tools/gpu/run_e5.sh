cd $GRAFT_REPO_ROOT
for v in t512 t256; do
echo "== $v"
LLAMPC_LIB=$PWD/llampc_b200/libllampc_b200_$v.so python tools/gpu_k1e_trace.py 65536 50 2>&1 | tail -8
done | tee gpurun_out/r2e5_trace.txt

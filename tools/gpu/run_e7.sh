cd $GRAFT_REPO_ROOT
for n in 65536 131072 262144 524288 1048576; do
python tools/gpu_launch_timing.py $n 50 1 recompute k1p auto 20
for v in e512 e256; do
LLAMPC_LIB=$PWD/llampc_b200/libllampc_b200_$v.so python tools/gpu_launch_timing.py $n 50 1 recompute k1e auto 20
done
done 2>&1 | tee gpurun_out/r2e7_timing.txt

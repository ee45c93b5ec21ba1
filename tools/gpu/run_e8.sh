cd $GRAFT_REPO_ROOT
for v in r2mb3 r2mb2; do
export LLAMPC_LIB=$PWD/llampc_b200/libllampc_b200_$v.so
echo "== $v"
python tools/gpu_launch_timing.py 65536 50 1 recompute k1p auto 30
python tools/gpu_launch_timing.py 65536 50 1 recompute k1p auto 30 2
python tools/gpu_launch_timing.py 131072 50 1 recompute k1p auto 30
python tools/gpu_launch_timing.py 1048576 50 1 recompute k1p auto 20
done 2>&1 | tee gpurun_out/r2e8_timing.txt
export LLAMPC_LIB=$PWD/llampc_b200/libllampc_b200_r2mb3.so
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "packed or c2_full or splits" 2>&1 | tail -3

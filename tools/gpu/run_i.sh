cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for n in 65536 131072; do for sy in 1 2 4 8; do python tools/gpu_launch_timing.py $n 50 1 recompute k1p auto 40 $sy >> gpurun_out/r2i_split.txt 2>&1; done; done
python tools/gpu_launch_timing.py 65536 50 1 recompute k1 auto 40 2 >> gpurun_out/r2i_split.txt 2>&1
python tools/gpu_launch_timing.py 262144 50 1 recompute k1p auto 40 >> gpurun_out/r2i_split.txt 2>&1
python tools/gpu_launch_timing.py 524288 50 1 recompute k1p auto 40 >> gpurun_out/r2i_split.txt 2>&1
cat gpurun_out/r2i_split.txt

cd $GRAFT_REPO_ROOT
for n in 131072 262144 524288; do
for s in 1 2 4 8; do
python tools/gpu_launch_timing.py $n 50 1 recompute k1p auto 20 $s
done; done 2>&1 | tee gpurun_out/r2f6_split_sweep.txt

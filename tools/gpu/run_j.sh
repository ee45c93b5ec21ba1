cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for sy in 1 2 4; do python tools/gpu_launch_timing.py 75776 50 1 recompute k1p auto 40 $sy >> gpurun_out/r2j_wave.txt 2>&1; done
for sy in 2 4; do python tools/gpu_launch_timing.py 56832 50 1 recompute k1p auto 40 $sy >> gpurun_out/r2j_wave.txt 2>&1; done
for sy in 1 2; do python tools/gpu_launch_timing.py 151552 50 1 recompute k1p auto 40 $sy >> gpurun_out/r2j_wave.txt 2>&1; done
cat gpurun_out/r2j_wave.txt

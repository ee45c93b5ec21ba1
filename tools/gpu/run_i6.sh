cd $GRAFT_REPO_ROOT
timeout 120 python tools/gpu_tick_range_timing.py 2>&1 | tee gpurun_out/r2i6_tick_range.txt
timeout 200 python -m pytest tests -m gpu -q -x -k "low_speed or guard or packed or tree_tick or c2_full" 2>&1 | tail -2

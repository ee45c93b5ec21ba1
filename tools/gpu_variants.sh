#!/bin/bash
# GPU experiment: K1 timing for the default build and the launch-bounds variants (run under gpurun)
for lib in libllampc_b200.so libllampc_b200_mb7.so libllampc_b200_mb8.so; do
  echo "=== $lib"
  LLAMPC_LIB=/root/repo/llampc_b200/$lib python tools/gpu_accuracy.py 2>&1 | grep -E "K1 N|push|C2|1M"
done

cd $GRAFT_REPO_ROOT
timeout 280 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 120 python tools/gpu_tick_range_timing.py 2>&1 | tee gpurun_out/r2i4_tick_range.txt
timeout 200 python bench.py --no-cpu --no-extras > gpurun_out/r2i4_bench.json 2> gpurun_out/r2i4_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2i4_bench.json')); print(d['value'], d['ms_per_step'], d['l2_flushed']['ms_per_step'], d['e2e']['value'], d['tick_latency']['p50_us'])"

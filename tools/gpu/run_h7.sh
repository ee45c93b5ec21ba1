cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tools/gpu_tick_range_timing.py 2>&1 | tee gpurun_out/r2h7_tick_range.txt
python bench.py --no-cpu --no-extras > gpurun_out/r2h7_bench.json 2> gpurun_out/r2h7_bench.err; tail -c 300 gpurun_out/r2h7_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2h7_bench.json')); print(d['value'], d['ms_per_step'], d['l2_flushed']['ms_per_step'], {k:v for k,v in d['e2e'].items() if 'api' not in k})"

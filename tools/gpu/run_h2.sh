cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "replay_pipelined or push" 2>&1 | tail -3
python bench.py --no-cpu --no-extras > gpurun_out/r2h2_bench.json 2> gpurun_out/r2h2_bench.err; tail -c 500 gpurun_out/r2h2_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2h2_bench.json')); print(d['value'], d['ms_per_step']); print({k:v for k,v in d['e2e'].items() if 'api' not in k}); print(d['tick_latency'])"

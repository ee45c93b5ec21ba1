cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "pdl or lookahead" 2>&1 | tail -3
python bench.py --no-cpu > gpurun_out/r2g9_bench.json 2> gpurun_out/r2g9_bench.err; tail -c 300 gpurun_out/r2g9_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2g9_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac']); print(d['other_configs']['C3_lookahead_16384x32x20'])"

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(python -m pytest tests -m gpu -x -q) > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -30 gpurun_out/r2b_pytest.log
python tools/gpu_lookahead_timing.py > gpurun_out/r2b_la.txt 2>&1; cat gpurun_out/r2b_la.txt
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'])
for k,v in d['other_configs'].items(): print(k, v)
PY

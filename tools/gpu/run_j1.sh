cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r2j1_pytest.log 2>&1; tail -2 gpurun_out/r2j1_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/r2j1_bench.json 2> gpurun_out/r2j1_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2j1_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['sync_push_value'], d['tick_latency']['p50_us'], d['tick_latency']['n_over_1ms'])"

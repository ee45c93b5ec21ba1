cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r2i1_pytest.log 2>&1; tail -2 gpurun_out/r2i1_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --impl reference > gpurun_out/r2i1_bench_ref.json 2> gpurun_out/r2i1_bench_ref.err
timeout 400 python bench.py > gpurun_out/r2i1_bench.json 2> gpurun_out/r2i1_bench.err; tail -c 300 gpurun_out/r2i1_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2i1_bench.json')); r=json.load(open('gpurun_out/r2i1_bench_ref.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['sync_push_value'], 'ref', r['value'], 'e2e ratio', d['e2e']['value']/r['value'])
print(d['tick_latency']); print(d['l2_flushed']['ms_per_step'], d['parity'])
for k,v in d['other_configs'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('steps_per_s','ms_per_tick','us_per_tick','ms_per_call','roofline_frac','roofline_frac_tick')})"

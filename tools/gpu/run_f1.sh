cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2f1_pytest.log 2>&1; tail -3 gpurun_out/r2f1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/r2f1_bench_ref.json 2> gpurun_out/r2f1_bench_ref.err; tail -c 300 gpurun_out/r2f1_bench_ref.json
python bench.py > gpurun_out/r2f1_bench.json 2> gpurun_out/r2f1_bench.err; tail -c 400 gpurun_out/r2f1_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2f1_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('register_file'), d['e2e'])"

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "tree_tick or c2_full or c5_full or packed or splits or peer" 2>&1 | tail -2
for p in 0 1; do
LLAMPC_BENCH_PDL=$p python bench.py --no-cpu --no-extras > gpurun_out/r2g6_bench_pdl$p.json 2> gpurun_out/r2g6_bench_pdl$p.err; tail -c 300 gpurun_out/r2g6_bench_pdl$p.err; python -c "
import json; d=json.load(open('gpurun_out/r2g6_bench_pdl$p.json')); print('pdl=$p', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_us'], d['l2_flushed']['ms_per_step'], d['parity']['vs_oracle_f64'])"
done

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py > gpurun_out/r2g3_bench.json 2> gpurun_out/r2g3_bench.err; tail -c 400 gpurun_out/r2g3_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2g3_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_us'], d['roofline']['kernel_us_l2_flushed'], d['l2_flushed'], d['roofline']['register_file']['frac_whole_launch'], d['clocks'], d['config']['l2'], d['other_configs']['C5_1gpu_lookback_1048576x50'])"

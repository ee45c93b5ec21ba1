cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests -m gpu -q -x -k "replay or push or pdl or low_speed" 2>&1 | tail -2
timeout 100 python tools/gpu_replay_length_timing.py 2>&1 | tail -3
timeout 250 python bench.py --no-cpu --no-extras > gpurun_out/r2j2_bench.json 2> gpurun_out/r2j2_bench.err; tail -c 300 gpurun_out/r2j2_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2j2_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['sync_push_value'], d['parity']['e2e_replay_equals_push'])"

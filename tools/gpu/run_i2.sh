cd $GRAFT_REPO_ROOT
for f in ""; do
timeout 400 python bench.py --no-extras $f > gpurun_out/r2i2_bench.json 2> gpurun_out/r2i2_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2i2_bench.json'))
print('$f', d['e2e']['value'], d['e2e']['sync_push_value'], d['tick_latency'])"
done

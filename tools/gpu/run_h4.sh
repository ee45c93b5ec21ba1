cd $GRAFT_REPO_ROOT
python tools/gpu_replay_timing.py 2>&1 | head -7
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 60 --launch-skip 300 --csv --log-file gpurun_out/r2h4_replay_launches.csv python tools/gpu_replay_timing.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2h4_replay_launches.csv')) if len(r)>5]
ix={h:i for i,h in enumerate(rows[0])}
for r in rows[1:40]:
    print(r[ix['ID']], r[ix['Kernel Name']][:60], r[ix['Grid Size']] if 'Grid Size' in ix else '', r[ix['Metric Value']])
PY

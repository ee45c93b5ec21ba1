cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | head -c 160; echo
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"

# the end-of-round verification: every GPU test, the smoke test and the default bench line
mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
(time timeout 900 python bench.py > gpurun_out/final/default_line.json 2> gpurun_out/final/default_line.err) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final/default_line.json').read().strip().splitlines()[-1])
print('default', d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('traffic'), d['verified'], d['clocks'], d['gpu_launches'])
print('cfg2', d['configs1_cfg2']['value'], d['configs1_cfg2']['e2e']['value'], 'float', d['float']['value'], d['float']['e2e']['value'], 'orb', d['orb_extraction']['images_per_s'])
PY

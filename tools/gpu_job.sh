mkdir -p gpurun_out/r2k; (timeout 300 python -m pytest tests -m gpu -x -q --timeout 100 > gpurun_out/r2k/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k/pytest.log); tail -6 gpurun_out/r2k/pytest.log
timeout 400 python bench.py > gpurun_out/r2k/bench_default.json 2> gpurun_out/r2k/bench_default.err; tail -3 gpurun_out/r2k/bench_default.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2k/bench_default.json'))
print('value',l['value'],'e2e',l['e2e']['value'],'dev-only',l['resident_device_only']['value'],'frac',l['roofline']['frac'],l['roofline']['peak'],'verified',l.get('verified'))
print('cfg2',l['configs1_cfg2']['value'],l['configs1_cfg2']['e2e']['value'],l['configs1_cfg2']['roofline']['frac'])
print('float',l['float']['value'],l['float']['e2e']['value'],l['float']['roofline']['frac'])
print('orb',l['orb_extraction'])
print('alt',l['alt_engine']['pairs_per_s_per_gpu'],l['alt_engine']['roofline']['frac'],l['alt_engine']['roofline'].get('executed_popc_frac'))
print('cpu',l['cpu_baseline'])
PY

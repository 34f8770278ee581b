mkdir -p gpurun_out/r2z
timeout 1200 python -m pytest tests/test_parity_binary.py -m gpu -x -q 2>&1 | tail -3
B="--steps 5 --warmup 3 --no-extra --no-alt-engine --no-cpu-baseline --verify 2 --device-only-iters 3"
timeout 300 python bench.py --workload cfg5s $B > gpurun_out/r2z/cfg5s_f4x.json 2> gpurun_out/r2z/cfg5s_f4x.err
SFMM_NO_F4X=1 timeout 300 python bench.py --workload cfg5s $B > gpurun_out/r2z/cfg5s_f4p.json 2> gpurun_out/r2z/cfg5s_f4p.err
SFMM_NO_F4=1 timeout 300 python bench.py --workload cfg5s $B > gpurun_out/r2z/cfg5s_i8.json 2> gpurun_out/r2z/cfg5s_i8.err
timeout 300 python bench.py --workload cfg2 --cross-check $B > gpurun_out/r2z/cfg2_cross.json 2> gpurun_out/r2z/cfg2_cross.err
timeout 300 python bench.py --workload cfg3 --cross-check $B > gpurun_out/r2z/cfg3_cross.json 2> gpurun_out/r2z/cfg3_cross.err
timeout 300 python bench.py --workload orb $B > gpurun_out/r2z/orb.json 2> gpurun_out/r2z/orb.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2z/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['value'], d.get('resident_device_only',{}).get('value'), d['roofline'].get('frac'), d['roofline'].get('tensor_kind'), d.get('verified'))
    except Exception as e: print(f, 'ERR', e)
PY
(time timeout 900 python bench.py > gpurun_out/r2z/default_line.json 2> gpurun_out/r2z/default_line.err) 2>&1 | tail -3
tail -c 600 gpurun_out/r2z/default_line.json

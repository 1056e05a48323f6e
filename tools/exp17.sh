timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_b17.json 2> gpurun_out/r2_b17.err || tail -c 400 gpurun_out/r2_b17.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b17.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['achieved'])
for k in ('large_batch','ddpm1000','train','pipelined'):
    v=d.get(k) or {}
    print(k, v.get('value'), v.get('ms_per_step'), v.get('ms_per_denoise_step'))
print(d['cpu_baseline'])
PY

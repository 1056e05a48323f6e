python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t9.log; tail -3 gpurun_out/r2_t9.log
bash tools/exp2.sh; tail -n 20 gpurun_out/exp_chain_timing.txt
A="--steps 5 --warmup 3 --ddpm-batch 0 --no-train --no-cpu-baseline"
python bench.py $A > gpurun_out/r2_b9.json 2> gpurun_out/r2_b9.err; tail -c 300 gpurun_out/r2_b9.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b9.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('legs',{}).get('large_batch',{}).get('value'))
kc=d['roofline']['kernel_classes_per_denoise_step']
print({k:(v['ms'],v['launches']) for k,v in kc.items() if v['ms']>0})
PY

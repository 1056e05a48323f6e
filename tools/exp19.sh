timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -m gpu -x -q 2>&1 | tail -3
A="--steps 5 --warmup 3 --ddpm-batch 0 --no-train --no-cpu-baseline"
run() {
  env $1 timeout 600 python bench.py $A > gpurun_out/r2_b19.json 2> gpurun_out/r2_b19.err || tail -c 400 gpurun_out/r2_b19.err
  python - "$1" <<'PY'
import json, sys
d=json.loads(open('gpurun_out/r2_b19.json').read().strip().splitlines()[-1])
lb=d['large_batch']
print(sys.argv[1], '| b256', d['value'], '| b4096', lb['value'], lb['ms_per_denoise_step'], 'conv ms', lb['kernel_classes_ms'].get('conv3x3'), lb['kernel_classes_ms'].get('gn_apply'))
PY
}
run "SPDM_NO_PIX256=0"
run "SPDM_NO_PIX256=1"

A="--steps 3 --warmup 3 --ddpm-batch 0 --no-train --no-cpu-baseline --large-batch 0 --batch 4096"
run() {
  env $1 SPDM_PROF_DUMP=1 timeout 600 python bench.py $A > gpurun_out/r2_b16.json 2> gpurun_out/r2_b16.err
  echo "== $1"; grep "spdm prof" gpurun_out/r2_b16.err | tail -70 | awk '{ if ($4 == 0 || $4 == 1 || $4==8 || $4==7) printf "%s:%s:%s ", $3,$4,$5 }'; echo
  python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b16.json').read().strip().splitlines()[-1])
print('   value', d['value'], d['ms_per_step']/50)
PY
}
run "SPDM_FUSE_SHORT=0"
run "SPDM_FUSE_SHORT=1"

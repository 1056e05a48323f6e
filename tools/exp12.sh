A="--steps 5 --warmup 3 --ddpm-batch 0 --no-train --no-cpu-baseline --large-batch 0"
for sp in 1 2 4; do
python bench.py $A --split $sp > gpurun_out/r2_b12_$sp.json 2> gpurun_out/r2_b12.err || tail -c 400 gpurun_out/r2_b12.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_b12_$sp.json').read().strip().splitlines()[-1])
print("split $sp", d['value'], d['ms_per_step'], d['e2e']['value'], d.get('pipelined',{}) if 'pipelined' in d else [ (k,v.get('value')) for k,v in d.get('legs',{}).items()])
PY
done

for c in 1 0; do
SPDM_COMM_PRIORITY=$c python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$c bench.py --gpus 8 --workload train --steps 20 --warmup 5 > gpurun_out/r2_b22_p$c.json 2> gpurun_out/r2_b22_p$c.err || tail -c 300 gpurun_out/r2_b22_p$c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_b22_p$c.json').read().strip().splitlines()[-1])
print("comm priority $c", d['value'], d['ms_per_step'])
PY
done

A="--steps 9 --warmup 3 --large-batch 0 --ddpm-batch 0 --no-train --no-cpu-baseline"
for d in 2 3 4; do python bench.py $A --e2e-depth $d --pipeline-depth $d > gpurun_out/exp_depth_$d.json 2> gpurun_out/exp_depth_$d.err; done

python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t4.log; tail -3 gpurun_out/r2_t4.log
A="--steps 5 --warmup 3 --ddpm-batch 0 --no-train --no-cpu-baseline"
python bench.py $A > gpurun_out/r2_b4.json 2> gpurun_out/r2_b4.err; tail -c 300 gpurun_out/r2_b4.err

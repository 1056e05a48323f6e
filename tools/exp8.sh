python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t5.log; tail -3 gpurun_out/r2_t5.log
python tools/margins.py 2>&1 | tail -3
A="--steps 5 --warmup 3 --ddpm-batch 0 --no-train --no-cpu-baseline"
python bench.py $A > gpurun_out/r2_b5.json 2> gpurun_out/r2_b5.err; tail -c 300 gpurun_out/r2_b5.err

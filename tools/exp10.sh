python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t10.log; tail -3 gpurun_out/r2_t10.log
python tools/train_bench.py 512 10 bf16 attn 2>&1 | tail -2
SPDM_FILM_SIMT=1 python tools/train_bench.py 512 10 bf16 attn 2>&1 | tail -1
python tools/train_bench.py 96 10 bf16 attn 2>&1 | tail -1

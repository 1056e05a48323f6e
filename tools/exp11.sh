python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -2
python tools/train_bench.py 512 10 bf16 attn 2>&1 | tail -1
python tools/train_bench.py 512 10 bf16 attn 2>&1 | tail -1

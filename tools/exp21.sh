nproc
python tools/train_bench.py 512 20 bf16 attn 2>&1 | tail -1
taskset -c 0,1 python tools/train_bench.py 512 20 bf16 attn 2>&1 | tail -1
taskset -c 0 python tools/train_bench.py 512 20 bf16 attn 2>&1 | tail -1

for m in -1 1 2; do
  SPDM_FUSE_MODE=$m python bench.py --batch 4096 --large-batch 0 --no-train --no-cpu-baseline --pipeline-depth 1 --steps 2 --warmup 3 > gpurun_out/exp_fuse_$m.json 2> gpurun_out/exp_fuse_$m.err
done

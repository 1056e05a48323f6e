#!/usr/bin/env python
"""Prints the bf16-vs-fp32 parity margins of the benchmarked configurations (the asserts of tests/test_gpu_bench_configs.py as numbers)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import fixtures
import test_gpu_bench_configs as T

def run(kind, K, B, stride, seedb, seedx, noise_seed=None):
    sd = fixtures.make_unet_weights(attention=True, seed=0)
    esd = fixtures.make_encoder_weights()
    batch = fixtures.make_batch(B, seed=seedb)
    x_T = fixtures.make_xT(B, seed=seedx)
    noise = fixtures.make_noise(K, B, seed=noise_seed) if noise_seed else None
    got, _ = T._sample("bf16", kind, K, B, 10, sd, esd, batch, x_T, noise=None if noise is None else noise.cuda())
    idx = torch.arange(0, B, stride)
    want, _ = T._sample("fp32", kind, K, idx.numel(), 10, sd, esd, T._subset(batch, idx), x_T[idx].contiguous(),
                        noise=None if noise is None else noise[:, idx].contiguous().cuda())
    print("%s-%d B=%d: rel %.3e  per-sample max %.3e" % (kind, K, B, T.rel(got[idx], want), float(T.per_sample_rel(got[idx], want).max())), flush=True)

run("ddim", 50, 256, 1, 1234, 77)
run("ddim", 50, 4096, 7, 4321, 78)
if len(sys.argv) > 1:
    run("ddpm", 1000, 512, 5, 99, 79, 100)

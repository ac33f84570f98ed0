#!/usr/bin/env python
"""First-party radix sort + head compaction (ops/csrc/radix_sort.cu) vs torch.sort / CUB through
the deduplicated-update entry point.  Sizes follow the synthetic models (items per rank and step).

  python tools/bench_sort.py            # standalone kernels vs torch.sort
  DE_B200_SORT=own python tools/bench_sort.py --path   # whole sort_items path with the own sort
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from distributed_embeddings_b200.ops import _native

ops = _native.require()


def timeit(fn, iters=10, warmup=3):
  for _ in range(warmup):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(iters):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / iters * 1e3  # us


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--sizes", default="1703936,4194304,16777216,33554432")
  ap.add_argument("--bits", default="20,28,32")
  ap.add_argument("--alpha", type=float, default=1.05, help="power-law skew of the keys (0: uniform)")
  args = ap.parse_args()
  rows = []
  gen = torch.Generator(device="cuda").manual_seed(1)
  for n in [int(x) for x in args.sizes.split(",")]:
    for bits in [int(x) for x in args.bits.split(",")]:
      hi = 1 << bits
      if args.alpha > 0:
        u = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64)
        # inverse-CDF of a truncated power law on [1, hi]
        a = 1.0 - args.alpha
        keys = (((hi**a - 1.0) * u + 1.0)**(1.0 / a)).long().clamp_(1, hi) - 1
      else:
        keys = torch.randint(0, hi, (n,), generator=gen, device="cuda", dtype=torch.int64)
      items = torch.arange(n, device="cuda", dtype=torch.int32)
      t_own = timeit(lambda: ops.radix_sort_pairs(keys, items, bits))
      t_clone = timeit(lambda: (keys.clone(), items.clone()))  # the standalone op clones its inputs
      t_torch = timeit(lambda: torch.sort(keys, stable=True))
      t_own32, t_clone32 = None, 0.0
      if bits <= 32:
        k32 = torch.where(keys >= 2**31, keys - 2**32, keys).to(torch.int32)
        t_own32 = timeit(lambda: ops.radix_sort_pairs32(k32, items, bits))
        t_clone32 = timeit(lambda: (k32.clone(), items.clone()))
      ks, _ = ops.radix_sort_pairs(keys, items, bits)
      t_heads = timeit(lambda: ops.head_segments(ks))
      t_uniq = timeit(lambda: torch.unique_consecutive(ks, return_counts=True))
      res = {"n": n, "bits": bits, "own_sort_us": round(t_own - t_clone, 1),
             "own_sort32_us": round(t_own32 - t_clone32, 1) if t_own32 is not None else None,
             "torch_sort_us": round(t_torch, 1), "own_heads_us": round(t_heads, 1),
             "torch_unique_consecutive_us": round(t_uniq, 1),
             "own_sort_GBps": round(n * 32.0 * ((bits + 7) // 8) / (t_own - t_clone) / 1e3, 1)}
      rows.append(res)
      print(json.dumps(res), flush=True)
  os.makedirs("gpurun_out", exist_ok=True)
  json.dump(rows, open("gpurun_out/bench_sort.json", "w"), indent=1)


if __name__ == "__main__":
  main()

#!/usr/bin/env python
"""Op micro-benchmark: ragged pooled lookup forward / gradient / SGD vs torch.nn.EmbeddingBag
(reference examples/benchmarks/benchmark.py: voc 1M, dim 128, batch 16384, hotness <= 500),
timed with CUDA events (host clock with --device cpu, for smoke runs)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch

from distributed_embeddings_b200.ops import embedding_lookup_ops as elo
from distributed_embeddings_b200.ops.ragged import RaggedIds


def timeit(fn, iters=20, warmup=5):
  if not torch.cuda.is_available():
    fn()
    t0 = time.perf_counter()
    for _ in range(3):
      fn()
    return (time.perf_counter() - t0) / 3 * 1e3
  for _ in range(warmup):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(iters):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / iters


def main():
  p = argparse.ArgumentParser()
  p.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
  p.add_argument("--voc", type=int, default=1000000)
  p.add_argument("--dim", type=int, default=128)
  p.add_argument("--batch", type=int, default=16384)
  p.add_argument("--max_hot", type=int, default=500)
  args = p.parse_args()
  dev = args.device
  voc, dim, batch, max_hot = args.voc, args.dim, args.batch, args.max_hot
  gen = torch.Generator().manual_seed(0)
  lens = torch.randint(1, max_hot + 1, (batch,), generator=gen)
  vals = torch.randint(0, voc, (int(lens.sum()),), generator=gen)
  ids = RaggedIds.from_row_lengths(vals, lens).to(dev)
  param = torch.rand(voc, dim, device=dev, requires_grad=True)
  bag = torch.nn.EmbeddingBag(voc, dim, mode="sum", sparse=True, device=dev)
  grad = torch.rand(batch, dim, device=dev)
  nnz = int(lens.sum())
  bytes_fwd = nnz * dim * 4 + batch * dim * 4

  t = timeit(lambda: elo.embedding_lookup(param, ids, "sum"))
  print(f"custom fwd        {t:8.3f} ms  {bytes_fwd / t / 1e6:8.1f} GB/s")
  t = timeit(lambda: bag(ids.values, ids.row_splits[:-1]))
  print(f"EmbeddingBag fwd  {t:8.3f} ms  {bytes_fwd / t / 1e6:8.1f} GB/s")

  t = timeit(lambda: elo.lookup_grad_sparse(ids.values, ids.row_splits, 0, batch, 0, grad, voc))
  print(f"custom grad (sort+unique+segment reduce) {t:8.3f} ms")

  def sgd_custom():
    out = elo.embedding_lookup(param, ids, "sum")
    g, = torch.autograd.grad(out, param, grad)
    with torch.no_grad():
      param.add_(g, alpha=-0.1)

  def sgd_bag():
    out = bag(ids.values, ids.row_splits[:-1])
    g, = torch.autograd.grad(out, bag.weight, grad)
    with torch.no_grad():
      bag.weight.add_(g, alpha=-0.1)

  print(f"custom fwd+bwd+sgd        {timeit(sgd_custom):8.3f} ms")
  print(f"EmbeddingBag fwd+bwd+sgd  {timeit(sgd_bag):8.3f} ms")
  w = param.detach()
  t = timeit(lambda: elo.scatter_add_rows(w, ids.values, ids.row_splits, 0, batch, 0, grad, -0.1))
  print(f"fused scatter-add SGD update (no sparse grad) {t:8.3f} ms")


if __name__ == "__main__":
  main()

#!/usr/bin/env python
"""Per-kernel device-time breakdown of one DLRM training step (torch.profiler / CUPTI).
Usage: python tools/profile_step.py [--model dlrm-small] [--batch 65536] [--steps 5]
       [--min-table-rows N]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from bench import table_sizes_for
from distributed_embeddings_b200.models.dlrm import DLRM
from distributed_embeddings_b200.models.trainer import HybridTrainer

p = argparse.ArgumentParser()
p.add_argument("--model", default="dlrm-small")
p.add_argument("--batch", type=int, default=65536)
p.add_argument("--steps", type=int, default=5)
p.add_argument("--optimizer", default="sgd")
p.add_argument("--out", default="gpurun_out/profile_step.txt")
p.add_argument("--trainer", default="fast")
p.add_argument("--min-table-rows", type=int, default=0,
               help="lift every table to at least this many rows: comparing the scatter_add kernel "
               "time with 0 and with 1000000 isolates same-address atomic contention on tiny tables")
args = p.parse_args()

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
sizes = [max(s, args.min_table_rows) for s in table_sizes_for(args.model)]
model = DLRM(sizes, device=dev, compute_dtype=torch.bfloat16, backend="fused")
if args.trainer == "fast":
  from distributed_embeddings_b200.models.dlrm_fast import DLRMTrainStep
  trainer = DLRMTrainStep(model, lr=24.0, embedding_optimizer=args.optimizer, use_cuda_graph=False)
else:
  trainer = HybridTrainer(model, lr=24.0, embedding_optimizer=args.optimizer)
b = args.batch
num = torch.rand(b, 13, device=dev)
cat = [torch.randint(0, s, (b,), device=dev, dtype=torch.int32) for s in sizes]
if args.trainer == "fast":
  cat = torch.stack(cat)
lab = torch.randint(0, 2, (b, 1), device=dev).float()
for _ in range(5):
  trainer.step(num, cat, lab)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
  for _ in range(args.steps):
    trainer.step(num, cat, lab)
  torch.cuda.synchronize()
table = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90)
os.makedirs(os.path.dirname(args.out), exist_ok=True)
with open(args.out, "w") as f:
  f.write(f"# {args.model} batch {b} optimizer {args.optimizer}, {args.steps} steps\n")
  f.write(table)
print(table[-6000:])

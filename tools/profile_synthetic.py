#!/usr/bin/env python
"""Kernel breakdown of one synthetic-model training step (eager, torch.profiler)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from distributed_embeddings_b200.models.configs import synthetic_models_v3
from distributed_embeddings_b200.models.synthetic import InputGenerator, SyntheticModel
from distributed_embeddings_b200.models.trainer import HybridTrainer

p = argparse.ArgumentParser()
p.add_argument("--model", default="tiny")
p.add_argument("--optimizer", default="adagrad")
p.add_argument("--batch", type=int, default=65536)
p.add_argument("--alpha", type=float, default=1.05)
p.add_argument("--out", default="gpurun_out/profile_synth.txt")
a = p.parse_args()
dev = torch.device("cuda", 0)
cfg = synthetic_models_v3[a.model]
model = SyntheticModel(cfg, dp_input=True, device=dev, compute_dtype=torch.bfloat16)
gen = InputGenerator(cfg, a.batch, alpha=a.alpha, num_batches=1, device=dev)
tr = HybridTrainer(model, lr=0.001, embedding_optimizer=a.optimizer)
(num, cat), lab = gen[0]
for _ in range(4):
  tr.step(num, cat, lab)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  for _ in range(3):
    tr.step(num, cat, lab)
  torch.cuda.synchronize()
t = prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=80)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
open(a.out, "w").write(f"# {a.model} {a.optimizer} batch {a.batch} alpha {a.alpha}: 3 steps\n" + t)

#!/usr/bin/env python
"""IntegerLookup (on-the-fly vocabulary) + DLRM end to end (BASELINE config 5).

Raw 64-bit hashed categorical keys arrive model-parallel (every rank receives the global batch of
the features it owns, `dp_input=False`), go through that rank's `IntegerLookup` tables (GPU hash
insert/probe kernel) and the resulting contiguous indices feed the hybrid-parallel DLRM.
Run under torchrun for N > 1.  Prints one JSON line on rank 0 (device timed, max over ranks)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import distributed_embeddings_b200 as de
from distributed_embeddings_b200.models.dlrm import DLRM
from distributed_embeddings_b200.models.trainer import HybridTrainer

p = argparse.ArgumentParser()
p.add_argument("--vocab", type=int, default=1000000)
p.add_argument("--batch_size", type=int, default=65536)
p.add_argument("--steps", type=int, default=30)
p.add_argument("--warmup", type=int, default=8)
p.add_argument("--data_batches", type=int, default=4)
a = p.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lr_ = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr_); dev = torch.device("cuda", lr_)
if world > 1:
  os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
  dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(1)
sizes = 26 * [a.vocab + 1]
model = DLRM(sizes, device=dev, compute_dtype=torch.bfloat16, dp_input=False)
de.broadcast_variables(model)
my_feats = model.embedding.strategy.input_ids_list[rank]
lookups = [de.IntegerLookup(a.vocab, device=dev) for _ in my_feats]
trainer = HybridTrainer(model, lr=1.0, embedding_optimizer="sgd", use_cuda_graph=False)
lb = a.batch_size // world
g = torch.Generator().manual_seed(5)  # same stream on all ranks -> consistent global batch
pool = []
for _ in range(a.data_batches):
  u = torch.rand(a.batch_size, 26, generator=g)
  raw = ((u.pow(3) * 4 * a.vocab).long() * 2654435761 + 12345) % (2**40)  # skewed hashed keys
  num = torch.rand(a.batch_size, 13, generator=g)
  lab = torch.randint(0, 2, (a.batch_size, 1), generator=g).float()
  pool.append((num[rank * lb:(rank + 1) * lb].to(dev), [raw[:, f].contiguous().to(dev) for f in my_feats],
               lab[rank * lb:(rank + 1) * lb].to(dev)))

def step(i):
  num, keys, lab = pool[i % len(pool)]
  ids = [lk(k) for lk, k in zip(lookups, keys)]     # raw keys -> contiguous indices (GPU hash)
  return trainer.step(num, ids, lab)

for i in range(a.warmup):
  step(i)
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for i in range(a.steps):
  loss = step(i)
t1.record(); torch.cuda.synchronize()
ms = torch.tensor([t0.elapsed_time(t1) / a.steps], device=dev)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
vocab_sizes = [lk.vocabulary_size() for lk in lookups[:3]]
if rank == 0:
  print(json.dumps({"bench": "integer_lookup+dlrm", "n_gpus": world, "global_batch": a.batch_size,
                    "ms_per_step": float(ms), "samples_per_sec": a.batch_size / float(ms) * 1e3,
                    "vocab": a.vocab, "vocab_sizes_first3": vocab_sizes, "loss": float(loss)}))
if world > 1: dist.destroy_process_group()

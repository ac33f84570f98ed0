#!/usr/bin/env python
"""DLRM training with hybrid-parallel embeddings (reference examples/dlrm/main.py).

  torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/dlrm/main.py \
      --dataset_path /data/criteo_split_binary        # real data (split binary Criteo)
  python examples/dlrm/main.py --num_batches 100                       # synthetic data

Embeddings are model parallel (memory_balanced placement), MLPs data parallel; SGD lr 24 with
warm-up + polynomial decay; AUC on the eval split; embedding weights saved with np.savez in the
global (sharding independent) layout.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import distributed_embeddings_b200 as de
from distributed_embeddings_b200.models.dlrm import DLRM
from distributed_embeddings_b200.models.trainer import HybridTrainer
from distributed_embeddings_b200.utils.criteo import DummyDataset, RawBinaryDataset
from distributed_embeddings_b200.utils.lr_schedule import LearningRateScheduler
from distributed_embeddings_b200.utils.metrics import binary_auc


def parse():
  p = argparse.ArgumentParser()
  p.add_argument("--dataset_path", default=None, help="dir with model_size.json + train/ test/")
  p.add_argument("--learning_rate", type=float, default=24)
  p.add_argument("--batch_size", type=int, default=64 * 1024, help="global batch size")
  p.add_argument("--top_mlp_dims", default="1024,1024,512,256,1")
  p.add_argument("--bottom_mlp_dims", default="512,256,128")
  p.add_argument("--num_numerical_features", type=int, default=13)
  p.add_argument("--num_batches", type=int, default=340, help="synthetic train batches")
  p.add_argument("--table_sizes", default=",".join(["1000"] * 26))
  p.add_argument("--embedding_dim", type=int, default=128)
  p.add_argument("--dp_input", action="store_true")
  p.add_argument("--test_combiner", action="store_true")
  p.add_argument("--dist_strategy", default="memory_balanced")
  p.add_argument("--fast", action="store_true", help="hand-scheduled step + CUDA graph")
  p.add_argument("--amp", action="store_true", default=True)
  p.add_argument("--warmup_steps", type=int, default=8000)
  p.add_argument("--decay_start_step", type=int, default=48000)
  p.add_argument("--decay_steps", type=int, default=24000)
  p.add_argument("--epochs", type=int, default=1)
  p.add_argument("--save_path", default="/tmp/embedding_weights")
  p.add_argument("--save_dir", default=None,
                 help="write one global-layout .npy per table into this directory instead; every "
                      "rank writes its own slices in parallel (no gather to rank 0)")
  return p.parse_args()


def main():
  args = parse()
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  cuda = torch.cuda.is_available()
  device = torch.device("cuda", local_rank) if cuda else torch.device("cpu")
  if cuda:
    torch.cuda.set_device(device)
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl" if cuda else "gloo", device_id=device if cuda else None)

  if args.dataset_path is not None:
    with open(os.path.join(args.dataset_path, "model_size.json"), encoding="utf-8") as f:
      table_sizes = [s + 1 for s in json.load(f).values()]
  else:
    table_sizes = [int(s) for s in args.table_sizes.split(",")]

  model = DLRM(table_sizes, embedding_dim=args.embedding_dim,
               bottom_mlp_dims=[int(d) for d in args.bottom_mlp_dims.split(",")],
               top_mlp_dims=[int(d) for d in args.top_mlp_dims.split(",")],
               num_numerical_features=args.num_numerical_features, dp_input=args.dp_input,
               dist_strategy=args.dist_strategy, test_combiner=args.test_combiner, device=device,
               compute_dtype=torch.bfloat16 if (args.amp and cuda) else torch.float32)
  table_ids = list(range(len(table_sizes))) if args.dp_input else \
      model.embedding.strategy.input_ids_list[rank]
  lbs = args.batch_size // world
  if args.dataset_path is not None:
    kw = dict(batch_size=args.batch_size, numerical_features=args.num_numerical_features,
              categorical_features=table_ids, categorical_feature_sizes=table_sizes,
              prefetch_depth=10, drop_last_batch=True, offset=lbs * rank, lbs=lbs,
              dp_input=args.dp_input)
    train = RawBinaryDataset(args.dataset_path, **kw)
    evald = RawBinaryDataset(args.dataset_path, valid=True, **kw)
  else:
    train = DummyDataset(args.batch_size, args.num_numerical_features, world, len(table_ids), True,
                         args.dp_input, args.num_batches)
    evald = DummyDataset(args.batch_size, args.num_numerical_features, world, len(table_ids),
                         False, args.dp_input, max(1, args.num_batches // 10))

  sched = LearningRateScheduler(args.learning_rate, warmup_steps=args.warmup_steps,
                                decay_start_step=args.decay_start_step,
                                decay_steps=args.decay_steps)
  de.broadcast_variables(model)
  if args.fast and cuda and args.dp_input:
    from distributed_embeddings_b200.models.dlrm_fast import DLRMTrainStep
    trainer = DLRMTrainStep(model, lr=args.learning_rate, scheduler=sched)
    step = lambda n, c, l: trainer.step(n, torch.stack([x.to(torch.int32) for x in c]), l)
  else:
    trainer = HybridTrainer(model, lr=args.learning_rate, scheduler=sched)
    step = trainer.step

  def batches():
    for _ in range(args.epochs):
      yield from train

  for i, (num, cat, lab) in enumerate(batches()):
    num, lab = num.to(device).float(), lab.to(device)
    cat = [c.to(device) for c in cat]
    if args.test_combiner:
      cat = [c.reshape(-1, 1) for c in cat]
    loss = step(num, cat, lab)
    if i % 1000 == 0:
      loss = loss.detach().clone().reshape(())
      if world > 1:
        dist.all_reduce(loss)
        loss /= world
      if rank == 0:
        print("step: ", i, " loss: ", float(loss))

  # evaluation: predictions of the local batch gathered on every rank, AUC on rank 0
  preds, labels = [], []
  model.eval()
  with torch.no_grad():
    for num, cat, lab in evald:
      cat = [c.to(device) for c in cat]
      if args.test_combiner:
        cat = [c.reshape(-1, 1) for c in cat]
      p = torch.sigmoid(model(num.to(device).float(), cat).float())
      if world > 1:
        out = [torch.empty_like(p) for _ in range(world)]
        dist.all_gather(out, p)
        p = torch.cat(out)
      preds.append(p.cpu())
      labels.append(lab.reshape(-1, 1).float().cpu())
  if rank == 0 and preds:
    p, y = torch.cat(preds).reshape(-1), torch.cat(labels).reshape(-1)
    n = min(p.numel(), y.numel())
    auc = binary_auc(y[:n], p[:n])
    bce = torch.nn.functional.binary_cross_entropy(p[:n].clamp(1e-7, 1 - 1e-7), y[:n])
    print(f"Evaluation completed, AUC: {auc}, test_loss: {float(bce)}")

  if args.save_dir:
    paths = model.embedding.save_weights(args.save_dir)
    if rank == 0:
      print(f"saved {len(paths)} tables to {args.save_dir}/")
  else:
    weights = model.embedding.get_weights()
    if rank == 0:
      np.savez(args.save_path, *weights)
      print(f"saved {len(weights)} tables to {args.save_path}.npz")
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()

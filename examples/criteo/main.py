#!/usr/bin/env python
"""Single-GPU model on raw (hashed hex) Criteo categorical features with on-the-fly vocabulary
building: IntegerLookup(100000) -> Embedding(100000, 128) x 26 -> MLP
(reference examples/criteo/main.py:56-91).

  python examples/criteo/main.py                                   # synthetic raw ids
  python examples/criteo/main.py --data train.txt --epochs 2       # Criteo Kaggle TSV: label,
                                                                   # 13 counts, 26 hex strings
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch
from torch import nn

import distributed_embeddings_b200 as de


class CriteoModel(nn.Module):

  def __init__(self, num_cat=26, vocab=100000, dim=128, num_numerical=13, device=None):
    super().__init__()
    self.lookups = nn.ModuleList([de.IntegerLookup(vocab, device=device) for _ in range(num_cat)])
    self.embeddings = nn.ModuleList(
        [de.Embedding(vocab + 1, dim, device=device) for _ in range(num_cat)])
    d = num_cat * dim + num_numerical
    self.mlp = nn.Sequential(nn.Linear(d, 512), nn.ReLU(), nn.Linear(512, 256), nn.ReLU(),
                             nn.Linear(256, 1)).to(device)

  def forward(self, numerical, categorical):
    outs = []
    for lk, emb, c in zip(self.lookups, self.embeddings, categorical):
      outs.append(emb(lk(c)))
    return self.mlp(torch.cat(outs + [numerical], dim=1))


def synthetic_batches(n_batches, batch, num_cat, device, seed=0):
  """Raw 32-bit hashed ids with a power-law popularity, like hashed hex Criteo columns."""
  g = torch.Generator().manual_seed(seed)
  for _ in range(n_batches):
    u = torch.rand(batch, num_cat, generator=g)
    raw = (u.pow(4) * 2**31).to(torch.int64) * 2654435761 % 2**32  # spread over the id space
    yield (torch.rand(batch, 13, generator=g).to(device), [raw[:, i].to(device)
                                                            for i in range(num_cat)],
           torch.randint(0, 2, (batch, 1), generator=g).float().to(device))


def tsv_batches(path, batch, device, max_rows=None, epochs=1):
  """Criteo (Kaggle / Terabyte) text format: tab-separated ``label, I1..I13, C1..C26`` with empty fields
  for missing values.  Missing values become 0, the numerical columns are min-max scaled over
  the rows read, the categorical hex strings become int64 keys for ``IntegerLookup`` - what the
  reference example does with pandas + sklearn (examples/criteo/main.py:26-53)."""
  import numpy as np
  labels, nums, cats = [], [], []
  with open(path, encoding="ascii") as f:
    for n, line in enumerate(f):
      if max_rows is not None and n >= max_rows:
        break
      parts = line.rstrip("\n").split("\t")
      parts += [""] * (40 - len(parts))
      labels.append(int(parts[0] or 0))
      nums.append([float(x) if x else 0.0 for x in parts[1:14]])
      cats.append([int(x, 16) if x else 0 for x in parts[14:40]])
  lab = torch.tensor(labels, dtype=torch.float32).reshape(-1, 1)
  num = np.asarray(nums, dtype=np.float32)
  lo, hi = num.min(0, keepdims=True), num.max(0, keepdims=True)
  num = torch.from_numpy((num - lo) / np.maximum(hi - lo, 1e-12))
  cat = torch.tensor(cats, dtype=torch.int64)
  for _ in range(epochs):
    for i in range(0, len(lab), batch):
      sl = slice(i, i + batch)
      yield (num[sl].to(device), [cat[sl, j].to(device) for j in range(26)], lab[sl].to(device))


def main():
  p = argparse.ArgumentParser()
  p.add_argument("--batch_size", type=int, default=16384)
  p.add_argument("--steps", type=int, default=50)
  p.add_argument("--vocab", type=int, default=100000)
  p.add_argument("--data", default=None, help="Criteo text file (label, 13 counts, 26 hex ids)")
  p.add_argument("--epochs", type=int, default=1)
  p.add_argument("--max_rows", type=int, default=None)
  args = p.parse_args()
  device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
  model = CriteoModel(vocab=args.vocab, device=device)
  dense = [p_ for n, p_ in model.named_parameters() if "embeddings" not in n]
  sparse = [p_ for n, p_ in model.named_parameters() if "embeddings" in n]
  opt_d = torch.optim.Adam(dense, lr=1e-3)
  opt_s = torch.optim.SparseAdam(sparse, lr=1e-3)
  bce = nn.BCEWithLogitsLoss()
  if args.data:
    batches = tsv_batches(args.data, args.batch_size, device, args.max_rows, args.epochs)
  else:
    batches = synthetic_batches(args.steps, args.batch_size, 26, device)
  for i, (num, cat, lab) in enumerate(batches):
    opt_d.zero_grad()
    opt_s.zero_grad()
    loss = bce(model(num, cat), lab)
    loss.backward()
    opt_d.step()
    opt_s.step()
    if i % 10 == 0:
      sizes = [lk.vocabulary_size() for lk in model.lookups[:3]]
      print(f"step {i} loss {loss.item():.4f} vocab sizes (first 3 features) {sizes}")


if __name__ == "__main__":
  main()

#!/usr/bin/env python
"""Write a small synthetic dataset in the *split binary* Criteo layout that
``distributed_embeddings_b200.utils.criteo.RawBinaryDataset`` and ``examples/dlrm/main.py
--dataset_path`` read:

    <out>/model_size.json            {"cat_0": max id, ...}   (cardinality - 1, like the real set)
    <out>/{train,test}/label.bin     bool   [n]
    <out>/{train,test}/numerical.bin fp16   [n, num_numerical]
    <out>/{train,test}/cat_<i>.bin   int8 / int16 / int32 by cardinality, [n]

Labels follow a logistic model of the features, so a few hundred training steps lift the AUC
clearly above 0.5 - enough to exercise the data path end to end without the 1 TB original.

  python tools/make_synthetic_criteo.py /tmp/criteo_small --train 65536 --test 8192
"""
import argparse
import json
import os

import numpy as np


def cat_dtype(size):
  for t in (np.int8, np.int16, np.int32):
    if size < np.iinfo(t).max:
      return t
  raise ValueError(size)


def write_split(path, n, sizes, num_numerical, rng, weights):
  os.makedirs(path, exist_ok=True)
  num = rng.standard_normal((n, num_numerical)).astype(np.float16)
  logit = num.astype(np.float32) @ weights["num"]
  cats = []
  for i, s in enumerate(sizes):
    # skewed ids (most mass on the small ids), like real categorical features
    ids = np.minimum((rng.pareto(1.2, size=n) * s / 50).astype(np.int64), s - 1)
    cats.append(ids)
    logit += weights["cat"][i][ids]
  p = 1.0 / (1.0 + np.exp(-logit))
  label = rng.random(n) < p
  label.astype(np.bool_).tofile(os.path.join(path, "label.bin"))
  num.tofile(os.path.join(path, "numerical.bin"))
  for i, (s, ids) in enumerate(zip(sizes, cats)):
    ids.astype(cat_dtype(s)).tofile(os.path.join(path, f"cat_{i}.bin"))
  return float(label.mean())


def main(argv=None):
  ap = argparse.ArgumentParser(description=__doc__,
                               formatter_class=argparse.RawTextHelpFormatter)
  ap.add_argument("out")
  ap.add_argument("--train", type=int, default=65536, help="training samples")
  ap.add_argument("--test", type=int, default=8192, help="evaluation samples")
  ap.add_argument("--table_sizes", default=None,
                  help="comma separated cardinalities (default: 26 sizes from 3 to 40000)")
  ap.add_argument("--num_numerical", type=int, default=13)
  ap.add_argument("--seed", type=int, default=0)
  args = ap.parse_args(argv)
  if args.table_sizes:
    sizes = [int(s) for s in args.table_sizes.split(",")]
  else:
    sizes = [int(x) for x in np.unique(np.geomspace(3, 40000, 26).astype(np.int64))]
    while len(sizes) < 26:
      sizes.append(sizes[-1] + 7)
  rng = np.random.default_rng(args.seed)
  weights = {"num": (rng.standard_normal(args.num_numerical) * 0.5).astype(np.float32),
             "cat": [(rng.standard_normal(s) * 0.7).astype(np.float32) for s in sizes]}
  os.makedirs(args.out, exist_ok=True)
  with open(os.path.join(args.out, "model_size.json"), "w", encoding="utf-8") as f:
    json.dump({f"cat_{i}": s - 1 for i, s in enumerate(sizes)}, f)
  pos_tr = write_split(os.path.join(args.out, "train"), args.train, sizes, args.num_numerical,
                       rng, weights)
  pos_te = write_split(os.path.join(args.out, "test"), args.test, sizes, args.num_numerical, rng,
                       weights)
  print(f"wrote {args.train} train / {args.test} test samples, {len(sizes)} categorical features "
        f"to {args.out} (positives: {pos_tr:.3f} / {pos_te:.3f})")
  return sizes


if __name__ == "__main__":
  main()

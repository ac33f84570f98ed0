"""Utilities: initializers, lr schedule, AUC, model zoo facts, regularizer/constraint hooks."""
import math

import numpy as np
import pytest
import torch

import distributed_embeddings_b200 as de
from distributed_embeddings_b200.models import configs
from distributed_embeddings_b200.utils import initializers
from distributed_embeddings_b200.utils.lr_schedule import LearningRateScheduler
from distributed_embeddings_b200.utils.metrics import binary_auc


def test_initializers():
  t = initializers.get("uniform")((1000, 8))
  assert t.abs().max() <= 0.05 and t.std() > 0.01
  d = initializers.DLRMInitializer()((400, 4))
  assert d.abs().max() <= 1 / math.sqrt(400) + 1e-7
  c = initializers.ConcatInitializer(initializers.DLRMInitializer(), [100, 10000])((10100, 4))
  assert c[:100].abs().max() > c[100:].abs().max() * 3  # every member sees its own row count
  assert torch.count_nonzero(initializers.get("zeros")((3, 3))) == 0
  z = initializers.get({"class_name": "RandomUniform", "config": {"minval": 1.0, "maxval": 2.0}})
  assert z((50, 2)).min() >= 1.0
  f = initializers.get(lambda shape, dtype=None, device=None: torch.full(shape, 7.0))
  assert f((2, 2)).eq(7).all()
  with pytest.raises(ValueError):
    initializers.get("nope")


def test_lr_schedule_matches_reference_shape():
  s = LearningRateScheduler(24.0, warmup_steps=8000, decay_start_step=48000, decay_steps=24000)
  assert s.lr_at(0) == 0.0
  assert s.lr_at(4000) == pytest.approx(12.0)
  assert s.lr_at(8000) == s.lr_at(47999) == 24.0
  assert s.lr_at(60000) == pytest.approx(24.0 * 0.25)
  assert s.lr_at(72000) == 0.0
  assert [s.step() for _ in range(3)] == [s.lr_at(0), s.lr_at(1), s.lr_at(2)]


def test_auc():
  y = torch.tensor([0, 0, 1, 1.])
  assert binary_auc(y, torch.tensor([0.1, 0.4, 0.35, 0.8])) == pytest.approx(0.75)
  assert binary_auc(y, torch.tensor([0.1, 0.2, 0.7, 0.8])) == 1.0
  assert binary_auc(y, torch.tensor([0.5, 0.5, 0.5, 0.5])) == pytest.approx(0.5)
  g = torch.Generator().manual_seed(0)
  s = torch.rand(5000, generator=g)
  lab = (torch.rand(5000, generator=g) < s).float()
  # compare with a direct pair count on a subsample
  pos, neg = s[lab > 0.5][:300], s[lab < 0.5][:300]
  direct = ((pos[:, None] > neg[None, :]).float().mean() +
            0.5 * (pos[:, None] == neg[None, :]).float().mean())
  sub_y = torch.cat([torch.ones(len(pos)), torch.zeros(len(neg))])
  assert binary_auc(sub_y, torch.cat([pos, neg])) == pytest.approx(float(direct), abs=1e-6)


def test_model_zoo_matches_published_sizes():
  # reference README table: tiny 55 tables / 4.2 GiB, small 107 / 26.3, large 612 / 773.8
  facts = {"tiny": (55, 58, 4.2, 672, 85), "small": (107, 116, 26.3, 2512, 377),
           "medium": (311, 337, 206.2, 17280, 1611), "large": (612, 669, 773.8, 36608, 6312),
           "jumbo": (1022, 1097, 3109.5, 118400, 16022)}
  for name, (tables, inputs, gib, width, lookups) in facts.items():
    s = configs.summary(configs.synthetic_models_v3[name])
    assert (s["tables"], s["inputs"], s["output_width"], s["lookups_per_sample"]) == \
        (tables, inputs, width, lookups)
    assert s["gib_fp32"] == pytest.approx(gib, abs=0.06)


def test_regularizer_and_constraint_hooks():
  e = de.Embedding(10, 4, embeddings_regularizer=lambda w: 0.5 * w.pow(2).sum(),
                   embeddings_constraint=lambda w: w.clamp(-0.01, 0.01))
  loss = e.regularization_loss()
  assert loss.item() == pytest.approx(0.5 * e.embeddings.detach().pow(2).sum().item())
  loss.backward()
  torch.testing.assert_close(e.embeddings.grad, e.embeddings.detach())
  e.apply_constraint()
  assert e.embeddings.abs().max() <= 0.01 + 1e-9
  assert de.Embedding(3, 2).regularization_loss().item() == 0.0


def test_power_law_generator_is_skewed_and_in_range():
  from distributed_embeddings_b200.models.synthetic import gen_power_law_data
  ids = gen_power_law_data(20000, 3, 1000, 1.05, np.random.default_rng(0))
  assert ids.min() >= 0 and ids.max() < 1000
  assert (ids == 0).float().mean() > 0.05  # heavy head


def test_p2p_domain_detection():
  from distributed_embeddings_b200.parallel.comm import CommContext, host_identity, single_p2p_domain
  me = host_identity()
  assert me == host_identity() and "/" in me
  assert single_p2p_domain([me] * 8, 16)
  assert not single_p2p_domain([me] * 4 + ["other-host/1"] * 4, 16)   # two nodes
  assert not single_p2p_domain([me] * 32, 16)                         # more ranks than peer slots
  ctx = CommContext(device="cpu")
  assert not ctx.p2p and ctx.world_size == 1 and ctx.p2p_unavailable_reason is None


def test_model_zoo_details():
  """Table counts and fp32 sizes of the synthetic zoo as published by the reference
  (examples/benchmarks/synthetic_models/README.md:11-16; config_v3.py:46-92)."""
  from distributed_embeddings_b200.models.configs import expand, scaled, summary, synthetic_models_v3
  published = {"tiny": (55, 4.2), "small": (107, 26.3), "medium": (311, 206.2),
               "large": (612, 773.8), "jumbo": (1022, 3109.5), "colossal": (2002, 22327.4)}
  for name, (tables, gib) in published.items():
    s = summary(synthetic_models_v3[name])
    assert s["tables"] == tables
    assert round(s["gib_fp32"], 1) == gib
  small = summary(synthetic_models_v3["small"])
  assert (small["inputs"], small["elements"], small["output_width"],
          small["lookups_per_sample"]) == (116, 7058084800, 2512, 377)
  large = summary(synthetic_models_v3["large"])
  assert (large["inputs"], large["output_width"], large["lookups_per_sample"]) == (669, 36608, 6312)
  crit = summary(synthetic_models_v3["criteo"])
  assert (crit["tables"], crit["rows"], crit["output_width"]) == (26, 2600000, 3328)
  # scaling shrinks rows only; expand() yields one entry per table / input
  tiny = scaled(synthetic_models_v3["tiny"], 0.001)
  st = summary(tiny)
  assert st["tables"] == 55 and st["output_width"] == 672 and st["rows"] < 100000
  tables, imap, hots = expand(tiny)[:3]
  assert len(tables) == 55 and len(imap) == 58 and len(hots) == 58


def _make_dataset(tmp_path, n_train=96, n_test=32, sizes="5,300,70000"):
  import importlib.util
  import os
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  spec = importlib.util.spec_from_file_location("mk", os.path.join(root, "tools",
                                                                   "make_synthetic_criteo.py"))
  mk = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mk)
  out = str(tmp_path / "criteo")
  got = mk.main([out, "--train", str(n_train), "--test", str(n_test), "--table_sizes", sizes,
                 "--num_numerical", "4", "--seed", "3"])
  return out, got


def test_raw_binary_dataset_round_trip(tmp_path):
  """The split-binary Criteo reader returns exactly what is on disk: integer widths chosen by
  cardinality (int8 / int16 / int32), fp16 numericals, bool labels; rank slices; both input
  modes; the prefetching iterator (reference examples/dlrm/utils.py:116-307)."""
  import os
  from distributed_embeddings_b200.utils.criteo import RawBinaryDataset
  path, sizes = _make_dataset(tmp_path)
  assert sizes == [5, 300, 70000]
  tr = os.path.join(path, "train")
  lab = np.fromfile(os.path.join(tr, "label.bin"), dtype=np.bool_)
  num = np.fromfile(os.path.join(tr, "numerical.bin"), dtype=np.float16).reshape(-1, 4)
  cats = [np.fromfile(os.path.join(tr, f"cat_{i}.bin"), dtype=t)
          for i, t in enumerate((np.int8, np.int16, np.int32))]
  assert len(lab) == 96 and all(len(c) == 96 for c in cats) and int(cats[2].max()) < 70000
  bs = 32
  # model-parallel inputs: rank 1 of 2 owns features [2, 0]; dp tensors are its batch slice
  ds = RawBinaryDataset(path, batch_size=bs, numerical_features=4, categorical_features=[2, 0],
                        categorical_feature_sizes=sizes, offset=16, lbs=16, pin_memory=False)
  assert len(ds) == 3
  items = list(ds)  # prefetch thread
  assert len(items) == 3
  for b, (n_, c_, l_) in enumerate(items):
    sl = slice(b * bs, (b + 1) * bs)
    np.testing.assert_array_equal(n_.numpy(), num[sl][16:32])
    np.testing.assert_array_equal(l_.numpy().reshape(-1), lab[sl][16:32].astype(np.float32))
    assert [x.dtype for x in c_] == [torch.int32, torch.int32]
    np.testing.assert_array_equal(c_[0].numpy(), cats[2][sl])  # global batch of an owned feature
    np.testing.assert_array_equal(c_[1].numpy(), cats[0][sl])
  # data-parallel inputs: everything sliced
  dp = RawBinaryDataset(path, batch_size=bs, numerical_features=4, categorical_features=[0, 1, 2],
                        categorical_feature_sizes=sizes, offset=0, lbs=16, dp_input=True,
                        pin_memory=False)
  n_, c_, l_ = dp[1]
  np.testing.assert_array_equal(c_[1].numpy(), cats[1][32:48])
  assert n_.shape == (16, 4) and l_.shape == (16, 1)
  with pytest.raises(IndexError):
    dp[3]  # pylint: disable=pointless-statement
  # evaluation split keeps the labels of the global batch (AUC is computed on gathered scores)
  ev = RawBinaryDataset(path, batch_size=bs, numerical_features=4, categorical_features=[0],
                        categorical_feature_sizes=sizes, valid=True, offset=16, lbs=16,
                        pin_memory=False)
  n_, c_, l_ = ev[0]
  assert len(ev) == 1 and l_.shape == (32, 1) and n_.shape == (16, 4)
  # a ragged tail is dropped or kept
  full = RawBinaryDataset(path, batch_size=40, categorical_features=[0],
                          categorical_feature_sizes=sizes, pin_memory=False)
  drop = RawBinaryDataset(path, batch_size=40, categorical_features=[0],
                          categorical_feature_sizes=sizes, drop_last_batch=True, pin_memory=False)
  assert len(full) == 3 and len(drop) == 2 and full[0][0] is None

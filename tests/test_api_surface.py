"""Small public entry points that the larger suites only reach indirectly: ragged / sparse
containers, the functional lookups, initializer (de)serialisation, shape inference, parameter
partitioning (reference API parity: embedding_lookup_ops.py:37-128, embedding.py:65-170)."""
import numpy as np
import pytest
import torch

import distributed_embeddings_b200 as de
from distributed_embeddings_b200 import dist_model_parallel as dmp
from distributed_embeddings_b200.ops import embedding_lookup_ops as elo
from distributed_embeddings_b200.ops.ragged import RaggedIds, SparseIds
from distributed_embeddings_b200.utils import initializers


def test_ragged_and_sparse_containers():
  ids = torch.tensor([[1, 2, 3], [4, 5, 6]])
  r = RaggedIds.from_dense(ids)
  assert r.nrows == 2 and r.shape == (2, None) and r.row_splits.tolist() == [0, 3, 6]
  assert r.values.tolist() == [1, 2, 3, 4, 5, 6] and r.dtype == torch.int64
  l = RaggedIds.from_lists([[7], [], [8, 9]])
  assert l.nrows == 3 and l.row_splits.tolist() == [0, 1, 1, 3]
  assert l.slice_rows(1, 3).values.tolist() == [8, 9]
  coo = torch.sparse_coo_tensor(torch.tensor([[0, 0, 2], [0, 1, 0]]), torch.tensor([5, 6, 7]),
                                size=(3, 2))
  s = SparseIds.from_torch_sparse(coo)
  assert s.values.tolist() == [5, 6, 7] and tuple(s.dense_shape) == (3, 2)
  assert s.indices.tolist() == [[0, 0], [0, 1], [2, 0]]


@pytest.mark.parametrize("combiner", ["sum", "mean"])
def test_functional_lookups_agree(combiner):
  torch.manual_seed(0)
  w = torch.randn(20, 6, requires_grad=True)
  ids = torch.randint(0, 20, (5, 3))
  ref = w[ids].sum(1) if combiner == "sum" else w[ids].mean(1)
  a = elo.embedding_lookup(w, ids, combiner)
  b = elo.embedding_lookup_fixed_hotness(w, ids, combiner)
  r = RaggedIds.from_dense(ids)
  c = elo.embedding_lookup_variable_hotness(w, r.values, r.row_splits, combiner)
  for x in (a, b, c):
    torch.testing.assert_close(x, ref)
  # sparse (deduplicated) gradient with summed rows, like the reference's custom gradient
  g, = torch.autograd.grad(c.sum(), w)
  dense = torch.zeros(20, 6)
  scale = 1.0 if combiner == "sum" else 1.0 / 3
  dense.index_add_(0, ids.reshape(-1), torch.full((15, 6), scale))
  assert g.is_sparse and g.coalesce().indices().shape[1] == ids.unique().numel()
  torch.testing.assert_close(g.to_dense(), dense)
  with pytest.raises(ValueError):
    elo.embedding_lookup_variable_hotness(w, r.values, r.row_splits, "max")
  v = elo.read_var_no_copy(w)
  assert v.data_ptr() == w.data_ptr() and not v.requires_grad


def test_integer_lookup_cpu_path_and_state_round_trip():
  table = elo.integer_lookup_init(torch.zeros(16, 2, dtype=torch.int64))
  assert int(table.min()) == -1 and int(table.max()) == -1
  layer = de.IntegerLookup(max_tokens=5, use_gpu=False)
  keys = torch.tensor([100, 7, 100, 42, 7, 9, 11, 13, 100])
  out = layer(keys)
  assert out.tolist()[:5] == [1, 2, 1, 3, 2]
  assert out.tolist()[5:8].count(0) >= 1  # the vocabulary is full: later keys are OOV
  vocab = layer.get_vocabulary()
  assert list(vocab[1:4]) == [100, 7, 42]
  clone = de.IntegerLookup(max_tokens=5, use_gpu=False)
  clone.load_state_dict(layer.state_dict())
  assert clone(keys).tolist() == layer(keys).tolist()
  assert list(clone.get_vocabulary()) == list(layer.get_vocabulary())


def test_output_shape_inference_and_repr():
  e = de.Embedding(10, 4)
  assert e.compute_output_shape((7,)) == (7, 4) and e.compute_output_shape((7, 3)) == (7, 3, 4)
  assert tuple(e(torch.zeros(7, 3, dtype=torch.int64)).shape) == (7, 3, 4)
  p = de.Embedding(10, 4, combiner="mean")
  assert p.compute_output_shape((7, 3)) == (7, 4) and p.compute_output_shape((2, 7, 3)) == (2, 7, 4)
  assert tuple(p(torch.zeros(2, 7, 3, dtype=torch.int64)).shape) == (2, 7, 4)
  assert "10" in repr(e) and "4" in repr(e)
  d = dmp.DistributedEmbedding([e, p])
  assert "world_size=1" in repr(d) and "strategy=basic" in repr(d)


def test_initializer_identifiers_round_trip():
  for name in ("uniform", "zeros", "ones"):
    assert isinstance(initializers.get(name), initializers.Initializer)
  n = initializers.RandomNormal(mean=1.0, stddev=0.0)
  cfg = initializers.serialize(n)
  assert cfg == {"class_name": "RandomNormal", "config": {"mean": 1.0, "stddev": 0.0}}
  t = initializers.get(cfg).fill_(torch.empty(3, 2))
  torch.testing.assert_close(t, torch.ones(3, 2))
  c = initializers.get({"class_name": "Constant", "config": {"value": 2.5}})
  torch.testing.assert_close(c.fill_(torch.empty(2)), torch.full((2,), 2.5))
  f = initializers.get(lambda shape: np.arange(np.prod(shape)).reshape(shape))
  assert isinstance(f, initializers.FunctionInitializer) and initializers.serialize(f) is f
  torch.testing.assert_close(f.fill_(torch.empty(2, 2)), torch.tensor([[0., 1.], [2., 3.]]))
  with pytest.raises(ValueError):
    initializers.get("no_such_initializer")
  layer = de.Embedding(4, 2, embeddings_initializer=cfg)
  torch.testing.assert_close(layer.embeddings.detach(), torch.ones(4, 2))
  again = de.Embedding.from_config(layer.get_config())
  torch.testing.assert_close(again.embeddings.detach(), torch.ones(4, 2))


def test_parameter_partition_and_learning_rate():
  d = dmp.DistributedEmbedding([de.Embedding(10, 4, combiner="sum"),
                                de.Embedding(6, 4, combiner="sum")])
  mp, dp = d.mp_parameters(), d.dp_parameters()
  assert len(mp) + len(dp) == len(list(d.parameters())) and len(mp) >= 1 and not dp
  assert all(getattr(p, "de_local", False) for p in mp)
  with pytest.raises(RuntimeError):
    d.set_learning_rate(0.1)
  d.set_optimizer("sgd", lr=0.5)
  d.set_learning_rate(0.25)
  assert d._fused_optimizer["lr"] == 0.25


def test_module_state_dict_round_trip(tmp_path):
  """nn.Module checkpointing (per-rank shards; the sharding-independent form is get_weights /
  save_weights): state_dict -> torch.save -> load_state_dict restores outputs exactly."""
  torch.manual_seed(1)
  make = lambda: dmp.DistributedEmbedding([de.Embedding(10, 4, combiner="sum"),
                                           de.Embedding(6, 8, combiner="mean"),
                                           de.Embedding(9, 4, combiner="sum")])
  a, b = make(), make()
  ids = [torch.randint(0, n, (5, 2)) for n in (10, 6, 9)]
  assert not torch.equal(torch.cat(a(ids), 1), torch.cat(b(ids), 1))
  path = str(tmp_path / "shard.pt")
  torch.save(a.state_dict(), path)
  missing = b.load_state_dict(torch.load(path))
  assert not missing.missing_keys and not missing.unexpected_keys
  torch.testing.assert_close(torch.cat(a(ids), 1), torch.cat(b(ids), 1), rtol=0, atol=0)

"""Layer semantics (reference embedding_test.py:28-191)."""
import pytest
import torch
from torch import nn

import distributed_embeddings_b200 as de
from distributed_embeddings_b200.ops.ragged import RaggedIds, SparseIds

DEVICES = ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("device", DEVICES)
def test_shapes(device):
  e = de.Embedding(10, 4, device=device)
  assert e(torch.tensor([1, 2, 3], device=device)).shape == (3, 4)
  assert e(torch.zeros(3, 5, dtype=torch.int64, device=device)).shape == (3, 5, 4)
  assert e(torch.zeros(2, 3, 5, dtype=torch.int64, device=device)).shape == (2, 3, 5, 4)
  s = de.Embedding(10, 4, combiner="sum", device=device)
  assert s(torch.zeros(3, 5, dtype=torch.int64, device=device)).shape == (3, 4)
  assert s(torch.zeros(2, 3, 5, dtype=torch.int64, device=device)).shape == (2, 3, 4)
  assert s(RaggedIds.from_lists([[1, 2], [3]], device=device)).shape == (2, 4)
  with pytest.raises(ValueError):
    s(torch.tensor([1, 2], device=device))
  with pytest.raises(ValueError):
    de.Embedding(0, 4)


@pytest.mark.parametrize("device", DEVICES)
def test_values_and_adagrad_step(device):
  w = torch.arange(20, dtype=torch.float32).reshape(5, 4)
  e = de.Embedding(5, 4, combiner="sum", device=device)
  with torch.no_grad():
    e.embeddings.copy_(w)
  ids = torch.tensor([[0, 1], [1, 1], [4, 2]], device=device)
  out = e(ids)
  assert torch.equal(out.cpu(), torch.stack([w[0] + w[1], 2 * w[1], w[4] + w[2]]))
  ref = nn.Embedding(5, 4, sparse=True)
  with torch.no_grad():
    ref.weight.copy_(w)
  opt_a = torch.optim.Adagrad(e.parameters(), lr=0.1)
  opt_b = torch.optim.Adagrad(ref.parameters(), lr=0.1)
  out.sum().backward()
  ref(ids.cpu()).sum().backward()
  opt_a.step()
  opt_b.step()
  torch.testing.assert_close(e.embeddings.detach().cpu(), ref.weight.detach())


@pytest.mark.parametrize("device", DEVICES)
def test_mean_and_sparse_input(device):
  e = de.Embedding(6, 2, combiner="mean", device=device)
  rag = RaggedIds.from_lists([[0, 1, 2], [5]], device=device)
  out = e(rag)
  w = e.embeddings.detach()
  torch.testing.assert_close(out, torch.stack([w[:3].mean(0), w[5]]))
  coo = SparseIds(torch.tensor([[0, 0], [0, 1], [0, 2], [1, 0]], device=device),
                  torch.tensor([0, 1, 2, 5], device=device), (2, 3))
  torch.testing.assert_close(e(coo), out)


def test_config_roundtrip_and_stock_configs():
  e = de.Embedding(7, 3, combiner="sum", embeddings_initializer="zeros")
  e2 = de.Embedding.from_config(e.get_config())
  assert (e2.input_dim, e2.output_dim, e2.combiner) == (7, 3, "sum")
  assert torch.count_nonzero(e2.embeddings) == 0
  stock = {"input_dim": 5, "output_dim": 2, "mask_zero": False, "input_length": None}
  assert de.Embedding.from_config(stock).embeddings.shape == (5, 2)


@pytest.mark.parametrize("device", DEVICES)
def test_native_fallback_matches(device):
  a = de.Embedding(9, 4, combiner="sum", device=device)
  b = de.Embedding(9, 4, combiner="sum", use_custom_kernel=False, device=device)
  with torch.no_grad():
    b.embeddings.copy_(a.embeddings)
  ids = torch.randint(0, 9, (5, 3), device=device)
  torch.testing.assert_close(a(ids), b(ids))
  rag = RaggedIds.from_lists([[1], [2, 3, 4], [8, 8]], device=device)
  torch.testing.assert_close(a(rag), b(rag))


@pytest.mark.parametrize("device", DEVICES)
def test_concat_one_hot(device):
  layer = de.ConcatOneHotEmbedding([3, 5, 2], 4, device=device)
  ids = torch.tensor([[0, 4, 1], [2, 0, 0]], device=device)
  out = layer(ids)
  assert out.shape == (2, 3, 4)
  assert torch.equal(out[0, 1], layer.params[3 + 4].detach())
  assert torch.equal(out[1, 2], layer.params[8 + 0].detach())

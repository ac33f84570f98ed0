"""Op-level tests: pooled lookup forward / backward vs a plain PyTorch fp32 oracle.
Mirrors the reference's embedding_lookup_ops_test.py:22-115 (ragged variable hotness incl. rows
longer than 128 ids, fixed hotness, sparse COO; sum & mean) and adds kernel-level cases."""
import pytest
import torch

import distributed_embeddings_b200 as de
from distributed_embeddings_b200.ops import embedding_lookup_ops as elo
from distributed_embeddings_b200.ops.ragged import RaggedIds, SparseIds

DEVICES = ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)]


def oracle(param, rows_of_ids, combiner):
  out = []
  for ids in rows_of_ids:
    v = param[torch.tensor(ids, dtype=torch.int64, device=param.device)]
    out.append(v.sum(0) if combiner == "sum" else v.mean(0))
  return torch.stack(out)


def make_ragged(voc, batch, max_hot, gen, device, dtype=torch.int64):
  lens = torch.randint(1, max_hot + 1, (batch,), generator=gen)
  vals = torch.randint(0, voc, (int(lens.sum()),), generator=gen, dtype=dtype)
  return RaggedIds.from_row_lengths(vals, lens).to(device)


@pytest.mark.parametrize("device", DEVICES)
@pytest.mark.parametrize("combiner", ["sum", "mean"])
@pytest.mark.parametrize("width", [64, 8, 3, 200])
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
def test_variable_hotness(device, combiner, width, dtype):
  gen = torch.Generator().manual_seed(0)
  voc, batch, max_hot = 69, 15, 207
  param = torch.rand(voc, width, generator=gen).to(device).requires_grad_(True)
  ids = make_ragged(voc, batch, max_hot, gen, device, dtype)
  out = elo.embedding_lookup(param, ids, combiner)
  ref_param = param.detach().clone().requires_grad_(True)
  ref = oracle(ref_param, ids.to_lists(), combiner)
  torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
  w = torch.rand(batch, width, generator=gen).to(device)
  (out * w).sum().backward()
  (ref * w).sum().backward()
  assert param.grad.is_sparse
  g = param.grad.coalesce()
  assert torch.equal(g.indices()[0], torch.unique(ids.values.to(torch.int64)))
  torch.testing.assert_close(g.to_dense(), ref_param.grad, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("device", DEVICES)
@pytest.mark.parametrize("combiner", ["sum", "mean", None])
@pytest.mark.parametrize("hot", [1, 4, 7])
def test_fixed_hotness(device, combiner, hot):
  gen = torch.Generator().manual_seed(1)
  voc, batch, width = 50, 33, 16
  param = torch.rand(voc, width, generator=gen).to(device).requires_grad_(True)
  ids = torch.randint(0, voc, (batch, hot), generator=gen).to(device)
  out = elo.embedding_lookup(param, ids, combiner)
  gathered = param.detach()[ids]
  if combiner is None:
    ref = gathered
  else:
    ref = gathered.sum(1) if combiner == "sum" else gathered.mean(1)
  if combiner == "mean" or hot > 1:
    torch.testing.assert_close(out, ref, rtol=1e-6, atol=1e-6)
  else:
    assert torch.equal(out, ref)  # one-hot is a pure copy: bit exact
  out.sum().backward()
  dense = torch.zeros(voc, width, device=device)
  scale = 1.0 / hot if combiner == "mean" else 1.0
  dense.index_add_(0, ids.reshape(-1), torch.full((batch * hot, width), scale, device=device))
  torch.testing.assert_close(param.grad.to_dense(), dense, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("device", DEVICES)
@pytest.mark.parametrize("combiner", ["sum", "mean"])
def test_sparse_coo(device, combiner):
  gen = torch.Generator().manual_seed(2)
  voc, batch, width, max_hot = 40, 12, 32, 9
  param = torch.rand(voc, width, generator=gen).to(device)
  rag = make_ragged(voc, batch, max_hot, gen, "cpu")
  rows = torch.repeat_interleave(torch.arange(batch), rag.row_lengths())
  cols = torch.cat([torch.arange(int(n)) for n in rag.row_lengths()])
  coo = SparseIds(torch.stack([rows, cols], 1).to(device), rag.values.to(device), (batch, max_hot))
  out = elo.embedding_lookup(param, coo, combiner)
  torch.testing.assert_close(out, oracle(param, rag.to_lists(), combiner), rtol=1e-5, atol=1e-5)
  # torch sparse tensors are accepted as well (ids shifted by +1 so zero entries are not dropped)
  t = torch.sparse_coo_tensor(torch.stack([rows, cols]), rag.values + 1, (batch, max_hot)).to(device)
  big = torch.cat([torch.zeros(1, width, device=device), param])
  out2 = elo.embedding_lookup(big, t, combiner)
  torch.testing.assert_close(out2, out, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("device", DEVICES)
def test_row_to_split(device):
  rows = torch.tensor([0, 0, 2, 2, 2, 5], device=device)
  idx = torch.stack([rows, torch.zeros_like(rows)], 1)
  got = elo.row_to_split(idx, 7)
  assert got.tolist() == [0, 2, 2, 5, 5, 5, 6, 6]


@pytest.mark.parametrize("device", DEVICES)
def test_out_of_range_ids_contribute_zero(device):
  param = torch.ones(10, 8, device=device)
  ids = torch.tensor([[0, 12], [-3, 4]], device=device)
  out = elo.embedding_lookup(param, ids, "sum")
  assert torch.equal(out, torch.ones(2, 8, device=device))


@pytest.mark.parametrize("device", DEVICES)
def test_dense_grad_mode(device):
  gen = torch.Generator().manual_seed(3)
  param = torch.rand(30, 12, generator=gen).to(device).requires_grad_(True)
  ids = torch.randint(0, 30, (9, 3), generator=gen).to(device)
  out = elo.embedding_lookup(param, ids, "sum", sparse_grad=False)
  out.sum().backward()
  assert not param.grad.is_sparse
  ref = torch.zeros(30, 12, device=device)
  ref.index_add_(0, ids.reshape(-1), torch.ones(27, 12, device=device))
  torch.testing.assert_close(param.grad, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("width", [4, 16, 32, 128, 256, 1024])
@pytest.mark.parametrize("hot", [1, 3, 64])
def test_cuda_matches_cpu_large(width, hot):
  gen = torch.Generator().manual_seed(4)
  voc, batch = 5000, 1000
  param = torch.rand(voc, width, generator=gen)
  ids = torch.randint(0, voc, (batch, hot), generator=gen)
  ref = elo.embedding_lookup(param, ids, "sum")
  out = elo.embedding_lookup(param.cuda(), ids.cuda(), "sum")
  torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-4)
  grad = torch.rand(batch, width, generator=gen)
  ids_c, rows_c = elo.lookup_grad_sparse(ids.reshape(-1), None, hot, batch, 0, grad, voc)
  ids_g, rows_g = elo.lookup_grad_sparse(ids.reshape(-1).cuda(), None, hot, batch, 0, grad.cuda(),
                                         voc)
  assert torch.equal(ids_g.cpu(), ids_c)
  torch.testing.assert_close(rows_g.cpu(), rows_c, rtol=1e-5, atol=1e-4)
  # bf16 output / gradient paths
  out16 = elo.lookup_forward(param.cuda(), ids.reshape(-1).cuda(), None, hot, batch, 0, True)
  torch.testing.assert_close(out16.float().cpu(), ref, rtol=2e-2, atol=2e-2)


@pytest.mark.gpu
def test_cuda_determinism():
  gen = torch.Generator().manual_seed(5)
  voc, batch, width, hot = 100, 4096, 64, 8
  grad = torch.rand(batch, width, generator=gen).cuda()
  ids = torch.randint(0, voc, (batch * hot,), generator=gen).cuda()
  a = elo.lookup_grad_sparse(ids, None, hot, batch, 0, grad, voc)
  b = elo.lookup_grad_sparse(ids, None, hot, batch, 0, grad, voc)
  assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])  # sorted segments: bitwise stable

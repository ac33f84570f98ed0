"""SparseRowOptimizer (parallel/hybrid.py): the row-sparse update the torch / NCCL back end applies
to model-parallel tables.  Oracle: torch.optim on the dense gradient for the rows that were
touched (the optimizers are *lazy*: untouched rows and their state must not move), hand-written
formulas for row-wise Adagrad.  Reference counterpart: the Keras sparse-apply kernels behind
examples/benchmarks/synthetic_models/main.py:96-101."""
import numpy as np
import pytest
import torch

from distributed_embeddings_b200.parallel.hybrid import SparseRowOptimizer


def _sparse_grad(rows, width, ids, vals):
  # uncoalesced on purpose: duplicate ids must be summed before the non-linear optimizers see them
  return torch.sparse_coo_tensor(torch.tensor(ids).unsqueeze(0), vals, size=(rows, width))


def _batches(seed, rows, width, steps, nnz):
  g = torch.Generator().manual_seed(seed)
  out = []
  for _ in range(steps):
    ids = torch.randint(0, rows, (nnz,), generator=g).tolist()
    ids[1] = ids[0]  # at least one duplicate
    out.append((ids, torch.randn(nnz, width, generator=g)))
  return out


@pytest.mark.parametrize("kind", ["sgd", "adagrad", "adam"])
def test_matches_torch_optim_on_touched_rows(kind):
  rows, width, lr = 40, 8, 0.1
  torch.manual_seed(0)
  w0 = torch.randn(rows, width)
  p = torch.nn.Parameter(w0.clone())
  opt = SparseRowOptimizer([p], kind, lr=lr)
  q = torch.nn.Parameter(w0.clone())
  if kind == "sgd":
    ref = torch.optim.SGD([q], lr=lr)
  elif kind == "adagrad":
    ref = torch.optim.Adagrad([q], lr=lr, initial_accumulator_value=0.1, eps=1e-7)
  else:
    ref = torch.optim.Adam([q], lr=lr, eps=1e-8)
  touched_ever = torch.zeros(rows, dtype=torch.bool)
  for step, (ids, vals) in enumerate(_batches(1, rows, width, 4, 12)):
    p.grad = _sparse_grad(rows, width, ids, vals)
    dense = p.grad.to_dense()
    touched = torch.zeros(rows, dtype=torch.bool)
    touched[ids] = True
    before = p.detach().clone()
    opt.step()
    assert p.grad is None
    # untouched rows do not move (lazy semantics)
    torch.testing.assert_close(p.detach()[~touched], before[~touched], rtol=0, atol=0)
    if kind == "adam" and step > 0:
      continue  # dense Adam decays the moments of untouched rows: only step 1 is comparable
    q.grad = dense
    prev = q.detach().clone()
    ref.step()
    with torch.no_grad():  # dense optimizers also move untouched rows only through state decay
      q[~touched] = prev[~touched]
    touched_ever |= touched
    torch.testing.assert_close(p.detach()[touched], q.detach()[touched], rtol=1e-5, atol=1e-6)


def test_lazy_adam_formula():
  rows, width, lr = 10, 4, 0.05
  w = np.random.default_rng(0).standard_normal((rows, width)).astype(np.float32)
  p = torch.nn.Parameter(torch.from_numpy(w.copy()))
  opt = SparseRowOptimizer([p], "adam", lr=lr)
  m = np.zeros_like(w)
  v = np.zeros_like(w)
  for t, (ids, vals) in enumerate(_batches(3, rows, width, 5, 6), start=1):
    p.grad = _sparse_grad(rows, width, ids, vals)
    g = p.grad.to_dense().numpy()
    opt.step()
    rowsel = np.unique(ids)
    m[rowsel] = 0.9 * m[rowsel] + 0.1 * g[rowsel]
    v[rowsel] = 0.999 * v[rowsel] + 0.001 * g[rowsel]**2
    w[rowsel] -= lr * (m[rowsel] / (1 - 0.9**t)) / (np.sqrt(v[rowsel] / (1 - 0.999**t)) + 1e-8)
    np.testing.assert_allclose(p.detach().numpy(), w, rtol=1e-5, atol=1e-6)


def test_rowwise_adagrad_formula_and_weight_decay():
  rows, width, lr, wd = 12, 8, 0.2, 0.01
  w = np.random.default_rng(1).standard_normal((rows, width)).astype(np.float32)
  p = torch.nn.Parameter(torch.from_numpy(w.copy()))
  opt = SparseRowOptimizer([p], "rowwise_adagrad", lr=lr, weight_decay=wd)
  acc = np.full((rows,), 0.1, dtype=np.float32)
  for ids, vals in _batches(4, rows, width, 3, 5):
    p.grad = _sparse_grad(rows, width, ids, vals)
    g = p.grad.to_dense().numpy()
    opt.step()
    rowsel = np.unique(ids)
    gg = g[rowsel] + wd * w[rowsel]
    acc[rowsel] += (gg * gg).mean(axis=1)
    w[rowsel] -= lr * gg / (np.sqrt(acc[rowsel])[:, None] + 1e-7)
    np.testing.assert_allclose(p.detach().numpy(), w, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(opt.state[0][0].numpy(), acc, rtol=1e-5, atol=1e-6)


def test_dense_gradients_frozen_params_and_bad_kind():
  a = torch.nn.Parameter(torch.ones(3, 2))
  b = torch.nn.Parameter(torch.ones(3, 2), requires_grad=False)
  opt = SparseRowOptimizer([a, b], "sgd", lr=0.5)
  assert len(opt.params) == 1
  a.grad = torch.full((3, 2), 2.0)
  opt.step()
  torch.testing.assert_close(a.detach(), torch.zeros(3, 2))
  opt.set_lr(0.25)
  a.grad = torch.full((3, 2), 4.0)
  opt.step()
  torch.testing.assert_close(a.detach(), torch.full((3, 2), -1.0))
  opt.step()  # no gradient: nothing happens
  torch.testing.assert_close(a.detach(), torch.full((3, 2), -1.0))
  with pytest.raises(ValueError):
    SparseRowOptimizer([a], "lamb")

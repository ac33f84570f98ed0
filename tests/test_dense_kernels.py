"""Dense-side kernels vs plain PyTorch fp32 references."""
import pytest
import torch

from distributed_embeddings_b200.ops import _native

pytestmark = pytest.mark.gpu


def ref_interact(bottom, emb, n_emb):
  b, d = bottom.shape
  f = torch.cat([bottom.unsqueeze(1), emb.view(b, n_emb, d)], 1).float()
  z = torch.bmm(f, f.transpose(1, 2))
  ii, jj = torch.tril_indices(n_emb + 1, n_emb + 1, -1)
  return torch.cat([z[:, ii, jj], bottom.float()], 1)


@pytest.mark.parametrize("n_emb,dim", [(26, 128), (7, 64), (31, 32), (3, 128)])
def test_interaction_fwd_bwd(n_emb, dim):
  ops = _native.require()
  torch.manual_seed(0)
  b = 777
  bottom = (torch.randn(b, dim, device="cuda") * 0.5).bfloat16()
  emb = (torch.randn(b, n_emb * dim, device="cuda") * 0.5).bfloat16()
  n_out = (n_emb + 1) * n_emb // 2 + dim
  zw = (n_out + 7) // 8 * 8
  z = torch.full((b, zw), 7.0, device="cuda", dtype=torch.bfloat16)
  ops.interact_fwd(bottom, emb, n_emb, z, [])
  bf = bottom.float().requires_grad_(True)
  ef = emb.float().requires_grad_(True)
  ref = ref_interact(bf, ef, n_emb)
  torch.testing.assert_close(z[:, :n_out].float(), ref, rtol=2e-2, atol=2e-2)
  assert torch.count_nonzero(z[:, n_out:]) == 0
  dz = (torch.randn(b, zw, device="cuda") * 0.1).bfloat16()
  ref.backward(dz[:, :n_out].float())
  dbottom = torch.empty(b, dim, device="cuda", dtype=torch.bfloat16)
  demb = torch.empty(b, n_emb * dim + 16, device="cuda", dtype=torch.bfloat16)
  ops.interact_bwd(bottom, emb, n_emb, dz, dbottom, demb.data_ptr(), demb.stride(0), 1.0, None, 0,
                   [], None, 0)
  torch.testing.assert_close(dbottom.float(), bf.grad, rtol=3e-2, atol=3e-2)
  torch.testing.assert_close(demb[:, :n_emb * dim].float(), ef.grad, rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("n_emb,dim,b", [(26, 128, 5003), (7, 64, 9001), (3, 128, 4099),
                                         (26, 128, 37)])
def test_interaction_bwd_double_buffered(n_emb, dim, b):
  """Batches with several samples per warp exercise the cp.async double buffer and the in-place
  shared-memory epilogue of the default interaction backward."""
  ops = _native.require()
  torch.manual_seed(1)
  bottom = (torch.randn(b, dim, device="cuda") * 0.5).bfloat16()
  emb = (torch.randn(b, n_emb * dim, device="cuda") * 0.5).bfloat16()
  n_out = (n_emb + 1) * n_emb // 2 + dim
  zw = (n_out + 7) // 8 * 8
  bf = bottom.float().requires_grad_(True)
  ef = emb.float().requires_grad_(True)
  ref = ref_interact(bf, ef, n_emb)
  dz = (torch.randn(b, zw, device="cuda") * 0.1).bfloat16()
  ref.backward(dz[:, :n_out].float())
  dbottom = torch.empty(b, dim, device="cuda", dtype=torch.bfloat16)
  demb = torch.empty(b, n_emb * dim + 16, device="cuda", dtype=torch.bfloat16)
  ops.interact_bwd(bottom, emb, n_emb, dz, dbottom, demb.data_ptr(), demb.stride(0), 1.0, None, 0,
                   [], None, 0)
  torch.cuda.synchronize()
  torch.testing.assert_close(dbottom.float(), bf.grad, rtol=3e-2, atol=3e-2)
  torch.testing.assert_close(demb[:, :n_emb * dim].float(), ef.grad, rtol=3e-2, atol=3e-2)


def test_interaction_bwd_routed():
  """Gradient pieces routed to separate destination buffers (what the distributed step does
  with the owners' receive buffers): column slices of a feature, permuted features, and a
  destination row stride different from the source."""
  import numpy as np
  from distributed_embeddings_b200.ops._native import GRAD_ROUTE, upload_struct_array
  ops = _native.require()
  torch.manual_seed(2)
  n_emb, dim, b = 5, 128, 3001
  bottom = (torch.randn(b, dim, device="cuda") * 0.5).bfloat16()
  emb = (torch.randn(b, n_emb * dim, device="cuda") * 0.5).bfloat16()
  n_out = (n_emb + 1) * n_emb // 2 + dim
  zw = (n_out + 7) // 8 * 8
  bf = bottom.float().requires_grad_(True)
  ef = emb.float().requires_grad_(True)
  ref = ref_interact(bf, ef, n_emb)
  dz = (torch.randn(b, zw, device="cuda") * 0.1).bfloat16()
  ref.backward(dz[:, :n_out].float())
  # "owner" A gets features 4, 0 and the upper half of feature 2; owner B the rest
  a = torch.zeros(b, 2 * 128 + 64 + 8, device="cuda", dtype=torch.bfloat16)
  bb = torch.zeros(b, 2 * 128 + 64, device="cuda", dtype=torch.bfloat16)
  pieces = [  # (src_col, width, dst, dst_col)
      (4 * 128, 128, a, 0), (0, 128, a, 128), (2 * 128 + 64, 64, a, 256),
      (1 * 128, 128, bb, 64), (2 * 128, 64, bb, 0), (3 * 128, 128, bb, 192)]
  routes = np.zeros(len(pieces), dtype=GRAD_ROUTE)
  for i, (sc, w, dst, dc) in enumerate(sorted(pieces)):
    routes[i]["dst"], routes[i]["dst_stride"] = dst.data_ptr(), dst.stride(0)
    routes[i]["src_col"], routes[i]["width"], routes[i]["dst_col"] = sc, w, dc
  dbottom = torch.empty(b, dim, device="cuda", dtype=torch.bfloat16)
  ops.interact_bwd(bottom, emb, n_emb, dz, dbottom, 0, 0, 0.5, upload_struct_array(routes, "cuda"),
                   len(pieces), [], None, 0)
  torch.cuda.synchronize()
  torch.testing.assert_close(dbottom.float(), bf.grad, rtol=3e-2, atol=3e-2)
  for sc, w, dst, dc in pieces:
    torch.testing.assert_close(dst[:, dc:dc + w].float(), 0.5 * ef.grad[:, sc:sc + w],
                               rtol=3e-2, atol=3e-2)
  assert torch.count_nonzero(a[:, 320:]) == 0


@pytest.mark.parametrize("cols", [128, 256, 512, 1024])
def test_relu_bwd_bias(cols):
  ops = _native.require()
  torch.manual_seed(1)
  rows = 3000
  y = torch.relu(torch.randn(rows, cols, device="cuda")).bfloat16()
  dy = torch.randn(rows, cols, device="cuda").bfloat16()
  want = dy.float() * (y > 0)
  db = torch.zeros(cols, device="cuda")
  ops.relu_bwd_bias(dy, y, db)
  torch.testing.assert_close(dy.float(), want.bfloat16().float())
  torch.testing.assert_close(db, want.sum(0), rtol=1e-3, atol=1e-2)


def test_head_loss():
  ops = _native.require()
  torch.manual_seed(2)
  b, k = 1000, 256
  x = torch.relu(torch.randn(b, k, device="cuda")).bfloat16()
  w = (torch.randn(k, device="cuda") * 0.1).bfloat16()
  bias = torch.tensor([0.3], device="cuda").bfloat16()
  labels = torch.randint(0, 2, (b,), device="cuda").float()
  xf = x.float().requires_grad_(True)
  wf = w.float().requires_grad_(True)
  bf = bias.float().requires_grad_(True)
  logit = xf @ wf + bf
  loss = torch.nn.functional.binary_cross_entropy_with_logits(logit, labels)
  loss.backward()
  dx = torch.empty_like(x)
  dw = torch.zeros(k, device="cuda")
  db = torch.zeros(8, device="cuda")
  dbp = torch.zeros(k, device="cuda")
  ls = torch.zeros(1, device="cuda")
  logits = torch.empty(b, device="cuda")
  ops.head_loss(x, w, bias, labels, 1.0 / b, dx, dw, db, dbp, ls, logits)
  torch.testing.assert_close(ls[0], loss.detach(), rtol=1e-4, atol=1e-5)
  torch.testing.assert_close(logits, logit.detach(), rtol=1e-4, atol=1e-4)
  want_dx = xf.grad * (x.float() > 0)
  torch.testing.assert_close(dx.float(), want_dx, rtol=2e-2, atol=1e-6)
  torch.testing.assert_close(dw, wf.grad, rtol=1e-3, atol=1e-5)
  torch.testing.assert_close(db[0], bf.grad[0], rtol=1e-3, atol=1e-6)
  torch.testing.assert_close(dbp, want_dx.sum(0), rtol=1e-3, atol=1e-5)


def test_dense_sgd_and_cast_pad():
  ops = _native.require()
  torch.manual_seed(3)
  n = 4096 + 64
  p = torch.randn(n, device="cuda")
  g = torch.randn(n, device="cuda")
  p16 = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
  lr = torch.tensor([0.5], device="cuda")
  want = p - 0.5 * g
  ops.dense_sgd(p, p16, g, lr, 1.0)
  torch.testing.assert_close(p, want)
  torch.testing.assert_close(p16.float(), want.bfloat16().float())
  assert torch.count_nonzero(g) == 0
  src = torch.randn(100, 13, device="cuda")
  dst = torch.ones(100, 16, device="cuda", dtype=torch.bfloat16)
  ops.cast_pad(src, dst)
  torch.testing.assert_close(dst[:, :13].float(), src.bfloat16().float())
  assert torch.count_nonzero(dst[:, 13:]) == 0

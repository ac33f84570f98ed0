"""Fused in-backward optimizers of the fused engine vs explicit PyTorch math, on heavily skewed
(power-law) ids so that segments span many chunks of the occurrence-balanced update."""
import pytest
import torch

import distributed_embeddings_b200 as de

pytestmark = pytest.mark.gpu


def reference_update(kind, w, state, ids_list, grad_cols, lr, eps, init_acc, scale, combiner,
                     hots):
  """Dense reference: accumulate the gradient of every row, then apply the optimizer to touched
  rows only."""
  g = torch.zeros_like(w)
  for ids, gcol, h in zip(ids_list, grad_cols, hots):
    wgt = 1.0 / h if combiner == "mean" else 1.0
    g.index_add_(0, ids.reshape(-1), gcol.repeat_interleave(h, dim=0) * wgt)
  g *= scale
  touched = torch.zeros(w.shape[0], dtype=torch.bool, device=w.device)
  for ids in ids_list:
    touched[ids.reshape(-1)] = True
  w = w.clone()
  if kind == "sgd":
    w -= lr * g
  elif kind == "adagrad":
    acc = state[0]
    acc[touched] += g[touched]**2
    w[touched] -= lr * g[touched] / (acc[touched].sqrt() + eps)
  elif kind == "rowwise_adagrad":
    acc = state[0]
    acc[touched] += (g[touched]**2).mean(dim=1)
    w[touched] -= lr * g[touched] / (acc[touched].sqrt().unsqueeze(1) + eps)
  elif kind == "adam":
    m, v = state
    b1, b2 = 0.9, 0.999
    m[touched] = b1 * m[touched] + (1 - b1) * g[touched]
    v[touched] = b2 * v[touched] + (1 - b2) * g[touched]**2
    mh, vh = m[touched] / (1 - b1), v[touched] / (1 - b2)
    w[touched] -= lr * mh / (vh.sqrt() + 1e-8)
  return w


@pytest.mark.parametrize("kind", ["sgd", "adagrad", "rowwise_adagrad", "adam"])
@pytest.mark.parametrize("width", [16, 32, 128])
@pytest.mark.parametrize("combiner", ["sum", "mean"])
def test_fused_optimizer_on_skewed_ids(kind, width, combiner):
  dev = torch.device("cuda", 0)
  torch.manual_seed(0)
  rows, b = 500, 4096
  hots = [1, 7, 3]
  tables = [de.Embedding(rows, width, combiner=combiner, device=dev) for _ in range(2)]
  demb = de.DistributedEmbedding(tables, input_table_map=[0, 1, 0], device=dev, backend="fused",
                                 world_size=1, rank=0)
  kw = {"deterministic": True} if kind == "sgd" else {}
  lr = 0.1
  demb.set_optimizer(kind, lr=lr, **kw)
  g = torch.Generator().manual_seed(1)
  # power law: most ids are 0/1/2, a long tail elsewhere
  ids = [(torch.rand(b, h, generator=g).pow(6) * rows).long().clamp(max=rows - 1).to(dev)
         for h in hots]
  w_before = demb.get_weights()
  out = demb(ids, concat=True)
  grad = torch.randn(out.shape, generator=g).to(dev)
  out.backward(grad)
  torch.cuda.synchronize()
  w_after = demb.get_weights()
  # both tables are fused into one local table; emulate per global table
  cols = [grad[:, 0:width], grad[:, width:2 * width], grad[:, 2 * width:3 * width]]
  eps = 1e-8 if kind == "adam" else 1e-7
  for t in range(2):
    sel = [i for i in range(3) if [0, 1, 0][i] == t]
    w0 = torch.from_numpy(w_before[t]).to(dev)
    if kind in ("adagrad",):
      state = [torch.full_like(w0, 0.1)]
    elif kind == "rowwise_adagrad":
      state = [torch.full((rows,), 0.1, device=dev)]
    elif kind == "adam":
      state = [torch.zeros_like(w0), torch.zeros_like(w0)]
    else:
      state = []
    want = reference_update(kind, w0, state, [ids[i] for i in sel], [cols[i] for i in sel], lr,
                            eps, 0.1, 1.0, combiner, [hots[i] for i in sel])
    torch.testing.assert_close(torch.from_numpy(w_after[t]).to(dev), want, rtol=2e-4, atol=2e-4)

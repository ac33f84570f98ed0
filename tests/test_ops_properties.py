"""Property-based checks (hypothesis) of the CPU paths of the op layer: pooled lookup forward and
its deduplicated sparse gradient against naive loops, and IntegerLookup against a dict model."""
import numpy as np
import pytest
import torch

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

import distributed_embeddings_b200 as de  # noqa: E402
from distributed_embeddings_b200.ops.ragged import RaggedIds  # noqa: E402


@st.composite
def ragged_cases(draw):
  voc = draw(st.integers(1, 30))
  width = draw(st.sampled_from([1, 3, 4, 8, 33]))
  batch = draw(st.integers(1, 12))
  lens = draw(st.lists(st.integers(0, 6), min_size=batch, max_size=batch))
  vals = [draw(st.integers(0, voc - 1)) for _ in range(sum(lens))]
  combiner = draw(st.sampled_from(["sum", "mean"]))
  seed = draw(st.integers(0, 1000))
  return voc, width, lens, vals, combiner, seed


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(ragged_cases())
def test_ragged_lookup_and_sparse_gradient(c):
  voc, width, lens, vals, combiner, seed = c
  g = torch.Generator().manual_seed(seed)
  param = torch.randn(voc, width, generator=g, requires_grad=True)
  ids = RaggedIds.from_row_lengths(torch.tensor(vals, dtype=torch.int64),
                                   torch.tensor(lens, dtype=torch.int64))
  out = de.embedding_lookup(param, ids, combiner=combiner)
  # naive oracle; an empty row pools to zero
  ref = torch.zeros(len(lens), width)
  pos = 0
  for r, n in enumerate(lens):
    if n:
      rows = param.detach()[vals[pos:pos + n]]
      ref[r] = rows.sum(0) if combiner == "sum" else rows.mean(0)
    pos += n
  torch.testing.assert_close(out.detach(), ref, rtol=1e-5, atol=1e-5)
  up = torch.randn(len(lens), width, generator=g)
  out.backward(up)
  grad = param.grad
  dense = grad.to_dense() if grad.is_sparse else grad
  exp = torch.zeros(voc, width)
  pos = 0
  for r, n in enumerate(lens):
    for v in vals[pos:pos + n]:
      exp[v] += up[r] / (n if combiner == "mean" else 1)
    pos += n
  torch.testing.assert_close(dense, exp, rtol=1e-5, atol=1e-5)
  if grad.is_sparse:
    # reference semantics: one entry per *unique* id (IndexedSlices of unique ids)
    idx = grad.coalesce().indices()[0] if not grad.is_coalesced() else grad.indices()[0]
    assert idx.numel() == len(set(vals))


@settings(max_examples=100, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.integers(1, 12), st.lists(st.lists(st.integers(-50, 50), min_size=1, max_size=9),
                                    min_size=1, max_size=6))
def test_integer_lookup_matches_dict_model(max_tokens, batches):
  layer = de.IntegerLookup(max_tokens, device="cpu")
  model = {}
  counts = {}
  for keys in batches:
    out = layer(torch.tensor(keys, dtype=torch.int64)).tolist()
    for k, o in zip(keys, out):
      if k == -1:
        assert o == 0  # the empty-slot marker is never a vocabulary entry
      elif k in model:
        assert o == model[k]
      elif len(model) < max_tokens:
        # a new key gets the next free index, whatever its position among this batch's new keys
        assert 1 <= o <= max_tokens and o not in model.values()
        model[k] = o
      else:
        assert o == 0
      if o:
        counts[o] = counts.get(o, 0) + 1
  vocab = layer.get_vocabulary()
  assert len(vocab) == len(model) + 1
  for k, i in model.items():
    assert vocab[i] == k


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.lists(st.integers(1, 5), min_size=1, max_size=4), st.integers(1, 9),
       st.sampled_from([None, "sum", "mean"]), st.integers(0, 100))
def test_embedding_layer_nd_inputs(shape, width, combiner, seed):
  """N-D dense inputs: without a combiner the output is shape + (width,); with one the last
  axis is pooled (reference embedding.py:120-147)."""
  voc = 13
  g = torch.Generator().manual_seed(seed)
  layer = de.Embedding(voc, width, combiner=combiner, device="cpu")
  ids = torch.randint(0, voc, tuple(shape), generator=g)
  if combiner is not None and len(shape) < 2:
    with pytest.raises(ValueError):
      layer(ids)
    return
  out = layer(ids)
  w = layer.embeddings.detach()
  ref = w[ids]
  if combiner == "sum":
    ref = ref.sum(-2)
  elif combiner == "mean":
    ref = ref.mean(-2)
  assert out.shape == ref.shape
  torch.testing.assert_close(out.detach(), ref, rtol=1e-5, atol=1e-5)


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.lists(st.integers(1, 20), min_size=1, max_size=6), st.integers(1, 8), st.integers(1, 7),
       st.integers(0, 100))
def test_concat_one_hot_embedding(sizes, width, batch, seed):
  g = torch.Generator().manual_seed(seed)
  layer = de.ConcatOneHotEmbedding(sizes, width, device="cpu")
  ids = torch.stack([torch.randint(0, s, (batch,), generator=g) for s in sizes], dim=1)
  out = layer(ids)
  assert out.shape == (batch, len(sizes), width)
  offs = np.concatenate([[0], np.cumsum(sizes)[:-1]])
  ref = layer.params.detach()[ids + torch.from_numpy(offs)]
  torch.testing.assert_close(out.detach(), ref)

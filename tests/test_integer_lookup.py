"""IntegerLookup (reference integer_lookup_test.py:29-70)."""
import pytest
import torch

import distributed_embeddings_b200 as de

DEVICES = ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("device", DEVICES)
@pytest.mark.parametrize("use_gpu", [True, False])
def test_small_vocab_with_oov(device, use_gpu):
  layer = de.IntegerLookup(4, use_gpu=use_gpu, device=device)
  data = torch.tensor([[12, 1138, 42], [42, 1000, 36], [7, 8, 9]], device=device)
  out = layer(data)
  vocab = layer.get_vocabulary()
  assert vocab[0] == -1 and len(vocab) == 5
  lut = {k: i for i, k in enumerate(vocab)}
  expect = [[lut.get(int(k), 0) for k in row] for row in data.tolist()]
  assert out.tolist() == expect
  assert sorted(vocab[1:]) == sorted({12, 1138, 42, 1000})
  # full: later keys are OOV, known keys keep their index
  again = layer(torch.tensor([36, 42, 5], device=device))
  assert again.tolist() == [0, lut[42], 0]


@pytest.mark.parametrize("device", DEVICES)
def test_random_vocab_fills_exactly(device):
  gen = torch.Generator().manual_seed(0)
  for _ in range(3):
    size = int(torch.randint(100, 1000, (1,), generator=gen))
    layer = de.IntegerLookup(size, device=device)
    keys = torch.randint(0, 1024, (4096,), generator=gen).to(device)
    out = layer(keys)
    vocab = layer.get_vocabulary()
    assert len(vocab) == min(size, len(set(keys.tolist()))) + 1
    assert len(set(vocab)) == len(vocab)
    lut = {k: i for i, k in enumerate(vocab)}
    allk = torch.arange(1024, device=device)
    got = layer(allk)
    if len(vocab) - 1 == size:  # table full: pure look-ups
      assert got.tolist() == [lut.get(k, 0) for k in range(1024)]
    assert int(out.max()) <= size and int(out.min()) >= 0
    # duplicates of one key inside a batch map to one index
    for k in keys[:50].tolist():
      vals = out[keys == k]
      assert len(set(vals.tolist())) == 1


@pytest.mark.parametrize("device", DEVICES)
def test_frequency_counts(device):
  layer = de.IntegerLookup(8, device=device)
  keys = torch.tensor([5, 5, 5, 9, 9, 2], device=device)
  out = layer(keys)
  counts = layer.count.tolist()
  for k, n in ((5, 3), (9, 2), (2, 1)):
    idx = int(out[keys == k][0])
    assert counts[idx] == n
  assert counts[0] == 1  # reserved OOV slot marker


@pytest.mark.gpu
def test_state_dict_roundtrip():
  a = de.IntegerLookup(16, device="cuda")
  keys = torch.arange(100, 110, device="cuda")
  first = a(keys)
  b = de.IntegerLookup(16, device="cuda")
  b.load_state_dict(a.state_dict())
  assert torch.equal(b(keys), first)
  assert b.get_vocabulary() == a.get_vocabulary()

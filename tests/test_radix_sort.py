"""First-party radix sort / head compaction (ops/csrc/radix_sort.cu) against torch.sort(stable)
and torch.unique_consecutive (the reference takes both from CUB, embedding_lookup_kernels.cu:645-661)."""
import pytest
import torch

from distributed_embeddings_b200.ops import _native

pytestmark = pytest.mark.gpu


def _keys(n, bits, skew, gen):
  hi = 1 << bits
  if skew:
    # power-law-ish: many repeats of a few keys
    u = torch.rand(n, generator=gen, device="cuda")
    return (u.pow(6) * hi).long().clamp_(max=hi - 1)
  return torch.randint(0, hi, (n,), generator=gen, device="cuda", dtype=torch.int64)


@pytest.mark.parametrize("n", [1, 31, 4096, 4097, 100003, 1 << 20])
@pytest.mark.parametrize("bits", [1, 8, 13, 28, 40])
@pytest.mark.parametrize("skew", [False, True])
def test_radix_sort_pairs_stable(n, bits, skew):
  ops = _native.require()
  gen = torch.Generator(device="cuda").manual_seed(n * 131 + bits)
  keys = _keys(n, bits, skew, gen)
  items = torch.arange(n, device="cuda", dtype=torch.int32)
  ks, it = ops.radix_sort_pairs(keys, items, bits)
  ref_k, ref_i = torch.sort(keys, stable=True)
  assert torch.equal(ks, ref_k)
  assert torch.equal(it.long(), ref_i)
  # inputs are not clobbered by the standalone op
  assert torch.equal(items, torch.arange(n, device="cuda", dtype=torch.int32))


@pytest.mark.parametrize("n", [1, 31, 4097, 100003, 1 << 20])
@pytest.mark.parametrize("bits", [1, 8, 13, 28, 32])
@pytest.mark.parametrize("skew", [False, True])
def test_radix_sort_pairs32_stable(n, bits, skew):
  """32-bit key variant (keys < 2^32 travel as uint32, the last pass widens them to int64)."""
  ops = _native.require()
  gen = torch.Generator(device="cuda").manual_seed(n * 17 + bits)
  keys = _keys(n, bits, skew, gen)
  k32 = (keys & 0xFFFFFFFF).to(torch.int64)
  k32 = torch.where(k32 >= 2**31, k32 - 2**32, k32).to(torch.int32)  # same bits, signed storage
  items = torch.arange(n, device="cuda", dtype=torch.int32)
  ks, it = ops.radix_sort_pairs32(k32, items, bits)
  ref_k, ref_i = torch.sort(keys, stable=True)
  assert ks.dtype == torch.int64 and torch.equal(ks, ref_k)
  assert torch.equal(it.long(), ref_i)


@pytest.mark.parametrize("n", [1, 33, 4096, 4097, 250001])
@pytest.mark.parametrize("distinct", [1, 7, 5000, 1 << 30])
def test_head_segments(n, distinct):
  ops = _native.require()
  gen = torch.Generator(device="cuda").manual_seed(n + distinct)
  keys = torch.randint(0, distinct, (n,), generator=gen, device="cuda", dtype=torch.int64)
  keys = torch.sort(keys).values
  seg, nu = ops.head_segments(keys)
  uniq, counts = torch.unique_consecutive(keys, return_counts=True)
  u = int(nu.item())
  assert u == uniq.numel()
  starts = torch.cumsum(counts, 0) - counts
  assert torch.equal(seg[:u], starts)
  assert int(seg[u].item()) == n


def test_update_path_with_library_sort():
  """The first-party sort is the default of the deduplicated optimizer path (covered by
  test_fused_optimizers); DE_B200_SORT=cub keeps the CUB route for A/B runs.  The switch is read
  once per process, hence the subprocess."""
  import os
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  env = dict(os.environ, DE_B200_SORT="cub")
  out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests",
                                                                      "test_fused_optimizers.py"),
                        "-x", "-q", "-k", "128 and sum"], env=env, cwd=root, capture_output=True,
                       text=True, timeout=600, check=False)
  assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
  assert " passed" in out.stdout

"""Minimal ragged / sparse id containers.

The reference takes ``tf.RaggedTensor`` (values + row_splits = CSR) and ``tf.SparseTensor``
(COO) id inputs (embedding_lookup_ops.py:68-96).  PyTorch has no direct equivalents, so the
framework defines two tiny value types that only carry what the kernels consume.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch


@dataclass
class RaggedIds:
  """2-D ragged batch of ids in CSR form: row ``i`` owns ``values[row_splits[i]:row_splits[i+1]]``."""
  values: torch.Tensor
  row_splits: torch.Tensor

  @property
  def nrows(self) -> int:
    return self.row_splits.numel() - 1

  @property
  def shape(self) -> Tuple[int, None]:
    return (self.nrows, None)

  @property
  def device(self):
    return self.values.device

  @property
  def dtype(self):
    return self.values.dtype

  def row_lengths(self) -> torch.Tensor:
    return self.row_splits[1:] - self.row_splits[:-1]

  def to(self, *args, **kwargs) -> "RaggedIds":
    return RaggedIds(self.values.to(*args, **kwargs), self.row_splits.to(*args, **kwargs))

  @staticmethod
  def from_row_lengths(values: torch.Tensor, row_lengths: torch.Tensor) -> "RaggedIds":
    splits = torch.zeros(row_lengths.numel() + 1, dtype=torch.int64, device=values.device)
    torch.cumsum(row_lengths.to(torch.int64), 0, out=splits[1:])
    return RaggedIds(values, splits)

  @staticmethod
  def from_dense(ids: torch.Tensor) -> "RaggedIds":
    b, h = ids.shape
    splits = torch.arange(0, (b + 1) * h, h, dtype=torch.int64, device=ids.device)
    return RaggedIds(ids.reshape(-1), splits)

  @staticmethod
  def from_lists(rows, dtype=torch.int64, device=None) -> "RaggedIds":
    lengths = torch.tensor([len(r) for r in rows], dtype=torch.int64, device=device)
    flat = [x for r in rows for x in r]
    return RaggedIds.from_row_lengths(torch.tensor(flat, dtype=dtype, device=device), lengths)

  def slice_rows(self, start: int, end: int) -> "RaggedIds":
    lo = int(self.row_splits[start])
    hi = int(self.row_splits[end])
    return RaggedIds(self.values[lo:hi], self.row_splits[start:end + 1] - lo)

  def to_lists(self):
    s = self.row_splits.tolist()
    v = self.values.tolist()
    return [v[s[i]:s[i + 1]] for i in range(len(s) - 1)]


@dataclass
class SparseIds:
  """2-D COO batch of ids: ``indices[k] = (row, col)`` sorted by row, ``values[k]`` the id."""
  indices: torch.Tensor  # [nnz, 2]
  values: torch.Tensor  # [nnz]
  dense_shape: Tuple[int, int]

  @property
  def shape(self):
    return tuple(self.dense_shape)

  @property
  def device(self):
    return self.values.device

  @property
  def dtype(self):
    return self.values.dtype

  @staticmethod
  def from_torch_sparse(t: torch.Tensor) -> "SparseIds":
    t = t.coalesce()
    return SparseIds(t.indices().t().contiguous(), t.values(), tuple(t.shape))

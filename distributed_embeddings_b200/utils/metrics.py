"""Evaluation metrics and structured logging helpers."""
from __future__ import annotations

import json
import time
from typing import Any, Dict

import torch


def binary_auc(labels: torch.Tensor, scores: torch.Tensor) -> float:
  """Exact ROC AUC by rank statistics (ties get the average rank).

  The reference example uses a Keras AUC metric with 8000 thresholds (examples/dlrm/main.py:
  223-243); the exact value is cheaper and deterministic."""
  labels = labels.reshape(-1).double()
  scores = scores.reshape(-1).double()
  n_pos = float(labels.sum())
  n_neg = float(labels.numel() - n_pos)
  if n_pos == 0 or n_neg == 0:
    return float("nan")
  order = torch.argsort(scores)
  s = scores[order]
  ranks = torch.arange(1, s.numel() + 1, dtype=torch.float64)
  # average ranks over ties
  uniq, inv, counts = torch.unique_consecutive(s, return_inverse=True, return_counts=True)
  ends = torch.cumsum(counts, 0).double()
  starts = ends - counts.double() + 1
  avg = ((starts + ends) / 2)[inv]
  pos_rank_sum = float(avg[labels[order] > 0.5].sum())
  return (pos_rank_sum - n_pos * (n_pos + 1) / 2) / (n_pos * n_neg)


class MetricsLogger:
  """Structured (JSON lines) metrics: one record per call, rank-0 only by default."""

  def __init__(self, path=None, rank: int = 0, enabled_ranks=(0,)):
    self.path, self.rank, self.enabled = path, rank, rank in enabled_ranks
    self.t0 = time.time()

  def log(self, **record: Any) -> Dict[str, Any]:
    record = {"t": round(time.time() - self.t0, 3), "rank": self.rank, **record}
    if self.enabled:
      line = json.dumps(record)
      if self.path:
        with open(self.path, "a", encoding="utf-8") as f:
          f.write(line + "\n")
      else:
        print(line, flush=True)
    return record


def bus_bandwidth_gbs(bytes_per_rank_out: float, seconds: float) -> float:
  """All-to-all bus bandwidth per GPU (bytes leaving one GPU / time)."""
  return bytes_per_rank_out / max(seconds, 1e-12) / 1e9

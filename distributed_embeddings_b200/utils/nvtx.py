"""NVTX ranges around the phases of a step (enable with DE_B200_NVTX=1; no-ops otherwise).

The reference has no tracing hooks at all; these ranges show up in Nsight
Systems / Compute timelines as `de:<phase>`."""
from __future__ import annotations

import contextlib
import os

import torch

ENABLED = os.environ.get("DE_B200_NVTX", "0") == "1"


@contextlib.contextmanager
def range(name: str):  # pylint: disable=redefined-builtin
  if ENABLED and torch.cuda.is_available():
    torch.cuda.nvtx.range_push("de:" + name)
    try:
      yield
    finally:
      torch.cuda.nvtx.range_pop()
  else:
    yield

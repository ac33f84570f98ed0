"""Host-side contract of the hand-scheduled synthetic step that does not need a GPU: which models
qualify (``SyntheticTrainStep.unsupported_reason``), i.e. when ``--trainer auto`` of
examples/benchmarks/synthetic_models/main.py falls back to the autograd trainer."""
import torch
from torch import nn

from distributed_embeddings_b200.models.configs import scaled, synthetic_models_v3
from distributed_embeddings_b200.models.synthetic import SyntheticModel, SyntheticModelNative
from distributed_embeddings_b200.models.synthetic_fast import SyntheticTrainStep


def _tiny(**kw):
  cfg = scaled(synthetic_models_v3["tiny"], 1e-4)
  return SyntheticModel(cfg, device=torch.device("cpu"), **kw)


def test_reasons():
  reason = SyntheticTrainStep.unsupported_reason
  assert "SyntheticModel" in reason(nn.Linear(2, 2))
  native = SyntheticModelNative(scaled(synthetic_models_v3["tiny"], 1e-4),
                                device=torch.device("cpu"))
  assert "SyntheticModel" in reason(native)
  m = _tiny(backend="torch", compute_dtype=torch.bfloat16)
  assert "fused" in reason(m)
  # pretend the engine is available: the remaining checks only look at the module
  m.embedding.backend = "fused"
  assert reason(m) is None
  m.compute_dtype = torch.float32
  assert "bf16" in reason(m)
  m.compute_dtype = torch.bfloat16
  lins = [l for l in m.mlp if isinstance(l, nn.Linear)]
  lins[-2].out_features = 96
  assert "last hidden layer" in reason(m)
  lins[-2].out_features = 256
  lins[0].out_features = 100
  assert "multiples of 8" in reason(m)


def test_constructor_rejects_unsupported_model():
  m = _tiny(backend="torch", compute_dtype=torch.bfloat16)
  try:
    SyntheticTrainStep(m)
  except (ValueError, RuntimeError) as e:
    assert "fused" in str(e)
  else:
    raise AssertionError("expected the constructor to refuse a torch-back-end model")

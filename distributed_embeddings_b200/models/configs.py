"""Synthetic recommender model zoo (tiny ... colossal + criteo).

Same model sizes as the reference's benchmark suite (examples/benchmarks/synthetic_models/
config_v3.py:30-142): every ``EmbeddingConfig`` describes ``num_tables`` tables of ``num_rows`` x
``width``; ``nnz`` lists the hotness of each input reading the table (shared = one table serves
all of them).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional


@dataclass(frozen=True)
class EmbeddingConfig:
  num_tables: int
  nnz: tuple
  num_rows: int
  width: int
  shared: bool


@dataclass(frozen=True)
class ModelConfig:
  name: str
  embedding_configs: tuple
  mlp_sizes: tuple
  num_numerical_features: int
  interact_stride: Optional[int]


def _E(n, nnz, rows, width, shared):
  return EmbeddingConfig(n, tuple(nnz), rows, width, shared)


model_tiny = ModelConfig(
    "Tiny V3",
    (_E(1, [1, 10], 10000, 8, True), _E(1, [1, 10], 1000000, 16, True),
     _E(1, [1, 10], 25000000, 16, True), _E(1, [1], 25000000, 16, False), _E(16, [1], 10, 8, False),
     _E(10, [1], 1000, 8, False), _E(4, [1], 10000, 8, False), _E(2, [1], 100000, 16, False),
     _E(19, [1], 1000000, 16, False)), (256, 128), 10, None)

model_small = ModelConfig(
    "Small V3",
    (_E(5, [1, 30], 10000, 16, True), _E(3, [1, 30], 4000000, 32, True),
     _E(1, [1, 30], 50000000, 32, True), _E(1, [1], 50000000, 32, False),
     _E(30, [1], 10, 16, False), _E(30, [1], 1000, 16, False), _E(5, [1], 10000, 16, False),
     _E(5, [1], 100000, 32, False), _E(27, [1], 4000000, 32, False)), (512, 256, 128), 10, None)

model_medium = ModelConfig(
    "Medium v3",
    (_E(20, [1, 50], 100000, 64, True), _E(5, [1, 50], 10000000, 64, True),
     _E(1, [1, 50], 100000000, 128, True), _E(1, [1], 100000000, 128, False),
     _E(80, [1], 10, 32, False), _E(60, [1], 1000, 32, False), _E(80, [1], 100000, 64, False),
     _E(24, [1], 200000, 64, False), _E(40, [1], 10000000, 64, False)), (1024, 512, 256, 128), 25,
    7)

model_large = ModelConfig(
    "Large v3",
    (_E(40, [1, 100], 100000, 64, True), _E(16, [1, 100], 15000000, 64, True),
     _E(1, [1, 100], 200000000, 128, True), _E(1, [1], 200000000, 128, False),
     _E(100, [1], 10, 32, False), _E(100, [1], 10000, 32, False), _E(160, [1], 100000, 64, False),
     _E(50, [1], 500000, 64, False), _E(144, [1], 15000000, 64, False)), (2048, 1024, 512, 256),
    100, 8)

model_jumbo = ModelConfig(
    "Jumbo v3",
    (_E(50, [1, 200], 100000, 128, True), _E(24, [1, 200], 20000000, 128, True),
     _E(1, [1, 200], 400000000, 256, True), _E(1, [1], 400000000, 256, False),
     _E(100, [1], 10, 32, False), _E(200, [1], 10000, 64, False), _E(350, [1], 100000, 128, False),
     _E(80, [1], 1000000, 128, False), _E(216, [1], 20000000, 128, False)),
    (2048, 1024, 512, 256), 200, 20)

model_colossal = ModelConfig(
    "Colossal v3",
    (_E(100, [1, 300], 100000, 128, True), _E(50, [1, 300], 40000000, 256, True),
     _E(1, [1, 300], 2000000000, 256, True), _E(1, [1], 1000000000, 256, False),
     _E(100, [1], 10, 32, False), _E(400, [1], 10000, 128, False), _E(100, [1], 100000, 128, False),
     _E(800, [1], 1000000, 128, False), _E(450, [1], 40000000, 256, False)),
    (4096, 2048, 1024, 512, 256), 500, 30)

model_criteo = ModelConfig("Criteo-dlrm-like", (_E(26, [1], 100000, 128, False),), (512, 256, 128),
                           13, None)

synthetic_models_v3 = {
    "criteo": model_criteo,
    "tiny": model_tiny,
    "small": model_small,
    "medium": model_medium,
    "large": model_large,
    "jumbo": model_jumbo,
    "colossal": model_colossal,
}


def scaled(config: ModelConfig, row_scale: float) -> ModelConfig:
  """Shrink the row counts (for tests / smoke runs on small machines)."""
  embs = tuple(
      EmbeddingConfig(e.num_tables, e.nnz, max(4, int(e.num_rows * row_scale)), e.width, e.shared)
      for e in config.embedding_configs)
  return ModelConfig(config.name + f" x{row_scale}", embs, config.mlp_sizes,
                     config.num_numerical_features, config.interact_stride)


def expand(config: ModelConfig):
  """-> (table shapes [(rows, width)], input_table_map, hotness per input)."""
  tables, imap, hot = [], [], []
  for e in config.embedding_configs:
    if len(e.nnz) > 1 and not e.shared:
      raise NotImplementedError("Nonshared multihot embedding is not implemented yet")
    for _ in range(e.num_tables):
      for h in e.nnz:
        imap.append(len(tables))
        hot.append(h)
      tables.append((e.num_rows, e.width))
  return tables, imap, hot


def summary(config: ModelConfig) -> dict:
  tables, imap, hot = expand(config)
  elems = sum(r * w for r, w in tables)
  return {"tables": len(tables), "inputs": len(imap), "rows": sum(r for r, _ in tables),
          "elements": elems, "gib_fp32": elems * 4 / 2**30,
          "output_width": sum(tables[t][1] for t in imap), "lookups_per_sample": sum(hot)}

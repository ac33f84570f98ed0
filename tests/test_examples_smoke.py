"""The example scripts run end to end on the CPU with tiny configurations (reference examples:
examples/dlrm/main.py, examples/criteo/main.py, examples/benchmarks/*)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, timeout=600):
  env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
  out = subprocess.run([sys.executable] + args, cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=timeout, check=False)
  assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
  return out.stdout


def test_dlrm_example(tmp_path):
  save = str(tmp_path / "weights")
  out = run(["examples/dlrm/main.py", "--batch_size", "64", "--num_batches", "3", "--table_sizes",
             "50,60,70", "--embedding_dim", "16", "--bottom_mlp_dims", "32,16", "--top_mlp_dims",
             "32,1", "--save_path", save])
  assert "loss" in out and "Evaluation completed" in out
  saved = np.load(save + ".npz")
  assert [saved[k].shape for k in saved.files] == [(50, 16), (60, 16), (70, 16)]


def test_criteo_integer_lookup_example():
  out = run(["examples/criteo/main.py", "--batch_size", "64", "--steps", "2", "--vocab", "100"])
  assert "vocab sizes" in out


@pytest.mark.parametrize("api", ["de", "native"])
def test_synthetic_benchmark_example(api):
  out = run(["examples/benchmarks/synthetic_models/main.py", "--model", "tiny", "--row_scale",
             "0.001", "--batch_size", "16", "--num_steps", "3", "--device", "cpu",
             "--optimizer", "adagrad", "--embedding_api", api] +
            (["--dp_input"] if api == "native" else []))
  rec = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
  assert rec["model"] == "tiny" and rec["tables"] == 55 and rec["samples_per_sec"] > 0


def test_op_microbenchmark_example():
  out = run(["examples/benchmarks/benchmark.py", "--device", "cpu", "--voc", "1000", "--dim", "16",
             "--batch", "64", "--max_hot", "5"])
  assert "custom fwd" in out and "EmbeddingBag fwd+bwd+sgd" in out

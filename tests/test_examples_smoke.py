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


def test_dlrm_example_on_a_split_binary_dataset(tmp_path):
  """The real-data path of the DLRM example: split-binary Criteo layout (generated), prefetching
  reader, training, AUC evaluation, parallel file checkpoint."""
  data = str(tmp_path / "criteo")
  run(["tools/make_synthetic_criteo.py", data, "--train", "2048", "--test", "512", "--table_sizes",
       "5,300,70000,40", "--num_numerical", "13"])
  ckpt = str(tmp_path / "ckpt")
  out = run(["examples/dlrm/main.py", "--dataset_path", data, "--batch_size", "256",
             "--embedding_dim", "16", "--bottom_mlp_dims", "32,16", "--top_mlp_dims", "32,1",
             "--learning_rate", "1.0", "--save_dir", ckpt])
  assert "Evaluation completed" in out
  auc = float(out.split("AUC:")[1].split(",")[0])
  assert 0.0 <= auc <= 1.0
  shapes = [np.load(os.path.join(ckpt, f"table_{t}.npy")).shape for t in range(4)]
  assert shapes == [(5, 16), (300, 16), (70000, 16), (40, 16)]


@pytest.mark.parametrize("world", [1, 2])
def test_dlrm_example_learns(tmp_path, world):
  """End-to-end convergence, not just one-step equality: three epochs over a generated dataset
  whose labels follow a logistic model of the features lift the evaluation AUC from 0.5 to well
  above 0.7 (single process, and two gloo ranks with model-parallel tables)."""
  data = str(tmp_path / "criteo")
  run(["tools/make_synthetic_criteo.py", data, "--train", "16384", "--test", "4096",
       "--table_sizes", "5,300,7000,40,900,60,15,2000"])
  args = ["examples/dlrm/main.py", "--dataset_path", data, "--batch_size", "256",
          "--embedding_dim", "16", "--bottom_mlp_dims", "32,16", "--top_mlp_dims", "64,32,1",
          "--learning_rate", "2.0", "--warmup_steps", "20", "--decay_start_step", "100000",
          "--epochs", "3", "--save_path", str(tmp_path / "w")]
  if world > 1:
    args = ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
            "--master-addr", "127.0.0.1", "--master-port", str(29000 + os.getpid() % 900)] + args
  out = run(args, timeout=900)
  auc = float(out.split("AUC:")[1].split(",")[0])
  assert auc > 0.7, out[-500:]


def test_criteo_integer_lookup_example():
  out = run(["examples/criteo/main.py", "--batch_size", "64", "--steps", "2", "--vocab", "100"])
  assert "vocab sizes" in out


def test_criteo_integer_lookup_example_on_a_text_file(tmp_path):
  """Criteo text format (tab separated label, 13 counts, 26 hex ids, empty = missing) through
  IntegerLookup: the vocabulary grows to the number of distinct keys seen (+ the missing key)."""
  import random
  rng = random.Random(0)
  keys = [rng.getrandbits(32) for _ in range(9)]
  path = str(tmp_path / "train.txt")
  with open(path, "w", encoding="ascii") as f:
    for _ in range(200):
      nums = [str(rng.randint(0, 500)) if rng.random() > 0.2 else "" for _ in range(13)]
      cats = ["%08x" % rng.choice(keys) if rng.random() > 0.1 else "" for _ in range(26)]
      f.write("\t".join([str(rng.randint(0, 1))] + nums + cats) + "\n")
  out = run(["examples/criteo/main.py", "--data", path, "--batch_size", "200", "--vocab", "100",
             "--epochs", "1"])
  assert "vocab sizes (first 3 features) [10, 10, 10]" in out, out


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

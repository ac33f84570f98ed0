"""Multi-rank test cases, launched by ``tests/dist_utils.launch`` (one process per rank).

Mirrors the strategy matrix of the reference's ``dist_model_parallel_test.py`` (run_and_test
:244-291 and the test_* cases :319-640): the oracle is an undistributed model holding the full
tables whose gradients are averaged over ranks; the distributed model is loaded with the same
weights through ``set_weights`` and must match after one SGD step.
"""
import random

import numpy as np
import torch
import torch.distributed as dist
from torch import nn

import distributed_embeddings_b200 as de
from distributed_embeddings_b200 import dist_model_parallel as dmp
from distributed_embeddings_b200.ops.ragged import RaggedIds


class CustomEmbedding(nn.Module):
  """User-defined layer: only get_config / from_config + a 2-D parameter are required."""

  def __init__(self, input_dim, output_dim):
    super().__init__()
    self.input_dim, self.output_dim = input_dim, output_dim
    self.params = nn.Parameter(torch.rand(input_dim, output_dim))

  def forward(self, ids):
    return nn.functional.embedding(ids.to(torch.int64), self.params)

  def get_config(self):
    return {"input_dim": self.input_dim, "output_dim": self.output_dim}

  @classmethod
  def from_config(cls, cfg):
    return cls(cfg["input_dim"], cfg["output_dim"])


class EmbeddingListModel(nn.Module):

  def __init__(self, table_sizes, distribute=False, strategy="basic", dp_input=True,
               input_table_map=None, column_slice_threshold=None, test_custom_layer=False,
               combiner=None, row_slice_threshold=None, data_parallel_threshold=None,
               gpu_embedding_size=None, device="cpu", backend="auto", compute_dtype=None):
    super().__init__()
    embs = []
    for rows, width in table_sizes:
      if test_custom_layer:
        embs.append(CustomEmbedding(rows, width))
      else:
        embs.append(de.Embedding(rows, width, combiner=combiner, device=device))
    self.input_table_map = input_table_map
    if distribute:
      self.dist_embeddings = de.DistributedEmbedding(
          embs, strategy=strategy, dp_input=dp_input, input_table_map=input_table_map,
          column_slice_threshold=column_slice_threshold, row_slice_threshold=row_slice_threshold,
          data_parallel_threshold=data_parallel_threshold, gpu_embedding_size=gpu_embedding_size,
          device=device, backend=backend, compute_dtype=compute_dtype)
      self.embeddings = None
    else:
      self.dist_embeddings = None
      self.embeddings = nn.ModuleList(embs).to(device)
    imap = input_table_map or list(range(len(table_sizes)))
    total = sum(table_sizes[t][1] for t in imap)
    self.dense = nn.Linear(total, 5).to(device)

  def forward(self, inputs):
    if self.dist_embeddings is not None:
      outs = self.dist_embeddings(inputs)
    else:
      imap = self.input_table_map or list(range(len(self.embeddings)))
      outs = [self.embeddings[t](i) for i, t in zip(inputs, imap)]
    outs = [o.float() for o in outs]
    return self.dense(torch.cat(outs, dim=1))


def gen_table_sizes(seed, num_tables=None, world=1):
  rng = random.Random(seed)
  if num_tables is None:
    num_tables = rng.randint(1, 2 * world)
  return [[rng.randint(4, 20), rng.randint(4, 15)] for _ in range(num_tables)]


def gen_input_to_table_map(seed, num_tables):
  rng = random.Random(seed + 1)
  mapping = list(range(num_tables))
  for _ in range(3):
    mapping.append(rng.randint(0, num_tables - 1))
  rng.shuffle(mapping)
  return mapping


def gen_inputs(seed, global_batch, table_sizes, input_to_table_map=None, hotness=None,
               ragged=False, device="cpu", dtype=torch.int64):
  """Global inputs, identical on every rank (seeded)."""
  g = torch.Generator().manual_seed(seed + 7)
  imap = input_to_table_map or list(range(len(table_sizes)))
  if hotness is not None and isinstance(hotness, int):
    hotness = [hotness] * len(imap)
  outs = []
  for j, t in enumerate(imap):
    rows = table_sizes[t][0]
    if hotness is None:
      outs.append(torch.randint(0, rows, (global_batch,), generator=g, dtype=dtype).to(device))
    elif ragged:
      lens = torch.randint(1, hotness[j] + 1, (global_batch,), generator=g)
      vals = torch.randint(0, rows, (int(lens.sum()),), generator=g, dtype=dtype)
      outs.append(RaggedIds.from_row_lengths(vals, lens).to(device))
    else:
      outs.append(
          torch.randint(0, rows, (global_batch, hotness[j]), generator=g, dtype=dtype).to(device))
  return outs


def slice_batch(x, rank, local_batch):
  if isinstance(x, RaggedIds):
    return x.slice_rows(rank * local_batch, (rank + 1) * local_batch)
  return x[rank * local_batch:(rank + 1) * local_batch]


def run_and_test(ref_model, ref_inputs, test_model, test_inputs, fwd_tol=None, bwd_tol=1e-5,
                 lr=1.5):
  world = dist.get_world_size() if dist.is_initialized() else 1
  # same dense + embedding weights everywhere
  for p in ref_model.parameters():
    if world > 1:
      dist.broadcast(p.data, src=0)
  ref_tables = [e.embeddings if hasattr(e, "embeddings") else e.params
                for e in ref_model.embeddings]
  test_model.dist_embeddings.set_weights([t.detach().cpu().numpy() for t in ref_tables])
  with torch.no_grad():
    test_model.dense.weight.copy_(ref_model.dense.weight)
    test_model.dense.bias.copy_(ref_model.dense.bias)

  ref_params = list(ref_model.parameters())
  ref_out = torch.cumsum(ref_model(ref_inputs), dim=1)
  ref_grads = torch.autograd.grad(ref_out.sum(), ref_params)
  ref_grads = [g.to_dense() if g.is_sparse else g for g in ref_grads]
  if world > 1:
    for g in ref_grads:
      dist.all_reduce(g)
      g /= world

  test_params = list(test_model.parameters())
  test_out = torch.cumsum(test_model(test_inputs), dim=1)
  tape = dmp.DistributedGradientTape()
  test_grads = tape.gradient(test_out.sum(), test_params)

  if fwd_tol is None:
    assert torch.equal(ref_out, test_out), (ref_out - test_out).abs().max()
  else:
    torch.testing.assert_close(ref_out, test_out, rtol=fwd_tol, atol=fwd_tol)

  with torch.no_grad():
    for p, g in zip(ref_params, ref_grads):
      p -= lr * g
    for p, g in zip(test_params, test_grads):
      if g is None:
        continue
      p -= lr * (g.to_dense() if g.is_sparse else g)

  test_weights = test_model.dist_embeddings.get_weights(all_ranks=True)
  for ref_w, test_w in zip(ref_tables, test_weights):
    torch.testing.assert_close(ref_w.detach().cpu(), torch.from_numpy(test_w), rtol=bwd_tol,
                               atol=bwd_tol)
  torch.testing.assert_close(ref_model.dense.weight, test_model.dense.weight, rtol=bwd_tol,
                             atol=bwd_tol)


def _generic_case(rank, world, device, backend, *, seed=0, table_sizes=None, num_tables=None,
                  strategy="basic", dp_input=True, shared=False, hotness=None, ragged=False,
                  combiner=None, global_batch=24, fwd_tol=None, bwd_tol=1e-5, id_dtype=torch.int64,
                  **kw):
  torch.manual_seed(seed)
  if table_sizes is None:
    table_sizes = gen_table_sizes(seed, num_tables, world)
  imap = gen_input_to_table_map(seed, len(table_sizes)) if shared else None
  if hotness is not None and combiner is None:
    combiner = "sum"
  ref = EmbeddingListModel(table_sizes, distribute=False, input_table_map=imap, combiner=combiner,
                           test_custom_layer=kw.get("test_custom_layer", False), device=device)
  test = EmbeddingListModel(table_sizes, distribute=True, strategy=strategy, dp_input=dp_input,
                            input_table_map=imap, combiner=combiner, device=device,
                            backend=backend, **kw)
  glob = gen_inputs(seed, global_batch, table_sizes, imap, hotness, ragged, device, id_dtype)
  local_batch = global_batch // world
  dp_inputs = [slice_batch(x, rank, local_batch) for x in glob]
  if dp_input:
    test_inputs = dp_inputs
  else:
    ids = test.dist_embeddings.strategy.input_ids_list[rank]
    test_inputs = [glob[i] for i in ids]
  run_and_test(ref, dp_inputs, test, test_inputs, fwd_tol, bwd_tol)
  return test


def case_basic(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=11, **kw)


def case_memory_balanced(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=12, strategy="memory_balanced",
                num_tables=2 * world + 1, **kw)


def case_memory_optimized(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=13, strategy="memory_optimized",
                num_tables=2 * world + 1, **kw)


def case_row_slice(rank, world, device, backend, **kw):
  sizes = gen_table_sizes(14, 2 * world, world)
  _generic_case(rank, world, device, backend, seed=14, table_sizes=sizes,
                row_slice_threshold=8 * 13, **kw)


def case_data_parallel(rank, world, device, backend, **kw):
  sizes = gen_table_sizes(15, 2 * world + 1, world)
  _generic_case(rank, world, device, backend, seed=15, table_sizes=sizes,
                data_parallel_threshold=8 * 9, **kw)


def case_shared_dp(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=16, num_tables=world + 2, shared=True, **kw)


def case_shared_mp(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=17, num_tables=world + 2, shared=True,
                dp_input=False, **kw)


def case_mp_input(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=18, num_tables=2 * world, dp_input=False, **kw)


def case_column_slice_merge(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=19,
                table_sizes=[[100, 8], [5, 8], [10, 8], [25, 4]], strategy="memory_balanced",
                column_slice_threshold=45, **kw)


def case_column_slice_threshold(rank, world, device, backend, **kw):
  sizes = gen_table_sizes(20, world + 1, world)
  _generic_case(rank, world, device, backend, seed=20, table_sizes=sizes,
                column_slice_threshold=30, **kw)


def case_column_slice_dup_worker(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=21,
                table_sizes=[[10, 4], [11, 2], [4, 2], [4, 2]], strategy="memory_balanced",
                column_slice_threshold=10, **kw)


def case_fewer_tables_than_workers(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=22, table_sizes=[[16, 12]], **kw)


def case_custom_layer(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=23, num_tables=2 * world,
                test_custom_layer=True, **kw)


def case_all_modes(rank, world, device, backend, **kw):
  sizes = [[2, 8], [2, 16], [10, 8], [10, 16], [10, 16], [10, 32], [10, 128], [100, 16], [100, 32],
           [100, 128], [1000, 16], [1000, 48], [1000, 128], [10000, 64], [5000, 8], [50000, 8]]
  _generic_case(rank, world, device, backend, seed=24, table_sizes=sizes,
                strategy="memory_balanced", data_parallel_threshold=1000,
                column_slice_threshold=100000, row_slice_threshold=400000, **kw)


def case_multihot_dp(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=25, num_tables=2 * world, hotness=5,
                fwd_tol=1e-6, **kw)


def case_multihot_mp(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=26, num_tables=2 * world, hotness=5,
                dp_input=False, fwd_tol=1e-6, **kw)


def case_multihot_mean(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=27, num_tables=2 * world, hotness=4,
                combiner="mean", fwd_tol=1e-6, **kw)


def case_ragged_dp(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=28, num_tables=2 * world, hotness=6,
                ragged=True, fwd_tol=1e-6, **kw)


def case_ragged_mean(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=31, num_tables=2 * world, hotness=5,
                ragged=True, combiner="mean", fwd_tol=1e-6, **kw)


def case_ragged_mp(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=32, num_tables=2 * world, hotness=4,
                ragged=True, dp_input=False, fwd_tol=1e-6, **kw)


def case_cpu_offload(rank, world, device, backend, **kw):
  sizes = 4 * [[100, 32]] + 4 * [[1000, 64]]
  test = _generic_case(rank, world, device, backend, seed=29, table_sizes=sizes, hotness=3,
                       gpu_embedding_size=32000, fwd_tol=1e-6, **kw)
  assert any(l.cpu_offloaded for l in test.dist_embeddings.local_embedding_layers)


def case_int32_ids(rank, world, device, backend, **kw):
  _generic_case(rank, world, device, backend, seed=30, num_tables=2 * world,
                id_dtype=torch.int32, **kw)


def case_dp_to_mp_input(rank, world, device, backend, **kw):
  """Index exchange alone: every rank must receive the original global feature."""
  n_feat = 2 * world + 1
  rng = random.Random(5)
  owners = [rng.randrange(world) for _ in range(n_feat)]
  owners[0] = 0
  rank_to_features = {r: [k for k in range(n_feat) if owners[k] == r] for r in range(world)}
  if kw.get("unbalanced"):
    rank_to_features = {r: (list(range(n_feat)) if r == 0 else []) for r in range(world)}
  gb = 4 * world
  sizes = [[50, 4]] * n_feat
  glob = gen_inputs(3, gb, sizes, hotness=3, ragged=False, device=device)
  glob_r = gen_inputs(4, gb, sizes, hotness=4, ragged=True, device=device)
  mixed = [glob[k] if k % 2 == 0 else glob_r[k] for k in range(n_feat)]
  local = [slice_batch(x, rank, gb // world) for x in mixed]
  got = dmp.dp_to_mp_input(local, rank_to_features, rank, world)
  assert list(got.keys()) == rank_to_features[rank]
  for k, v in got.items():
    if isinstance(mixed[k], RaggedIds):
      assert v.to_lists() == mixed[k].to_lists()
    else:
      assert torch.equal(v.to(mixed[k].dtype), mixed[k])


def case_broadcast(rank, world, device, backend, **kw):
  torch.manual_seed(100 + rank)
  sizes = [[11, 7], [5, 8], [3, 8], [5, 8], [12, 25], [3, 12], [7, 13]]
  model = EmbeddingListModel(sizes, distribute=True, strategy="basic", device=device,
                             backend=backend)
  emb_w = [np.random.RandomState(1).rand(r, w).astype(np.float32) for r, w in sizes]
  model.dist_embeddings.set_weights(emb_w)
  ids = [torch.randint(0, s[0], (3,), generator=torch.Generator().manual_seed(2)).to(device)
         for s in sizes]
  out = model(ids)
  outs = [torch.empty_like(out) for _ in range(world)]
  dist.all_gather(outs, out.detach())
  if world > 1:
    assert not torch.allclose(outs[0], outs[1])
  dmp.broadcast_variables(model, root_rank=0)
  out = model(ids)
  dist.all_gather(outs, out.detach())
  for o in outs[1:]:
    assert torch.equal(outs[0], o)


def case_errors(rank, world, device, backend, **kw):
  import pytest
  sizes = [[10, 4], [10, 4], [10, 4]]
  model = EmbeddingListModel(sizes, distribute=True, device=device, backend=backend)
  with pytest.raises(ValueError):
    model.dist_embeddings.set_weights([np.zeros((10, 4), np.float32)])
  with pytest.raises(ValueError):
    model.dist_embeddings([torch.zeros(4, dtype=torch.int64, device=device)])
  mp_model = EmbeddingListModel(sizes, distribute=True, dp_input=False, device=device,
                                backend=backend)
  n_local = len(mp_model.dist_embeddings.strategy.local_maps[rank])
  if world > 1:
    with pytest.raises(ValueError, match="not divisible"):
      mp_model.dist_embeddings(
          [torch.zeros(world + 1, dtype=torch.int64, device=device)] * n_local)
  with pytest.raises(ValueError):
    de.DistributedEmbedding([de.Embedding(4, 4)], strategy="bogus", device=device)


def case_hybrid_optimizer(rank, world, device, backend, **kw):
  """DistributedOptimizer path (the reference's model.fit flow): loss + weights must match a
  manual all-reduce step."""
  torch.manual_seed(40)
  sizes = gen_table_sizes(40, 2 * world, world)
  ref = EmbeddingListModel(sizes, distribute=False, device=device)
  test = EmbeddingListModel(sizes, distribute=True, strategy="memory_balanced", device=device,
                            backend=backend)
  for p in ref.parameters():
    dist.broadcast(p.data, src=0)
  tables = [e.embeddings for e in ref.embeddings]
  test.dist_embeddings.set_weights([t.detach().cpu().numpy() for t in tables])
  torch.manual_seed(41 + rank)
  with torch.no_grad():
    test.dense.weight.normal_()  # deliberately different per rank before the broadcast
  dmp.BroadcastGlobalVariablesCallback(0)(test)
  with torch.no_grad():
    ref.dense.weight.copy_(test.dense.weight)
    ref.dense.bias.copy_(test.dense.bias)
  glob = gen_inputs(40, 8 * world, sizes, device=device)
  inputs = [slice_batch(x, rank, 8) for x in glob]
  opt = dmp.DistributedOptimizer(torch.optim.SGD(test.parameters(), lr=0.5))
  opt.zero_grad()
  loss = test(inputs).square().mean()
  loss.backward()
  opt.step()
  ref_loss = ref(inputs).square().mean()
  grads = torch.autograd.grad(ref_loss, list(ref.parameters()))
  with torch.no_grad():
    for p, g in zip(ref.parameters(), grads):
      g = g.to_dense() if g.is_sparse else g
      dist.all_reduce(g)
      p -= 0.5 * g / world
  torch.testing.assert_close(loss, ref_loss)
  got = test.dist_embeddings.get_weights(all_ranks=True)
  for t, w in zip(tables, got):
    torch.testing.assert_close(t.detach().cpu(), torch.from_numpy(w), rtol=1e-5, atol=1e-5)
  torch.testing.assert_close(ref.dense.weight, test.dense.weight, rtol=1e-5, atol=1e-5)


def case_dlrm_fast_step(rank, world, device, backend, optimizer="sgd", steps=2, **kw):
  """DLRMTrainStep on `world` ranks (local batches) == DLRMTrainStep in one process on the
  global batch: same loss, same dense weights, same embedding tables (global-mean contract)."""
  from distributed_embeddings_b200.models.dlrm import DLRM
  from distributed_embeddings_b200.models.dlrm_fast import DLRMTrainStep
  sizes = [200 + 13 * i for i in range(26)]
  torch.manual_seed(7)
  ref = DLRM(sizes, device=device, compute_dtype=torch.bfloat16, backend="fused", world_size=1,
             rank=0)
  torch.manual_seed(7)
  # dp_threshold: replicate the small tables (their update then rides on the dense SGD kernel)
  test = DLRM(sizes, device=device, compute_dtype=torch.bfloat16, backend="fused",
              data_parallel_threshold=kw.get("dp_threshold"))
  if kw.get("dp_threshold") and world > 1:
    assert len(test.embedding.dp_layers) > 0
  test.load_state_dict({k: v for k, v in ref.state_dict().items() if "embedding" not in k},
                       strict=False)
  test.embedding.set_weights(ref.embedding.get_weights(all_ranks=True))
  gb = 256 * world
  g = torch.Generator().manual_seed(3)
  lr = 0.5 if optimizer == "sgd" else 0.05
  t_ref = DLRMTrainStep(ref, lr=lr, embedding_optimizer=optimizer, use_cuda_graph=False)
  t_test = DLRMTrainStep(test, lr=lr, embedding_optimizer=optimizer,
                         use_cuda_graph=kw.get("graph", True))
  lb = gb // world
  for _ in range(steps):
    num = torch.rand(gb, 13, generator=g).to(device)
    cat = torch.stack([torch.randint(0, s, (gb,), generator=g, dtype=torch.int32)
                       for s in sizes]).to(device)
    lab = torch.randint(0, 2, (gb,), generator=g).float().to(device)
    l_ref = t_ref.step(num, cat, lab).clone()
    sl = slice(rank * lb, (rank + 1) * lb)
    l_test = t_test.step(num[sl], cat[:, sl].contiguous(), lab[sl]).clone()
    dist.all_reduce(l_test)
    torch.testing.assert_close(l_test / world, l_ref, rtol=1e-2, atol=1e-3)
  t_test.ctx.check_errors()
  for (n1, p1), (n2, p2) in zip(ref.named_parameters(), test.named_parameters()):
    if "embedding" in n1:
      continue
    torch.testing.assert_close(p2, p1, rtol=3e-2, atol=3e-3, msg=lambda m: f"{n1}: {m}")
  w_ref = ref.embedding.get_weights(all_ranks=True)
  w_test = test.embedding.get_weights(all_ranks=True)
  for a, b in zip(w_ref, w_test):
    torch.testing.assert_close(torch.from_numpy(b), torch.from_numpy(a), rtol=3e-2, atol=3e-3)


def case_synthetic_fast_step(rank, world, device, backend, optimizer="adagrad", model="tiny",
                             dp_input=False, steps=2, graph=True, **kw):
  """SyntheticTrainStep (hand-scheduled kernels, CUDA graph) on `world` ranks against a
  single-process plain-PyTorch oracle on the global batch: tables as dense tensors, index + sum
  pooling, concat (+ average pooling), nn.functional MLP under bf16 autocast, autograd, dense
  SGD / Adagrad.  Shares nothing with the framework but the initial weights."""
  from distributed_embeddings_b200.models.configs import (ModelConfig, expand, scaled,
                                                          synthetic_models_v3)
  from distributed_embeddings_b200.models.synthetic import SyntheticModel
  from distributed_embeddings_b200.models.synthetic_fast import SyntheticTrainStep
  cfg = scaled(synthetic_models_v3[model], kw.get("row_scale", 2e-4))
  if kw.get("interact_stride"):
    cfg = ModelConfig(cfg.name, cfg.embedding_configs, cfg.mlp_sizes, cfg.num_numerical_features,
                      kw["interact_stride"])
  tables, imap, hots = expand(cfg)
  torch.manual_seed(21)
  net = SyntheticModel(cfg, dp_input=dp_input, device=device, compute_dtype=torch.bfloat16,
                       backend="fused", column_slice_threshold=kw.get("column_slice_threshold"))
  de.broadcast_variables(net)
  lr = 0.05 if optimizer == "sgd" else 0.01
  w0 = net.embedding.get_weights(all_ranks=True)
  lins = [m for m in net.mlp if isinstance(m, torch.nn.Linear)]
  dense0 = [(l.weight.detach().float().clone(), l.bias.detach().float().clone()) for l in lins]
  step = SyntheticTrainStep(net, lr=lr, embedding_optimizer=optimizer, use_cuda_graph=graph)
  gb = 64 * world
  lb = gb // world
  g = torch.Generator().manual_seed(4)
  batches = []
  for _ in range(steps):
    num = (torch.rand(gb, cfg.num_numerical_features, generator=g) * 2).to(device)
    ids = [torch.randint(0, tables[t][0], (gb, h), generator=g).to(device)
           for t, h in zip(imap, hots)]
    lab = torch.randint(0, 2, (gb, 1), generator=g).float().to(device)
    batches.append((num, ids, lab))
  mine = net.embedding.strategy.input_ids_list[rank]
  losses = []
  sl = slice(rank * lb, (rank + 1) * lb)
  for num, ids, lab in batches:
    cat = [x[sl] for x in ids] if dp_input else [ids[i] for i in mine]
    loss = step.step(num[sl], cat, lab[sl]).clone()
    if world > 1:
      dist.all_reduce(loss)
    losses.append(float(loss) / world)
  step.ctx.check_errors()
  w1 = net.embedding.get_weights(all_ranks=True)

  # ---- oracle
  tabs = [torch.from_numpy(w).to(device).requires_grad_(True) for w in w0]
  dense = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in dense0]
  acc = [torch.full_like(t, 0.1) for t in tabs]  # Adagrad initial accumulator value
  in_dim, in_pad = net._in_dim, net._in_pad
  olosses = []
  for num, ids, lab in batches:
    with torch.autocast("cuda", dtype=torch.bfloat16):
      outs = [tabs[t][x.long()].sum(1) for t, x in zip(imap, ids)]
      x = torch.cat(outs, dim=1).to(torch.bfloat16)
      if cfg.interact_stride:
        from distributed_embeddings_b200.models.synthetic import _interact
        x = _interact(x.float(), cfg.interact_stride).to(torch.bfloat16)
      h = torch.cat([x, num.to(torch.bfloat16),
                     torch.zeros(gb, in_pad, dtype=torch.bfloat16, device=device)], dim=1)
      for i, (w, b) in enumerate(dense):
        h = torch.nn.functional.linear(h, w, b)
        if i < len(dense) - 1:
          h = torch.relu(h)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(h.float(), lab)
    olosses.append(float(loss))
    params = tabs + [t for wb in dense for t in wb]
    grads = torch.autograd.grad(loss, params)
    with torch.no_grad():
      for i, (p, gr) in enumerate(zip(params, grads)):
        if i < len(tabs) and optimizer == "adagrad":
          acc[i] += gr * gr
          p -= lr * gr / (acc[i].sqrt() + 1e-7)
        else:
          p -= lr * gr
  for a, b in zip(losses, olosses):
    assert abs(a - b) < 2e-2, (losses, olosses)
  sq_e = sum(float(((torch.from_numpy(w1[t]).to(device) - tabs[t].detach()).double()**2).sum())
             for t in range(len(tabs)))
  sq_u = sum(float(((tabs[t].detach() - torch.from_numpy(w0[t]).to(device)).double()**2).sum())
             for t in range(len(tabs)))
  rel = (sq_e / max(sq_u, 1e-30))**0.5
  assert sq_u > 0 and rel < 0.1, f"table update differs from the oracle: rel L2 {rel:.4f}"
  for l, (w, b) in zip(lins, dense):
    torch.testing.assert_close(l.weight.detach().float(), w.detach(), rtol=5e-2, atol=5e-3)


def case_checkpoint_resharding(rank, world, device, backend, **kw):
  """Weights written under one sharding load under another and give identical outputs; also
  through .npy files (memory mapped) and with use_lock."""
  import os
  import tempfile
  torch.manual_seed(50)
  sizes = [[300, 8], [40, 16], [1000, 8], [64, 4], [500, 16], [9, 8]]
  a = EmbeddingListModel(sizes, distribute=True, strategy="memory_balanced", combiner="sum",
                         column_slice_threshold=1500, device=device, backend=backend)
  b = EmbeddingListModel(sizes, distribute=True, strategy="basic", combiner="sum",
                         row_slice_threshold=7000, data_parallel_threshold=100, device=device,
                         backend=backend)
  ref = [np.random.RandomState(3 + i).rand(r, w).astype(np.float32) for i, (r, w) in
         enumerate(sizes)]
  a.dist_embeddings.set_weights(ref)
  got = a.dist_embeddings.get_weights(all_ranks=True)
  for x, y in zip(ref, got):
    np.testing.assert_array_equal(x, y)
  # rank-0-only gather, files, lock-step loading
  only0 = a.dist_embeddings.get_weights()
  assert (len(only0) == len(sizes)) == (rank == 0)
  tmp = tempfile.mkdtemp(prefix=f"ckpt_r{rank}_")
  paths = []
  for i, w in enumerate(got):
    path = os.path.join(tmp, f"t{i}.npy")
    np.save(path, w)
    paths.append(path)
  b.dist_embeddings.set_weights(paths, chunk=1024, use_lock=True)
  for x, y in zip(ref, b.dist_embeddings.get_weights(all_ranks=True)):
    np.testing.assert_array_equal(x, y)
  glob = gen_inputs(51, 8 * world, sizes, hotness=3, device=device)
  inputs = [slice_batch(x, rank, 8) for x in glob]
  oa = torch.cat(a.dist_embeddings(inputs), 1)
  ob = torch.cat(b.dist_embeddings(inputs), 1)
  torch.testing.assert_close(oa, ob, rtol=1e-6, atol=1e-6)


def case_file_checkpoint(rank, world, device, backend, **kw):
  """save_weights: concurrent writers (one process per rank) into shared global-layout files,
  no gather; load_weights under another plan restores the tables bit for bit."""
  import os
  import shutil
  import tempfile
  sizes = [[300, 8], [40, 16], [1000, 8], [64, 4], [500, 16], [9, 8]]
  a = EmbeddingListModel(sizes, distribute=True, strategy="memory_balanced", combiner="sum",
                         column_slice_threshold=1500, device=device, backend=backend)
  b = EmbeddingListModel(sizes, distribute=True, strategy="basic", combiner="sum",
                         row_slice_threshold=7000, data_parallel_threshold=100, device=device,
                         backend=backend)
  ref = [np.random.RandomState(7 + i).rand(r, w).astype(np.float32) for i, (r, w) in
         enumerate(sizes)]
  a.dist_embeddings.set_weights(ref)
  box = [tempfile.mkdtemp(prefix="de_ckpt_") if rank == 0 else None]
  if world > 1:
    dist.broadcast_object_list(box, src=0)
  ckpt = os.path.join(box[0], "step_1")
  paths = a.dist_embeddings.save_weights(ckpt, chunk=512)
  assert len(paths) == len(sizes)
  for x, path in zip(ref, paths):  # every rank sees complete files after the call returns
    np.testing.assert_array_equal(np.load(path), x)
  b.dist_embeddings.load_weights(ckpt, chunk=1024, use_lock=True)
  for x, y in zip(ref, b.dist_embeddings.get_weights(all_ranks=True)):
    np.testing.assert_array_equal(x, y)
  # a second save from the row-sliced / replicated plan overwrites the same files consistently
  b.dist_embeddings.save_weights(ckpt)
  for x, path in zip(ref, paths):
    np.testing.assert_array_equal(np.load(path), x)
  if world > 1:
    dist.barrier()
  if rank == 0:
    shutil.rmtree(box[0], ignore_errors=True)


def case_ddp_interop(rank, world, device, backend, **kw):
  """torch DistributedDataParallel around a model with a DistributedEmbedding: model-parallel
  tables are excluded from DDP (exclude_model_parallel_from_ddp), replicated tables and the dense
  layer are averaged by DDP; one SGD step equals the undistributed model on the global batch."""
  from torch.nn.parallel import DistributedDataParallel as DDP
  torch.manual_seed(60)
  sizes = [[30, 8], [12, 4], [50, 8], [9, 4], [40, 8]]
  ref = EmbeddingListModel(sizes, distribute=False, combiner="sum", device=device)
  test = EmbeddingListModel(sizes, distribute=True, strategy="memory_balanced", combiner="sum",
                            data_parallel_threshold=60, device=device, backend=backend)
  for p in ref.parameters():
    dist.broadcast(p.data, src=0)
  test.dist_embeddings.set_weights([e.embeddings.detach().cpu().numpy() for e in ref.embeddings])
  with torch.no_grad():
    test.dense.weight.copy_(ref.dense.weight)
    test.dense.bias.copy_(ref.dense.bias)
  ignored = dmp.exclude_model_parallel_from_ddp(test)
  n_mp = len(test.dist_embeddings.mp_parameters())
  assert len(ignored) == n_mp > 0 and len(test.dist_embeddings.dp_layers) > 0
  dev = torch.device(device)
  ddp = DDP(test) if dev.type == "cpu" else DDP(test, device_ids=[dev.index])
  lb, lr = 6, 0.5
  glob = gen_inputs(61, lb * world, sizes, hotness=2, device=device)
  inputs = [slice_batch(x, rank, lb) for x in glob]
  loss = ddp(inputs).square().mean()
  loss.backward()
  with torch.no_grad():
    for p in test.parameters():
      if p.grad is None:
        continue
      g = p.grad.to_dense() if p.grad.is_sparse else p.grad
      p -= lr * g  # model-parallel gradients already follow the global-mean contract
  ref_loss = ref(glob).square().mean()
  grads = torch.autograd.grad(ref_loss, list(ref.parameters()))
  with torch.no_grad():
    for p, g in zip(ref.parameters(), grads):
      p -= lr * (g.to_dense() if g.is_sparse else g)
  got = test.dist_embeddings.get_weights(all_ranks=True)
  for e, w in zip(ref.embeddings, got):
    torch.testing.assert_close(torch.from_numpy(w), e.embeddings.detach().cpu(), rtol=1e-5,
                               atol=1e-6)
  torch.testing.assert_close(test.dense.weight, ref.dense.weight, rtol=1e-5, atol=1e-6)
  torch.testing.assert_close(test.dense.bias, ref.dense.bias, rtol=1e-5, atol=1e-6)


def case_batch_mismatch(rank, world, device, backend, **kw):
  import pytest
  sizes = [[10, 4], [10, 4]]
  model = EmbeddingListModel(sizes, distribute=True, device=device, backend=backend)
  bs = 4 + (2 if rank == 1 else 0)
  ids = [torch.zeros(bs, dtype=torch.int64, device=device) for _ in sizes]
  with pytest.raises(ValueError, match="same batchsize"):
    model.dist_embeddings(ids)


def case_fuzz(rank, world, device, backend, n_seeds=6, seed0=100, **kw):
  """Randomised plans (same on every rank): strategy x thresholds x shared tables x hotness x
  input mode, each checked forward + backward + checkpoint against the unsharded model."""
  rejected = 0
  for seed in range(seed0, seed0 + n_seeds):
    rng = random.Random(seed * 7919)
    strategy = rng.choice(["basic", "memory_balanced", "memory_optimized", "traffic_balanced"])
    num_tables = rng.randint(world, 3 * world)
    table_sizes = [[rng.randint(4, 40), rng.choice([4, 6, 8, 12, 16])] for _ in range(num_tables)]
    opts = {}
    if rng.random() < 0.5:
      opts["column_slice_threshold"] = rng.choice([40, 100, 200])
    dp_input = rng.random() < 0.7
    if dp_input and rng.random() < 0.4:
      opts["data_parallel_threshold"] = rng.choice([30, 60])
    if dp_input and rng.random() < 0.4:
      opts["row_slice_threshold"] = rng.choice([300, 500])
    hot = rng.choice([None, None, 1, 3])
    ragged = hot is not None and hot > 1 and rng.random() < 0.3 and \
        "row_slice_threshold" not in opts
    if rng.random() < 0.25:
      opts["gpu_embedding_size"] = rng.choice([100, 300, 800])  # some tables go to host memory
    id_dtype = rng.choice([torch.int64, torch.int32])
    if kw.get("backend_custom_ok", True) and hot is None and rng.random() < 0.15 and \
        backend != "fused":
      opts["test_custom_layer"] = True
    try:
      model = _generic_case(rank, world, device, backend, seed=seed, table_sizes=table_sizes,
                            strategy=strategy, dp_input=dp_input, shared=rng.random() < 0.5,
                            hotness=hot, ragged=ragged,
                            combiner=rng.choice(["sum", "mean"]) if hot else None,
                            global_batch=4 * world, fwd_tol=1e-5, bwd_tol=1e-4, id_dtype=id_dtype,
                            **opts, **kw)
      # many plans in one process: hand the peer-mapped buffers back in lock step
      model.dist_embeddings.close()
    except ValueError as e:
      # infeasible plans must be rejected identically on every rank
      if "Not enough table" not in str(e):
        raise
      rejected += 1
    except Exception as e:  # pylint: disable=broad-except
      raise AssertionError(f"fuzz seed {seed}: strategy={strategy} tables={table_sizes} "
                           f"dp_input={dp_input} hot={hot} ragged={ragged} opts={opts}: {e}") from e
  assert rejected <= n_seeds // 2, f"{rejected} of {n_seeds} random plans were infeasible"
  if rank == 0:
    print(f"case_fuzz world={world}: {n_seeds - rejected} plans checked, {rejected} infeasible")


def case_subgroups(rank, world, device, backend, **kw):
  """Two independent model-parallel groups inside one job (process_group argument): every
  collective of the wrapper must stay inside its group and use group ranks."""
  assert world % 2 == 0
  half = world // 2
  groups = [dist.new_group(list(range(0, half))), dist.new_group(list(range(half, world)))]
  gid = rank // half
  group = groups[gid]
  grank = rank - gid * half
  rng = np.random.default_rng(1234 + gid)  # different tables / inputs per group
  sizes = [[30, 8], [17, 4], [64, 16], [9, 8]]
  tables = [rng.standard_normal((r, w)).astype(np.float32) for r, w in sizes]
  embs = [de.Embedding(r, w, combiner="sum", device=device) for r, w in sizes]
  emb = de.DistributedEmbedding(embs, strategy="memory_balanced", column_slice_threshold=100,
                                row_slice_threshold=900, process_group=group, device=device,
                                backend=backend)
  assert emb.world_size == half and emb.rank == grank
  emb.set_weights(tables)
  lb, hot = 3, 2
  glob = [torch.from_numpy(rng.integers(0, r, size=(lb * half, hot))).to(device) for r, _ in sizes]
  local = [g[grank * lb:(grank + 1) * lb] for g in glob]
  outs = emb(local)
  for t, o in enumerate(outs):
    ref = tables[t][local[t].cpu().numpy()].sum(1)
    np.testing.assert_allclose(o.detach().float().cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
  # one SGD step on sum(outputs): d table[row] = -lr * (#occurrences in the group's global batch)
  # / group size (mean-of-ranks contract for model-parallel gradients)
  lr = 0.5
  loss = sum(o.float().sum() for o in outs)
  params = [p for p in emb.parameters()]
  grads = torch.autograd.grad(loss, params, allow_unused=True)
  with torch.no_grad():
    for p, g in zip(params, grads):
      if g is not None:
        p -= lr * (g.to_dense() if g.is_sparse else g)
  got = emb.get_weights(all_ranks=True)
  for t, (r, w) in enumerate(sizes):
    cnt = np.bincount(glob[t].cpu().numpy().reshape(-1), minlength=r).astype(np.float32)
    exp = tables[t] - lr * cnt[:, None] / half
    np.testing.assert_allclose(got[t], exp, rtol=1e-5, atol=1e-5)

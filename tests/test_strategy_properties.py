"""Property-based invariants of the sharding planner (hypothesis): whatever the table list,
thresholds and world size, the plan must tile every table exactly once and route every input to
a place that can serve it.  Complements the golden vectors of test_strategy.py."""
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

from distributed_embeddings_b200.parallel.strategy import STRATEGIES, DistEmbeddingStrategy  # noqa: E402


@st.composite
def plans(draw):
  n_tables = draw(st.integers(1, 14))
  rows = draw(st.lists(st.integers(1, 5000), min_size=n_tables, max_size=n_tables))
  widths = draw(st.lists(st.sampled_from([4, 8, 16, 32, 64, 128]), min_size=n_tables,
                         max_size=n_tables))
  world = draw(st.sampled_from([1, 2, 3, 4, 8]))
  strategy = draw(st.sampled_from(STRATEGIES))
  n_extra = draw(st.integers(0, 4))
  imap = list(range(n_tables)) + [draw(st.integers(0, n_tables - 1)) for _ in range(n_extra)]
  col_thr = draw(st.one_of(st.none(), st.integers(16, 200000)))
  row_thr = draw(st.one_of(st.none(), st.integers(50000, 700000)))
  dp_thr = draw(st.one_of(st.none(), st.integers(1, 3000)))
  cfgs = [{"input_dim": r, "output_dim": w, "combiner": "sum"} for r, w in zip(rows, widths)]
  return cfgs, world, strategy, imap, col_thr, row_thr, dp_thr


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(plans())
def test_plan_invariants(p):
  cfgs, world, strategy, imap, col_thr, row_thr, dp_thr = p
  try:
    plan = DistEmbeddingStrategy(cfgs, world, strategy, input_table_map=imap,
                                 column_slice_threshold=col_thr, row_slice_threshold=row_thr,
                                 data_parallel_threshold=dp_thr)
  except ValueError:
    return  # an infeasible combination rejected up front is fine
  n_tables = len(cfgs)
  # 1. every table is in exactly one group
  groups = plan.table_groups
  assert sorted(groups[0] + groups[1] + groups[2]) == list(range(n_tables))
  # 2. inputs are regrouped by a permutation and rev_group_ids undoes it
  flat_inputs = plan.input_groups[0] + plan.input_groups[1] + plan.input_groups[2]
  assert sorted(flat_inputs) == list(range(len(imap)))
  assert [flat_inputs[i] for i in plan.rev_group_ids] == list(range(len(imap)))
  # 3. table-parallel shards tile every column-group table exactly once, widths add up
  cover = {gt: [] for gt in range(len(groups[1]))}
  for r, shards in enumerate(plan.shards):
    for s in shards:
      assert s.rank == r and 0 <= s.col_start < s.col_end
      cover[s.table].append((s.col_start, s.col_end, s.rows))
  for gt, t in enumerate(groups[1]):
    rows, width = cfgs[t]["input_dim"], cfgs[t]["output_dim"]
    pieces = sorted(cover[gt])
    assert pieces, f"table {t} not placed"
    assert pieces[0][0] == 0 and pieces[-1][1] == width
    for (a0, a1, ra), (b0, _, _) in zip(pieces, pieces[1:]):
      assert a1 == b0, "column ranges must be contiguous and disjoint"
    assert all(pr == rows for _, _, pr in pieces)
    n = len(pieces)
    # a power of two capped by the world size and the width; slices meeting on a rank merge
    assert n <= min(world, width)
  # 4. local configs: fused tables hold the shards at their row offsets
  for r in range(world):
    for s in plan.shards[r]:
      lc = plan.local_configs[r][s.local_table]
      assert lc["output_dim"] == s.width
      assert 0 <= s.row_offset and s.row_offset + s.rows <= lc["input_dim"]
    assert len(plan.local_maps[r]) == len(plan.input_ids_list[r]) == \
        len(plan.local_input_offsets[r])
    for m in plan.local_maps[r]:
      assert 0 <= m < len(plan.local_configs[r])
  # 5. every (column-group input, slice) is served by exactly one rank
  served = sorted(i for r in range(world) for i in plan.input_ids_list[r])
  expect = []
  for gi, gt in enumerate(plan.map_groups[1]):
    expect += [gi] * len(cover[gt])
  assert served == sorted(expect)
  # 6. row slices: contiguous ranges covering all rows, one per rank
  for gt, t in enumerate(groups[2]):
    rr = plan.row_ranges[gt]
    assert len(rr) == world and rr[0][0] == 0 and rr[-1][1] == cfgs[t]["input_dim"]
    for (a0, a1), (b0, _) in zip(rr, rr[1:]):
      assert a0 <= a1 == b0
  # 7. the fingerprint is a pure function of the inputs
  again = DistEmbeddingStrategy(cfgs, world, strategy, input_table_map=imap,
                                column_slice_threshold=col_thr, row_slice_threshold=row_thr,
                                data_parallel_threshold=dp_thr)
  assert again.fingerprint() == plan.fingerprint()


@st.composite
def small_models(draw):
  n_tables = draw(st.integers(1, 6))
  rows = draw(st.lists(st.integers(1, 60), min_size=n_tables, max_size=n_tables))
  widths = draw(st.lists(st.sampled_from([1, 3, 4, 8, 20]), min_size=n_tables, max_size=n_tables))
  combiners = draw(st.lists(st.sampled_from(["sum", "mean"]), min_size=n_tables,
                            max_size=n_tables))
  n_extra = draw(st.integers(0, 3))
  imap = list(range(n_tables)) + [draw(st.integers(0, n_tables - 1)) for _ in range(n_extra)]
  hots = [draw(st.integers(1, 4)) for _ in imap]
  strategy = draw(st.sampled_from(STRATEGIES))
  col_thr = draw(st.one_of(st.none(), st.integers(4, 400)))
  seed = draw(st.integers(0, 2**16))
  return rows, widths, combiners, imap, hots, strategy, col_thr, seed


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(small_models())
def test_single_worker_forward_and_checkpoint_roundtrip(m):
  """World size 1, generic back end on the CPU: whatever the plan does internally (fusion of
  tables, column slicing requests, shared tables), the outputs equal a plain gather + pool of
  the global tables and get_weights returns what set_weights was given."""
  import numpy as np
  import torch

  from distributed_embeddings_b200 import DistributedEmbedding
  rows, widths, combiners, imap, hots, strategy, col_thr, seed = m
  rng = np.random.default_rng(seed)
  cfgs = [{"input_dim": r, "output_dim": w, "combiner": c}
          for r, w, c in zip(rows, widths, combiners)]
  de = DistributedEmbedding(cfgs, strategy=strategy, column_slice_threshold=col_thr,
                            input_table_map=imap, device="cpu", backend="torch", world_size=1,
                            rank=0)
  tables = [rng.standard_normal((r, w)).astype(np.float32) for r, w in zip(rows, widths)]
  de.set_weights(tables)
  back = de.get_weights()
  for a, b_ in zip(tables, back):
    np.testing.assert_array_equal(a, b_)
  batch = 5
  ids = [torch.from_numpy(rng.integers(0, rows[t], size=(batch, h))) for t, h in zip(imap, hots)]
  outs = de(ids)
  for i, (t, out) in enumerate(zip(imap, outs)):
    g = tables[t][ids[i].numpy()]  # [b, h, w]
    ref = g.sum(1) if combiners[t] == "sum" else g.mean(1)
    np.testing.assert_allclose(out.detach().numpy(), ref, rtol=1e-5, atol=1e-5)

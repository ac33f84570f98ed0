"""Multi-rank behaviour of DistributedEmbedding on CPU (gloo, torch back end).  The same cases
run on GPUs with the fused back end in test_dist_gpu.py."""
import pytest

from dist_utils import launch

CASES = [
    "case_basic", "case_memory_balanced", "case_memory_optimized", "case_row_slice",
    "case_data_parallel", "case_shared_dp", "case_shared_mp", "case_mp_input",
    "case_column_slice_threshold", "case_fewer_tables_than_workers", "case_custom_layer",
    "case_multihot_dp", "case_multihot_mp", "case_multihot_mean", "case_ragged_dp",
    "case_cpu_offload", "case_int32_ids", "case_dp_to_mp_input", "case_broadcast", "case_errors",
    "case_hybrid_optimizer", "case_checkpoint_resharding", "case_batch_mismatch",
    "case_ragged_mean", "case_ragged_mp",
]


@pytest.mark.parametrize("case", CASES)
def test_world2(case):
  launch(case, world=2)


@pytest.mark.parametrize("case", ["case_column_slice_merge", "case_column_slice_dup_worker",
                                  "case_all_modes", "case_basic", "case_row_slice",
                                  "case_checkpoint_resharding"])
def test_world4(case):
  launch(case, world=4)


def test_world3_uneven():
  launch("case_memory_optimized", world=3, global_batch=24)


def test_dp_to_mp_unbalanced():
  launch("case_dp_to_mp_input", world=2, unbalanced=True)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_fuzz_plans(world):
  launch("case_fuzz", world=world, n_seeds=6, seed0=100 * world)


def test_independent_subgroups():
  launch("case_subgroups", world=4)


@pytest.mark.parametrize("world", [2, 3])
def test_file_checkpoint_parallel_writers(world):
  launch("case_file_checkpoint", world=world)


@pytest.mark.parametrize("world", [2, 3])
def test_ddp_interop(world):
  launch("case_ddp_interop", world=world)

"""Fused execution engine of :class:`DistributedEmbedding` (CUDA, sm_100a).

Data flow of one training step on every rank (``W`` ranks, local batch ``b``, global ``B = W*b``):

forward
  1. ids of all features are staged in a *symmetric* buffer ``in_buf`` (data loaders can write it
     directly, so the H2D copy is the staging);
  2. flag barrier; one pull kernel copies the ids of this rank's features out of every peer's
     ``in_buf`` over NVLink (index all-to-all without NCCL);
  3. one descriptor-driven lookup kernel gathers + pools the rows of *all* local tables and stores
     every pooled row straight into the requester's ``out_buf`` at its final column offset
     (pooled-vector all-to-all + reorder + column-slice concat fused into the gather epilogue);
  4. flag barrier; ``out_buf`` is the ``[b, sum(widths)]`` activation (``concat=True`` returns it
     without a copy).
backward
  1. the incoming gradient is written to the symmetric ``grad_buf``; flag barrier;
  2. the owner pulls gradient rows from the peers' ``grad_buf`` inside the update kernel: either
     vector ``red.global.add`` straight into the table (SGD), or the sorted / deduplicated path
     that sums each unique row once and applies SGD / Adagrad / row-wise Adagrad / Adam in place.
     No sparse gradient tensor, no host sync, no NCCL.

Replaces ``_call_table_parallel`` / ``_call_row_slice`` / ``_call_data_parallel`` plus Horovod's
alltoall and the TF sparse optimizer kernels (reference dist_model_parallel.py:836-904).
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ..ops import _native
from ..ops._native import INPUT_DESC, TABLE_DESC
from ..ops.ragged import RaggedIds
from ..utils import nvtx
from .comm import CommContext

_OPT_KIND = {"sgd": _native.OPT_SGD, "adagrad": _native.OPT_ADAGRAD,
             "rowwise_adagrad": _native.OPT_ROWWISE_ADAGRAD, "adam": _native.OPT_ADAM}
_COMB = {None: 0, "sum": 0, "mean": 1}


def _weight(layer):
  return layer.embeddings


def _dev_ptr(t: torch.Tensor) -> int:
  """Device-visible address of a tensor: CUDA tensors directly, pinned host tensors (CPU
  offloaded tables / optimizer state) through their zero-copy UVA mapping."""
  if t.is_cuda:
    return t.data_ptr()
  if not t.is_pinned():
    raise RuntimeError("host-resident tables must live in pinned memory for the fused back end")
  return int(_native.require().host_device_pointer(t))


class _FusedFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, engine, token, *weights):  # pylint: disable=arguments-differ
    ctx.engine = engine
    ctx.n_weights = len(weights)
    out = engine._run_forward()
    return out

  @staticmethod
  def backward(ctx, grad_out):  # pylint: disable=arguments-differ
    grads = ctx.engine._run_backward(grad_out)
    return (None, None) + tuple(grads)


class FusedEngine:

  def __init__(self, de, dry=None):
    """``dry``: a :class:`dry_run.DryRank` - the engine then builds its descriptors against host
    buffers and runs them through the Python plan interpreter instead of the CUDA kernels (all
    ranks of a plan in one process, no GPU; see ``dry_run.py``)."""
    self.de = de
    self.st = de.strategy
    self.W, self.rank = de.world_size, de.rank
    self.device = de.device
    self.dry = dry is not None
    if self.dry:
      self.ops, self.ctx = dry.ops, dry.ctx
      dry.attach(self)
    else:
      self.ops = _native.require()
      self.ctx = CommContext.for_group(de.group, self.device)
    if self.W > 1 and not self.ctx.p2p:
      raise RuntimeError("fused back end needs CUDA peer access between all ranks")
    st = self.st
    imap = st.input_table_map
    self.out_widths = [int(st.global_configs[t]["output_dim"]) for t in imap]
    self.out_cols = [0]
    for w in self.out_widths:
      self.out_cols.append(self.out_cols[-1] + w)
    self.total_width = self.out_cols[-1]
    self.compute_dtype = de.compute_dtype
    if self.compute_dtype not in (torch.float32, torch.bfloat16):
      raise ValueError("fused back end supports fp32 and bf16 activations")
    self._key = None
    self._token = torch.zeros((), device=self.device)
    self.lr_t = torch.zeros(1, dtype=torch.float32, device=self.device)
    self.opt_state: Dict[int, List[torch.Tensor]] = {}
    self._tables_dirty = True
    # local model-parallel tables: table-parallel first, then row slices
    self.mp_layers = list(de.local_embedding_layers) + list(de.row_layers)
    self.n_col_tables = len(de.local_embedding_layers)
    # host-resident tables are read zero-copy over PCIe; their update must not use atomics
    self.has_offload = any(getattr(l, "cpu_offloaded", False) for l in self.mp_layers)

  def _ptr(self, t: torch.Tensor) -> int:
    return t.data_ptr() if self.dry else _dev_ptr(t)

  # ------------------------------------------------------------------ capabilities
  def supports(self, inputs) -> bool:
    st, de = self.st, self.de
    col_inputs = set(st.input_groups[1]) if de.dp_input else set(range(len(inputs)))
    for i, x in enumerate(inputs):
      if isinstance(x, RaggedIds):
        if i not in col_inputs:  # ragged only for table-parallel features
          return False
        continue
      if not isinstance(x, torch.Tensor) or x.dim() > 2:
        return False
    return True

  # ------------------------------------------------------------------ plan -> descriptors
  def _build(self, b: int, hots: Tuple[int, ...], ids64: bool):
    de, st, W, rank = self.de, self.st, self.W, self.rank
    dev = self.device
    id_dtype = torch.int64 if ids64 else torch.int32
    idsz = 8 if ids64 else 4
    B = b * W if de.dp_input else b
    lb = b if de.dp_input else b // W  # local (requester) batch
    self.B, self.lb, self.hots, self.ids64, self.id_dtype = B, lb, hots, ids64, id_dtype
    csz = 2 if self.compute_dtype == torch.bfloat16 else 4

    col_group = st.input_groups[1] if de.dp_input else list(range(len(hots)))
    col_map = st.map_groups[1]
    # --- staging of data-parallel inputs
    # hots[i] > 0: fixed hotness; hots[i] < 0: ragged input with capacity -hots[i] ids per sample
    self.ragged = [h < 0 for h in hots]
    self.any_ragged = any(self.ragged)
    n_rag = sum(self.ragged)
    rag_index = {}
    for i, r in enumerate(self.ragged):
      if r:
        rag_index[i] = len(rag_index)
    if de.dp_input:
      in_off, pos = [], 0
      for h in hots:
        in_off.append(pos)
        pos += b * abs(h)
      self.in_elems = max(pos, 1)
      if W > 1:
        self.in_buf = self.ctx.alloc(self.in_elems * idsz, "ids_in")
        in_flat = self.in_buf.view(id_dtype, (self.in_elems,))
        self.in_ptrs = self.in_buf.peer_ptrs()
      else:
        self.in_buf = None
        in_flat = torch.zeros(self.in_elems, dtype=id_dtype, device=dev)
        self.in_ptrs = [in_flat.data_ptr()]
      self.in_flat = in_flat
      self.in_views = [in_flat[o:o + b * abs(h)].view(b, abs(h)) for o, h in zip(in_off, hots)]
      # row_splits of ragged inputs (int64, [b + 1] each) live in their own symmetric buffer
      if n_rag:
        n_sp = n_rag * (b + 1)
        if W > 1:
          self.split_buf = self.ctx.alloc(n_sp * 8, "ragged_splits")
          self.split_flat = self.split_buf.view(torch.int64, (n_sp,))
          self.split_ptrs = self.split_buf.peer_ptrs()
        else:
          self.split_buf = None
          self.split_flat = torch.zeros(n_sp, dtype=torch.int64, device=dev)
          self.split_ptrs = [self.split_flat.data_ptr()]
        self.split_views = {i: self.split_flat[j * (b + 1):(j + 1) * (b + 1)]
                            for i, j in rag_index.items()}
    else:
      in_off = None
      self.in_buf, self.in_flat, self.in_views, self.in_ptrs = None, None, None, []
      self.split_views = {}

    # --- model-parallel id buffer (global batch of every local input)
    my_inputs = st.input_ids_list[rank] if st.table_groups[1] else []
    col_items, pos = [], 0
    for li, k in enumerate(my_inputs):
      gi = col_group[k] if de.dp_input else li
      h = abs(hots[gi])
      col_items.append(pos)
      pos += B * h
    row_items = []
    row_inputs = st.input_groups[2] if de.dp_input else []
    for gi in row_inputs:
      row_items.append(pos)
      pos += B * abs(hots[gi])
    self.n_items = pos
    need_copy = (W > 1) or (not de.dp_input)
    self.ids_mp = torch.zeros(max(pos, 1), dtype=id_dtype, device=dev) if need_copy else None
    # global-batch CSR offsets of the ragged local inputs ([B + 1] each)
    my_ragged = [li for li, k in enumerate(my_inputs)
                 if self.ragged[col_group[k] if de.dp_input else li]]
    self.goff = torch.zeros(max(len(my_ragged) * (B + 1), 1), dtype=torch.int64, device=dev)
    goff_index = {li: j for j, li in enumerate(my_ragged)}

    def ids_ptr(item_off, gi):
      if need_copy:
        return self.ids_mp.data_ptr() + item_off * idsz
      return self.in_flat.data_ptr() + in_off[gi] * idsz

    # --- table descriptors (sorted-update path)
    tdesc = np.zeros(len(self.mp_layers), dtype=TABLE_DESC)
    key = 0
    for m, layer in enumerate(self.mp_layers):
      w = _weight(layer)
      tdesc[m]["weight"] = self._ptr(w)
      tdesc[m]["rows"] = w.shape[0]
      tdesc[m]["key_base"] = key
      tdesc[m]["width"] = w.shape[1]
      key += w.shape[0]
    self.total_rows = key
    self.tdesc_np = tdesc
    self.max_width = max([int(_weight(l).shape[1]) for l in self.mp_layers] + [1])

    # --- output layout
    tw = self.total_width
    out_bytes = max(lb * tw * csz, 16)
    if W > 1:
      self.out_buf = self.ctx.alloc(out_bytes, "emb_out")
      self.grad_buf = self.ctx.alloc(out_bytes, "emb_grad")
      self.out = self.out_buf.view(self.compute_dtype, (lb, tw))
      self.grad = self.grad_buf.view(self.compute_dtype, (lb, tw))
      self.out_ptrs = self.out_buf.peer_ptrs()
      self.grad_ptrs = self.grad_buf.peer_ptrs()
    else:
      self.out_buf = self.grad_buf = None
      self.out = torch.zeros(lb, tw, dtype=self.compute_dtype, device=dev)
      self.grad = torch.zeros(lb, tw, dtype=self.compute_dtype, device=dev)
      self.out_ptrs = [self.out.data_ptr()]
      self.grad_ptrs = [self.grad.data_ptr()]

    # --- table-parallel descriptors
    pieces = {(p.rank, p.local_input): p for p in st.output_pieces}
    cdesc = np.zeros(len(my_inputs), dtype=INPUT_DESC)
    segs, rsegs = [], []
    for li, k in enumerate(my_inputs):
      gi = col_group[k] if de.dp_input else li
      t_in_group = col_map[k]
      shard = next(s for s in st.shards[rank] if s.table == t_in_group)
      layer = de.local_embedding_layers[shard.local_table]
      w = _weight(layer)
      d = cdesc[li]
      d["table"] = self._ptr(w)
      d["ids"] = ids_ptr(col_items[li], gi)
      d["ids_off"] = 0
      d["sub_rows"] = shard.rows
      d["row_base"] = shard.row_offset
      d["width"] = shard.width
      d["hotness"] = max(hots[gi], 0)
      if self.ragged[gi]:
        if need_copy:
          d["offsets"] = self.goff.data_ptr() + goff_index[li] * (B + 1) * 8
        else:  # single rank, dp input: the staged CSR is already the global one
          d["offsets"] = self.split_views[gi].data_ptr()
      gi_global = st.input_groups[1][k]
      d["dst_col"] = self.out_cols[gi_global] + pieces[(rank, li)].col_offset
      d["combiner"] = _COMB[layer.combiner]
      d["local_table"] = shard.local_table
      d["item_off"] = col_items[li]
      if layer.combiner is None and hots[gi] != 1:
        raise ValueError("table-parallel lookups without a combiner need one id per sample")
      if de.dp_input and W > 1:
        if self.ragged[gi]:
          rsegs.append([in_off[gi], col_items[li], rag_index[gi] * (b + 1),
                        goff_index[li] * (B + 1)])
        else:
          for s in range(W):
            segs.append([s, in_off[gi], col_items[li] + s * b * hots[gi], b * hots[gi]])
    self.cdesc_np = cdesc

    # --- row-slice descriptors (partial pools land in rs_buf[d][slot = my rank])
    rdesc = np.zeros(len(row_inputs), dtype=INPUT_DESC)
    self.rs_width = 0
    self.rs_cols = []
    for j, gi in enumerate(row_inputs):
      m = st.map_groups[2][j]
      layer = de.row_layers[m]
      w = _weight(layer)
      d = rdesc[j]
      d["table"] = w.data_ptr()
      d["ids"] = ids_ptr(row_items[j], gi)
      d["id_shift"] = st.row_inputs_offsets[rank][m]
      d["sub_rows"] = w.shape[0]
      d["width"] = w.shape[1]
      d["hotness"] = hots[gi]
      d["dst_col"] = self.rs_width
      d["combiner"] = _COMB[layer.combiner]
      d["local_table"] = self.n_col_tables + m
      d["item_off"] = row_items[j]
      d["flags"] = 1
      self.rs_cols.append((gi, self.rs_width, int(w.shape[1])))
      self.rs_width += int(w.shape[1])
      for s in range(W):
        segs.append([s, in_off[gi], row_items[j] + s * b * hots[gi], b * hots[gi]])
    self.rdesc_np = rdesc
    if len(row_inputs):
      self.rs_buf = self.ctx.alloc(W * lb * self.rs_width * 4, "row_slice_partials")
      self.rs = self.rs_buf.view(torch.float32, (W, lb, self.rs_width))
      self.rs_ptrs = self.rs_buf.peer_ptrs(rank * lb * self.rs_width * 4)
    else:
      self.rs_buf = None

    # --- replicated tables: plain local lookup of the local batch
    dp_inputs = st.input_groups[0] if de.dp_input else []
    ddesc = np.zeros(len(dp_inputs), dtype=INPUT_DESC)
    for j, gi in enumerate(dp_inputs):
      m = st.map_groups[0][j]
      layer = de.dp_layers[m]
      w = _weight(layer)
      d = ddesc[j]
      d["table"] = w.data_ptr()
      d["ids"] = self.in_flat.data_ptr() + in_off[gi] * idsz
      d["sub_rows"] = w.shape[0]
      d["width"] = w.shape[1]
      d["hotness"] = hots[gi]
      d["dst_col"] = self.out_cols[gi]
      d["combiner"] = _COMB[layer.combiner]
      d["local_table"] = m
    self.ddesc_np = ddesc

    self.segs = torch.tensor(segs, dtype=torch.int64, device=dev) if segs else None
    self.max_seg = max([s[3] for s in segs]) if segs else 0
    self.rsegs = torch.tensor(rsegs, dtype=torch.int64, device=dev) if rsegs else None
    self.max_rcap = b * max([abs(h) for h, r in zip(hots, self.ragged) if r] + [0])
    self.my_ragged_mp = my_ragged if not de.dp_input else []
    self.col_items = col_items
    widths = [int(x) for x in list(cdesc["width"]) + list(rdesc["width"]) + list(ddesc["width"])]
    cols = [int(x) for x in list(cdesc["dst_col"]) + list(ddesc["dst_col"])]
    self.vec4 = all(w % 4 == 0 for w in widths) and all(c % 4 == 0 for c in cols) and \
        tw % 4 == 0 and self.rs_width % 4 == 0
    # 16-byte gradient pulls (8 columns per lane).  Faster in isolation (151 -> 109 us at 8 GPUs)
    # but the whole step regressed at 4 and 8 GPUs in a same-box A/B (0.995 -> 1.11 ms at N=4), so
    # it is opt-in until that interaction is understood: DE_B200_VEC8_PULL=1.
    self.vec8 = os.environ.get("DE_B200_VEC8_PULL", "0") == "1" and self.vec4 and \
        all(w % 8 == 0 for w in widths) and all(c % 8 == 0 for c in cols) and tw % 8 == 0 and \
        not len(row_inputs)
    # shared-memory pre-reduction of tiny one-hot tables in the SGD backward (experimental)
    self.tiny_tables = os.environ.get("DE_B200_TINY_TABLES", "0") == "1"
    # TMA bulk row copies for one-hot table-parallel / replicated lookups (experimental):
    # every input one id per sample, rows of at most 128 fp32 columns, 16-byte aligned pieces
    def bulk_ok(desc):
      return len(desc) > 0 and bool(np.all(desc["hotness"] == 1)) and \
          bool(np.all(desc["offsets"] == 0)) and bool(np.all(desc["width"] % 4 == 0)) and \
          bool(np.all(desc["width"] <= 128)) and bool(np.all(desc["dst_col"] % 4 == 0)) and \
          bool(np.all(desc["flags"] == 0))
    bulk = os.environ.get("DE_B200_LOOKUP_BULK", "0") == "1" and tw % 4 == 0
    self.bulk_c = bulk and bulk_ok(cdesc)
    self.bulk_d = bulk and bulk_ok(ddesc)
    self._upload()
    self._key = (b, hots, ids64)

  def _upload(self):
    """(Re)upload descriptor arrays; table pointers / optimizer state may have changed."""
    dev = self.device
    up = _native.upload_struct_array
    self.cdesc = up(self.cdesc_np, dev) if len(self.cdesc_np) else None
    self.rdesc = up(self.rdesc_np, dev) if len(self.rdesc_np) else None
    self.ddesc = up(self.ddesc_np, dev) if len(self.ddesc_np) else None
    self._refresh_tables()
    # one descriptor array for the backward of all model-parallel inputs
    mp = np.concatenate([self.cdesc_np, self.rdesc_np]) if len(self.rdesc_np) else self.cdesc_np
    self.mpdesc = up(mp, dev) if len(mp) else None
    self.n_mp_inputs = len(mp)

  def _refresh_tables(self):
    opt = self.de._fused_optimizer
    t = self.tdesc_np
    for m, layer in enumerate(self.mp_layers):
      t[m]["weight"] = self._ptr(_weight(layer))
      st = self.opt_state.get(m)
      t[m]["state0"] = self._ptr(st[0]) if st else 0
      t[m]["state1"] = self._ptr(st[1]) if st and len(st) > 1 else 0
    self.tdesc = _native.upload_struct_array(t, self.device) if len(t) else None
    self._tables_dirty = False
    if opt is not None:
      self.lr_t.fill_(opt["lr"])

  # ------------------------------------------------------------------ optimizer state
  def reset_optimizer_state(self):
    self.opt_state = {}
    opt = self.de._fused_optimizer
    if opt is None:
      return
    kind = opt["kind"]
    def like(w, value, shape=None):
      t = torch.full(shape or tuple(w.shape), value, dtype=torch.float32, device=w.device)
      # state of offloaded tables stays on the host (pinned, read zero-copy by the kernels)
      return t.pin_memory() if not (w.is_cuda or self.dry) else t

    for m, layer in enumerate(self.mp_layers):
      w = _weight(layer)
      if kind == "adagrad":
        self.opt_state[m] = [like(w, opt["initial_accumulator_value"])]
      elif kind == "rowwise_adagrad":
        self.opt_state[m] = [like(w, opt["initial_accumulator_value"], (w.shape[0],))]
      elif kind == "adam":
        self.opt_state[m] = [like(w, 0.0), like(w, 0.0)]
    self._tables_dirty = True

  def update_lr(self, lr: float):
    self.lr_t.fill_(lr)

  def optimizer_state_dict(self) -> Dict[str, Any]:
    return {"state": {m: [s.detach().cpu() for s in st] for m, st in self.opt_state.items()},
            "step": (self.de._fused_optimizer or {}).get("step", 0)}

  def load_optimizer_state_dict(self, state):
    if not self.opt_state:
      self.reset_optimizer_state()
    for m, tensors in state.get("state", {}).items():
      for dst, src in zip(self.opt_state[int(m)], tensors):
        dst.copy_(src)
    if self.de._fused_optimizer is not None:
      self.de._fused_optimizer["step"] = state.get("step", 0)

  # ------------------------------------------------------------------ forward
  def prepare(self, local_batch: int, hotness: Sequence[int], ids64: bool = False):
    """Allocate buffers up front (e.g. so a data loader can H2D-copy straight into
    ``input_views``); ``local_batch`` is the per-call batch of the inputs."""
    key = (int(local_batch), tuple(int(h) for h in hotness), bool(ids64))
    if key != self._key:
      self._build(*key)
    return self

  @property
  def input_views(self) -> List[torch.Tensor]:
    """``[batch, hotness]`` views of the staging buffer, one per input (dp_input mode)."""
    return self.in_views

  def _ragged_capacity(self, inputs, b: int) -> int:
    """Ids-per-sample capacity reserved for ragged inputs: user set
    (``DistributedEmbedding.ragged_capacity``) or 2x the largest mean hotness seen at build time,
    agreed on by all ranks (symmetric buffers must have one size)."""
    cap = getattr(self.de, "ragged_capacity", None)
    if cap is None:
      need = max(int(x.values.numel()) for x in inputs if isinstance(x, RaggedIds))
      cap = max(8, 2 * -(-need // max(b, 1)))
      if self.W > 1:
        t = torch.tensor([cap], dtype=torch.int64, device=self.device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX, group=self.de.group)
        cap = int(t.item())
      self.de.ragged_capacity = cap
    return int(cap)

  def stage(self, inputs):
    b = inputs[0].nrows if isinstance(inputs[0], RaggedIds) else int(inputs[0].shape[0])
    any_rag = any(isinstance(x, RaggedIds) for x in inputs)
    cap = self._ragged_capacity(inputs, b) if any_rag else 0
    hots = tuple(-cap if isinstance(x, RaggedIds) else (1 if x.dim() == 1 else int(x.shape[1]))
                 for x in inputs)
    ids64 = any((x.values if isinstance(x, RaggedIds) else x).dtype == torch.int64 for x in inputs)
    if self._key is None or self._key[0] != b or self._key[1] != hots or \
        (ids64 and not self._key[2]):
      self._build(b, hots, ids64)
    if self.de.dp_input:
      for i, (v, x) in enumerate(zip(self.in_views, inputs)):
        if isinstance(x, RaggedIds):
          n = int(x.values.numel())
          if n > v.numel():
            raise ValueError(
                f"ragged input {i} holds {n} ids but only {v.numel()} are reserved; raise "
                "DistributedEmbedding.ragged_capacity (ids per sample) on every rank")
          v.view(-1)[:n].copy_(x.values, non_blocking=True)
          self.split_views[i].copy_(x.row_splits, non_blocking=True)
        elif x.data_ptr() != v.data_ptr():
          v.copy_(x.reshape(v.shape), non_blocking=True)
    else:
      for li, x in enumerate(inputs):
        pos = self.col_items[li]
        if isinstance(x, RaggedIds):
          n = int(x.values.numel())
          if n > self.B * cap:
            raise ValueError(f"ragged input {li} holds {n} ids, capacity is {self.B * cap}")
          self.ids_mp[pos:pos + n].copy_(x.values, non_blocking=True)
          j = self.my_ragged_mp.index(li)
          self.goff[j * (self.B + 1):(j + 1) * (self.B + 1)].copy_(x.row_splits,
                                                                     non_blocking=True)
        else:
          n = x.numel()
          self.ids_mp[pos:pos + n].copy_(x.reshape(-1), non_blocking=True)

  def forward(self, inputs, concat: bool):
    self.stage(inputs)
    return self.run(concat)

  def run(self, concat: bool = True):
    """Forward on already staged inputs (see :meth:`prepare` / :attr:`input_views`)."""
    if self.de._fused_optimizer is not None and not self.opt_state and \
        self.de._fused_optimizer["kind"] != "sgd":
      self.reset_optimizer_state()
    if self._tables_dirty:
      self._refresh_tables()
    weights = [_weight(l) for l in list(self.de.dp_layers) + self.mp_layers]
    out = _FusedFn.apply(self, self._token, *weights)
    if concat:
      return out
    return list(torch.split(out, self.out_widths, dim=1))

  def _run_forward(self):
    with nvtx.range("emb_forward"):
      return self._run_forward_impl()

  def _run_forward_impl(self):
    ops, W, rank = self.ops, self.W, self.rank
    B, lb = self.B, self.lb
    bf16 = self.compute_dtype == torch.bfloat16
    if self.rs_buf is not None:
      self.rs.zero_()
    if W > 1:
      self.ctx.barrier(0)  # every rank's ids are staged (and its partial buffer is cleared)
      if self.segs is not None:
        ops.gather_segments(self.segs, self.in_ptrs, self.ids_mp, self.max_seg)
      if self.rsegs is not None:
        ops.gather_ragged(self.rsegs, self.in_ptrs, self.split_ptrs, self.ids_mp, self.goff,
                          self.lb, self.max_rcap)
    if self.ddesc is not None:
      if self.bulk_d:
        ops.lookup_fwd_bulk(self.ddesc, len(self.ddesc_np), lb, lb, lb, self.total_width, [],
                            [self.out.data_ptr()], 0, self.ids64, bf16)
      else:
        ops.lookup_fwd(self.ddesc, len(self.ddesc_np), lb, lb, lb, self.total_width, [],
                       [self.out.data_ptr()], 0, self.ids64, bf16, self.vec4)
    if self.cdesc is not None:
      if self.bulk_c:
        ops.lookup_fwd_bulk(self.cdesc, len(self.cdesc_np), B, B, lb, self.total_width, [],
                            self.out_ptrs, rank, self.ids64, bf16)
      else:
        ops.lookup_fwd(self.cdesc, len(self.cdesc_np), B, B, lb, self.total_width, [],
                       self.out_ptrs, rank, self.ids64, bf16, self.vec4)
    if self.rdesc is not None:
      ops.lookup_fwd(self.rdesc, len(self.rdesc_np), B, B, lb, self.rs_width, [], self.rs_ptrs,
                     rank, self.ids64, False, self.vec4)
    if W > 1:
      self.ctx.barrier(1)  # all pooled rows have landed in my out_buf
    if self.rs_buf is not None:
      red = self.rs.sum(dim=0)
      for gi, c0, w in self.rs_cols:
        self.out[:, self.out_cols[gi]:self.out_cols[gi] + w] = red[:, c0:c0 + w]
    return self.out

  # ------------------------------------------------------------------ backward
  def _run_backward(self, grad_out: torch.Tensor):
    ops, W, rank, de = self.ops, self.W, self.rank, self.de
    B, lb = self.B, self.lb
    bf16 = self.compute_dtype == torch.bfloat16
    if grad_out.data_ptr() != self.grad.data_ptr():
      if grad_out.dtype not in (torch.float32, torch.bfloat16) or grad_out.stride(-1) != 1:
        grad_out = grad_out.float().contiguous()
      ops.copy_cast_2d(grad_out, self.grad.data_ptr(), self.total_width, bf16, 1.0)
    if W > 1:
      self.ctx.barrier(2)  # every rank's gradient buffer is complete
    grads: List[Optional[torch.Tensor]] = []
    # replicated tables: dense local gradients (all-reduced later with the MLP gradients)
    for m, layer in enumerate(de.dp_layers):
      w = _weight(layer)
      if not w.requires_grad:
        grads.append(None)
        continue
      g = torch.zeros_like(w)
      self._dp_grad_live = g  # keeps the buffer addressable for the plan interpreter (dry_run.py)
      sel = [j for j in range(len(self.ddesc_np)) if int(self.ddesc_np[j]["local_table"]) == m]
      d = self.ddesc_np[sel].copy()
      d["table"] = g.data_ptr()
      dd = _native.upload_struct_array(d, self.device)
      ops.scatter_add_bwd(dd, len(d), lb, lb, lb, self.total_width, [], [self.grad.data_ptr()], 0,
                          1.0, 0, self.ids64, bf16, self.vec4, False)
      grads.append(g)
    grads += self._backward_mp(bf16)
    return grads

  def set_dp_grad_targets(self, targets: Optional[Sequence[torch.Tensor]]):
    """Persistent dense-gradient buffers of the replicated tables, one fp32 ``[rows, width]``
    tensor per ``de.dp_layers`` entry (e.g. slices of a flat all-reduce bucket).  The caller
    zeroes them every step; :meth:`backward_inplace` accumulates the local-batch gradient into
    them, so a hand-scheduled step can all-reduce and apply them with its dense parameters."""
    if targets is not None:
      targets = list(targets)
      if len(targets) != len(self.de.dp_layers):
        raise ValueError(f"expected {len(self.de.dp_layers)} targets, got {len(targets)}")
      for t, layer in zip(targets, self.de.dp_layers):
        w = _weight(layer)
        if t.dtype != torch.float32 or tuple(t.shape) != tuple(w.shape) or not t.is_contiguous():
          raise ValueError("dp gradient targets must be contiguous fp32 tensors of the table shape")
    self._dp_targets = targets
    self._dp_target_desc = None

  def _scatter_dp_grads(self, bf16: bool):
    """Local-batch gradient of every replicated table into its persistent target (one launch)."""
    if not len(self.ddesc_np):
      return
    key = (self._key, tuple(t.data_ptr() for t in self._dp_targets))
    if self._dp_target_desc is None or self._dp_target_desc[0] != key:
      d = self.ddesc_np.copy()
      for j in range(len(d)):
        d[j]["table"] = self._dp_targets[int(d[j]["local_table"])].data_ptr()
      self._dp_target_desc = (key, _native.upload_struct_array(d, self.device), len(d))
    _, dd, n = self._dp_target_desc
    self.ops.scatter_add_bwd(dd, n, self.lb, self.lb, self.lb, self.total_width, [],
                             [self.grad.data_ptr()], 0, 1.0, 0, self.ids64, bf16, self.vec4, False)

  def backward_inplace(self):
    """Backward when the gradient was already written into ``self.grad`` (e.g. by the fused
    interaction-backward kernel): barrier + fused table update; replicated tables accumulate
    their dense gradient into the targets given to :meth:`set_dp_grad_targets`."""
    bf16 = self.compute_dtype == torch.bfloat16
    if self.W > 1:
      self.ctx.barrier(2)
    if len(self.de.dp_layers):
      if getattr(self, "_dp_targets", None) is None:
        raise RuntimeError("backward_inplace needs set_dp_grad_targets() for replicated tables")
      self._scatter_dp_grads(bf16)
    self._backward_mp(bf16)

  def _backward_mp(self, bf16: bool) -> List[Optional[torch.Tensor]]:
    with nvtx.range("emb_backward_update"):
      return self._backward_mp_impl(bf16)

  def _backward_mp_impl(self, bf16: bool) -> List[Optional[torch.Tensor]]:
    ops, de = self.ops, self.de
    n_mp = len(self.mp_layers)
    if self.mpdesc is None or not any(_weight(l).requires_grad for l in self.mp_layers):
      return [None] * n_mp
    opt = de._fused_optimizer
    if opt is not None and opt["kind"] != "sgd" and not self.opt_state:
      self.reset_optimizer_state()
    if self._tables_dirty:
      self._refresh_tables()
    B, lb = self.B, self.lb
    # row-slice partial gradients: the gradient of an input lives in grad_buf at the input's
    # output columns for *both* groups, but row descriptors carry rs_buf columns -> patch once
    if opt is not None and opt["kind"] == "sgd" and not opt.get("deterministic", False) and \
        not self.has_offload:
      if self.tiny_tables and self.vec4:
        # experimental: one-hot inputs of tables with <= 64 rows are pre-reduced in shared memory
        main, n_main, tiny, n_tiny, t_rows, t_width = self._bwd_desc_split()
        if n_main:
          ops.scatter_add_bwd(main, n_main, B, B, lb, self.total_width, [], self.grad_ptrs,
                              self.rank, -de.mp_grad_scale, self.lr_t.data_ptr(), self.ids64,
                              bf16, self.vec4, self.vec8 and self.W > 1)
        if n_tiny:
          ops.tiny_scatter_add_bwd(tiny, n_tiny, B, B, lb, self.total_width, [], self.grad_ptrs,
                                   -de.mp_grad_scale, self.lr_t.data_ptr(), self.ids64, bf16,
                                   t_rows, t_width)
        return [None] * n_mp
      ops.scatter_add_bwd(self._bwd_desc(), self.n_mp_inputs, B, B, lb, self.total_width, [],
                          self.grad_ptrs, self.rank, -de.mp_grad_scale, self.lr_t.data_ptr(),
                          self.ids64, bf16, self.vec4, self.vec8 and self.W > 1)
      return [None] * n_mp
    keys, items, seg, n_unique = ops.sort_items(self._bwd_desc(), self.tdesc, n_mp,
                                                self.n_mp_inputs, B, B, [], self.ids64,
                                                self.n_items, self.total_rows, self.any_ragged)
    if opt is not None:
      opt["step"] += 1
      t = opt["step"]
      bias1 = 1.0 - opt["beta1"]**t
      bias2 = 1.0 - opt["beta2"]**t
      ops.segment_update(self._bwd_desc(), self.tdesc, n_mp, B, lb, self.total_width,
                         self.grad_ptrs, keys, items, seg, n_unique, _OPT_KIND[opt["kind"]],
                         opt["lr"], opt["eps"], opt["beta1"], opt["beta2"], bias1, bias2,
                         de.mp_grad_scale, opt["weight_decay"], self.lr_t.data_ptr(), None, None,
                         self.max_width, bf16, self.vec4, self._balanced_scratch())
      return [None] * n_mp
    # no fused optimizer: materialise deduplicated sparse gradients (reference semantics)
    emit_keys = torch.empty(self.n_items, dtype=torch.int64, device=self.device)
    emit_rows = torch.empty(self.n_items, self.max_width, dtype=torch.float32, device=self.device)
    ops.segment_update(self._bwd_desc(), self.tdesc, n_mp, B, lb, self.total_width, self.grad_ptrs,
                       keys, items, seg, n_unique, _native.OPT_EMIT, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0,
                       de.mp_grad_scale, 0.0, 0, emit_keys, emit_rows, self.max_width, bf16,
                       self.vec4, None)
    nu = int(n_unique.item())
    emit_keys, emit_rows = emit_keys[:nu], emit_rows[:nu]
    bases = [int(x) for x in self.tdesc_np["key_base"]] + [self.total_rows]
    bounds = torch.searchsorted(emit_keys, torch.tensor(bases, device=self.device)).tolist()
    out = []
    for m, layer in enumerate(self.mp_layers):
      w = _weight(layer)
      lo, hi = bounds[m], bounds[m + 1]
      ids = (emit_keys[lo:hi] - bases[m]).unsqueeze(0).to(w.device)
      # a fresh [nnz, width] buffer with canonical strides: for nnz == 1 ``.contiguous()`` is a
      # no-op on the padded view (row stride max_width) and PyTorch's sparse -> dense kernels
      # then address the destination with that stride (heap overflow found by the plan fuzzer)
      rows = torch.empty(hi - lo, w.shape[1], dtype=torch.float32, device=emit_rows.device)
      rows.copy_(emit_rows[lo:hi, :w.shape[1]])
      rows = rows.to(w.device)
      out.append(torch.sparse_coo_tensor(ids, rows, size=tuple(w.shape), is_coalesced=True,
                                         check_invariants=False))
    return out

  def _balanced_scratch(self):
    """Zero-initialised scratch rows of the occurrence-balanced update (kept zero by the
    kernels); None selects the per-unique-row kernel (wide tables)."""
    if not self.vec4 or self.max_width > 128 or self.n_items == 0:
      return None
    sw = (self.max_width + 3) // 4 * 4
    need = ((self.n_items + 31) // 32) * sw
    cur = getattr(self, "_scratch", None)
    if cur is None or cur.numel() < need:
      self._scratch = torch.zeros(need, dtype=torch.float32, device=self.device)
    return self._scratch

  def _bwd_desc_split(self):
    """(main descs, n, tiny descs, n, max tiny rows, max tiny width): the backward descriptors
    split into inputs served by the RED scatter kernel and one-hot inputs of tiny tables."""
    if getattr(self, "_split_cache", None) is not None and self._split_key == self._key:
      return self._split_cache
    self._bwd_desc()
    mp = self._bwd_desc_np
    tiny = (mp["hotness"] == 1) & (mp["offsets"] == 0) & (mp["sub_rows"] <= 64) & \
        (mp["width"] % 4 == 0) & (mp["sub_rows"] * mp["width"] * 4 <= 96 * 1024)
    main_np, tiny_np = mp[~tiny], mp[tiny]
    main = _native.upload_struct_array(main_np, self.device) if len(main_np) else None
    tin = _native.upload_struct_array(tiny_np, self.device) if len(tiny_np) else None
    t_rows = int(tiny_np["sub_rows"].max()) if len(tiny_np) else 0
    t_width = int(tiny_np["width"].max()) if len(tiny_np) else 0
    self._split_cache = (main, len(main_np), tin, len(tiny_np), t_rows, t_width)
    self._split_key = self._key
    return self._split_cache

  def _bwd_desc(self):
    """Backward descriptors: same as forward but row-slice inputs read their gradient at the
    input's final output columns of grad_buf."""
    if getattr(self, "_bwd_desc_cache", None) is not None and self._bwd_key == self._key:
      return self._bwd_desc_cache
    mp = np.concatenate([self.cdesc_np, self.rdesc_np]) if len(self.rdesc_np) else \
        self.cdesc_np.copy()
    n_c = len(self.cdesc_np)
    for j, (gi, _, _) in enumerate(self.rs_cols):
      mp[n_c + j]["dst_col"] = self.out_cols[gi]
    self._bwd_desc_np = mp
    self._bwd_desc_cache = _native.upload_struct_array(mp, self.device)
    self._bwd_key = self._key
    return self._bwd_desc_cache

"""Hand-scheduled DLRM training step (no autograd on the hot path) + whole-step CUDA graph.

The generic :class:`DLRM` module is convenient but pays for autograd bookkeeping, autocast weight
casts, index-based interaction and ~150 tiny launches per step.  ``DLRMTrainStep`` runs the same
math as one static schedule over preallocated buffers:

* dense parameters live in one flat fp32 master buffer with a bf16 shadow (one fused kernel does
  SGD + re-cast + gradient zeroing); gradients live in one flat buffer (symmetric memory when
  world > 1) that the one-shot NVLink all-reduce kernel reduces in place;
* MLP layers: cuBLASLt bf16 GEMMs with fused bias+ReLU epilogues forward, plain GEMMs backward
  (K padded 13->16 and 479->480 so the tcgen05 library kernels are eligible), fused
  ReLU-backward+bias-gradient kernel;
* dot interaction forward/backward: tensor-core kernels; the forward's head waits for the
  embedding owners' "output ready" signals, the backward pushes every piece of the embedding
  gradient straight into its owner's receive buffer over NVLink and signals "gradient ready";
* final layer + BCE loss + their backward: one kernel;
* embedding forward/backward: the fused P2P engine (``parallel/fused.py``);
* the whole step is captured in a CUDA graph and replayed (launch bound otherwise).

Same model/optimizer as the reference example (examples/dlrm/main.py:76-209): SGD, shared lr.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch import nn

from ..ops import _native
from ..parallel.comm import CommContext
from ..parallel.fused import FusedEngine
from ..utils import nvtx
from ..utils.lr_schedule import LearningRateScheduler
from .dlrm import DLRM


def _pad8(n: int) -> int:
  return (n + 7) // 8 * 8


class _Layer:
  """One linear layer's slices of the flat buffers."""

  def __init__(self, lin: nn.Linear, relu: bool):
    self.lin = lin
    self.relu = relu
    self.out_f, self.in_f = lin.weight.shape
    self.in_pad = _pad8(self.in_f)
    self.w_off = self.b_off = 0
    self.w_numel = self.out_f * self.in_pad
    self.b_numel = _pad8(self.out_f)


class DLRMTrainStep:
  """Static-schedule training step for :class:`DLRM` on the fused embedding back end."""

  def __init__(self, model: DLRM, lr: float = 24.0, embedding_optimizer: str = "sgd",
               scheduler: Optional[LearningRateScheduler] = None, use_cuda_graph: bool = True,
               embedding_optimizer_kwargs: Optional[dict] = None, overlap: bool = True,
               gemm: str = "cublas"):
    if gemm not in ("cublas", "fused_dgrad", "tcgen05", "tcgen05_pair"):
      raise ValueError("gemm must be cublas | fused_dgrad | tcgen05 | tcgen05_pair")
    # cublas: cuBLASLt everywhere.  fused_dgrad: forward/wgrad on cuBLASLt, dgrad on the
    # first-party tcgen05 kernel with the ReLU-backward mask + bias gradient fused in its epilogue.
    # tcgen05: forward layers on the first-party kernel as well.  tcgen05_pair: same, with the
    # CTA-pair (cta_group::2) kernel for layers at least 256 wide.
    self.gemm = gemm
    self.model = model
    self.emb = model.embedding
    if self.emb.backend != "fused":
      raise ValueError("DLRMTrainStep needs the fused embedding back end")
    self.dev = self.emb.device
    self.ops = _native.require()
    self.world = self.emb.world_size
    self.ctx = CommContext.for_group(self.emb.group, self.dev)
    self.scheduler = scheduler
    self.use_cuda_graph = use_cuda_graph
    self.overlap = overlap
    self.emb.set_optimizer(embedding_optimizer, lr=lr, **(embedding_optimizer_kwargs or {}))
    if self.emb._engine is None:
      self.emb._engine = FusedEngine(self.emb)
    self.engine: FusedEngine = self.emb._engine
    self.n_emb = len(model.table_sizes)
    self.dim = model.embedding_dim
    # gradient all-to-all through local staging + a streaming copy kernel next to the interaction
    # backward (DE_B200_STREAM_PUSH=0: the interaction backward stores into peer memory itself)
    # measured: 0.705 vs 0.738 ms per step at 8 GPUs with it, but 1.41 vs 1.33 ms at 2 GPUs (the
    # copy kernel then competes with an interaction backward that keeps every SM busy), hence
    # on from 4 GPUs; DE_B200_STREAM_PUSH=0/1 overrides
    sp_env = os.environ.get("DE_B200_STREAM_PUSH", "auto")
    self._stream_push = self.world > 1 and (sp_env == "1" or (sp_env == "auto" and self.world >= 4))
    self._push_stream = torch.cuda.Stream(device=self.dev) if self._stream_push else None
    self._push_ready = torch.cuda.Event() if self._stream_push else None

    lins_b = [m for m in model.bottom_mlp.net if isinstance(m, nn.Linear)]
    lins_t = [m for m in model.top_mlp.net if isinstance(m, nn.Linear)]
    self.bottom = [_Layer(l, True) for l in lins_b]
    self.top = [_Layer(l, True) for l in lins_t[:-1]]
    self.head = _Layer(lins_t[-1], False)
    if self.head.out_f != 1:
      raise ValueError("the top MLP must end in a single logit")
    layers = self.bottom + self.top + [self.head]
    # replicated (data-parallel) embedding tables live in the flat dense buffers too: their
    # local-batch gradient is scattered into the gradient bucket, all-reduced with the MLP
    # gradients and applied by the same fused SGD kernel.  They come first so that they belong to
    # the bucket that is reduced last (DE_B200_AR_OVERLAP).  Validated on 2 and 8 GPUs
    # (tests/test_dist_gpu.py: replicated-table cases); bench.py replicates the < 2500-row tables.
    self._dp_slots = []
    pos = 0
    if len(self.emb.dp_layers):
      if embedding_optimizer != "sgd":
        raise ValueError("replicated tables in the fast step are updated by the dense SGD "
                         "kernel: use embedding_optimizer='sgd' or data_parallel_threshold=None")
      for layer in self.emb.dp_layers:
        w = layer.embeddings
        self._dp_slots.append((pos, tuple(w.shape)))
        pos += _pad8(w.numel())
    for L in layers:
      L.w_off = pos
      pos += L.w_numel
      L.b_off = pos
      pos += L.b_numel
    self.n_flat = pos
    dev = self.dev
    self.p32 = torch.zeros(pos, dtype=torch.float32, device=dev)
    self.p16 = torch.zeros(pos, dtype=torch.bfloat16, device=dev)
    if self.world > 1:
      # prefer an NVSwitch multicast mapping (in-switch reduction), else plain peer mappings
      self.gsym = self.ctx.alloc_multicast(pos * 4, "dense_grads") or \
          self.ctx.alloc(pos * 4, "dense_grads")
      self.allreduce_kind = "nvls_multimem" if getattr(self.gsym, "mc_ptr", 0) else "p2p"
      self.g32 = self.gsym.view(torch.float32, (pos,))
    else:
      self.gsym = None
      self.allreduce_kind = "none"
      self.g32 = torch.zeros(pos, dtype=torch.float32, device=dev)
    # move the module parameters into the flat master buffer (strided views keep the module usable)
    with torch.no_grad():
      dp_targets = []
      for layer, (off, shape) in zip(self.emb.dp_layers, self._dp_slots):
        n = shape[0] * shape[1]
        view = self.p32[off:off + n].view(shape)
        view.copy_(layer.embeddings.data)
        layer.embeddings.data = view
        dp_targets.append(self.g32[off:off + n].view(shape))
      if self._dp_slots:
        self.engine.set_dp_grad_targets(dp_targets)
        self.engine._tables_dirty = True
        self.engine._key = None  # descriptors built earlier point at the old table storage
      for L in layers:
        wv = self.p32[L.w_off:L.w_off + L.w_numel].view(L.out_f, L.in_pad)
        wv[:, :L.in_f].copy_(L.lin.weight)
        L.lin.weight.data = wv[:, :L.in_f]
        bv = self.p32[L.b_off:L.b_off + L.out_f]
        bv.copy_(L.lin.bias)
        L.lin.bias.data = bv
        L.w16 = self.p16[L.w_off:L.w_off + L.w_numel].view(L.out_f, L.in_pad)
        L.b16 = self.p16[L.b_off:L.b_off + L.out_f]
        L.gw = self.g32[L.w_off:L.w_off + L.w_numel].view(L.out_f, L.in_pad)
        L.gb = self.g32[L.b_off:L.b_off + L.b_numel]
        L.w16T = torch.empty(L.in_pad, L.out_f, dtype=torch.bfloat16, device=dev)
      self.p16.copy_(self.p32)
      self._refresh_transposes()
    self.lr_t = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
    self.engine.share_lr(self.lr_t)  # dense SGD and the fused embedding update read one word
    self.lr = float(lr)
    self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
    self._batch = None
    self._graph = None
    self._side = torch.cuda.Stream(device=dev) if overlap else None
    # weight-gradient GEMMs are off the critical path (head -> dgrads -> interaction -> embedding
    # backward): they run on a second side stream and join before the all-reduce
    # (measured: +1.5 % at local batch 65536, but at small local batches the full-GPU cuBLAS
    # kernels of the two streams interleave and stretch the critical chain: 0.66 -> 0.81 ms at 8
    # GPUs, so it is only enabled for large local batches; DE_B200_WGRAD_STREAM=0/1 overrides)
    self._wgrad_overlap = os.environ.get("DE_B200_WGRAD_STREAM", "auto")
    self._wstream = torch.cuda.Stream(device=dev) if overlap else None
    # The top-MLP + head gradients (93 % of the dense parameters, complete as soon as the top MLP
    # backward is done) are all-reduced on a third stream while the interaction backward, the
    # embedding exchange and the bottom MLP backward run; only the small bottom-MLP bucket is
    # reduced at the end.  The overlapped kernel is capped at 32 blocks so that its flag spins
    # cannot starve the kernels the peers wait for.  DE_B200_AR_OVERLAP=0/1 overrides the default
    # (on from 4 GPUs).
    ar_env = os.environ.get("DE_B200_AR_OVERLAP", "auto")
    # measured: +1 % at 8 GPUs, -4 % at 2 GPUs (there the large local batch keeps every SM busy
    # and the overlapped kernel only steals from the interaction backward)
    ar_on = ar_env == "1" or (ar_env == "auto" and self.world >= 4)
    self._ar_stream = torch.cuda.Stream(device=dev) if (overlap and self.world > 1 and ar_on) \
        else None

  def _refresh_transposes(self):
    """K-major copies of W^T for the dgrad GEMMs (2.4 M elements, a few microseconds)."""
    if self.gemm == "cublas":
      return
    for L in self.bottom[1:] + self.top[1:]:
      L.w16T.copy_(L.w16.t())

  def _linear_fwd(self, L, x):
    if self.gemm in ("tcgen05", "tcgen05_pair"):
      pair = self.gemm == "tcgen05_pair" and L.out_f >= 256
      self.ops.gemm_tn_bias_act(x, L.w16, L.b16, L.y, True, 512 if pair else 0)
    else:
      torch._addmm_activation(L.b16, x, L.w16.t(), out=L.y)
    return L.y

  def _wgrad(self, L, x):
    """gw = dy^T @ x (fp32), on the weight-gradient stream."""
    use = self._wstream is not None and (
        self._wgrad_overlap == "1" or
        (self._wgrad_overlap == "auto" and (self._batch or 0) >= 49152))
    if not use:
      torch.mm(L.dy.t(), x, out_dtype=torch.float32, out=L.gw)
      return
    self._w_used = True
    self._wstream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(self._wstream):
      torch.mm(L.dy.t(), x, out_dtype=torch.float32, out=L.gw)

  def _dgrad_relu(self, L, x_below, dx, gb_below):
    """dx = (dy @ W) * (x_below > 0); gb_below += colsum(dx)."""
    if self.gemm == "cublas":
      torch.mm(L.dy, L.w16, out=dx)
      self.ops.relu_bwd_bias(dx, x_below, gb_below)
    else:
      self.ops.gemm_dgrad_relu_bias(L.dy, L.w16T, x_below, dx, gb_below, 0)

  # ------------------------------------------------------------------ buffers
  def _alloc(self, b: int):
    dev, bf = self.dev, torch.bfloat16
    if self._stream_push:
      # chunks of >= 1024 samples, at most 16 of them: the last chunk's transfer is the only part
      # of the exchange that does not overlap the interaction backward
      self.engine.enable_streamed_push(max(1024, -(-b // 16)))
    self.engine.prepare(b, [1] * self.n_emb, ids64=False)
    self.cat_stage = self.engine.in_flat[:self.n_emb * b].view(self.n_emb, b)
    self.num_in = torch.zeros(b, self.bottom[0].in_f, dtype=torch.float32, device=dev)
    self.lab_in = torch.zeros(b, dtype=torch.float32, device=dev)
    self.x0 = torch.zeros(b, self.bottom[0].in_pad, dtype=bf, device=dev)
    for L in self.bottom + self.top:
      L.y = torch.empty(b, L.out_f, dtype=bf, device=dev)
      L.dy = torch.empty(b, L.out_f, dtype=bf, device=dev)
    self.z = torch.zeros(b, self.top[0].in_pad, dtype=bf, device=dev)
    self.dz = torch.empty(b, self.top[0].in_pad, dtype=bf, device=dev)
    self._batch = b
    self._graph = None
    # double-buffered device staging for the asynchronous input pipeline (prefetch())
    self._stage = [(torch.zeros_like(self.cat_stage), torch.zeros_like(self.num_in),
                    torch.zeros_like(self.lab_in)) for _ in range(2)]
    self._slot_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    self._slot_host = torch.tensor([[0], [1]], dtype=torch.int32).pin_memory()
    self._h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
    self._consumed = [torch.cuda.Event(), torch.cuda.Event()]
    self._prefetched = 0   # batches handed to prefetch()
    self._consumed_n = 0   # batches run
    self._use_stage = False

  # ------------------------------------------------------------------ the step
  def _forward(self):
    ops = self.ops
    if self._use_stage:
      a, b_ = self._stage
      ops.select_copy([a[0], a[1], a[2]], [b_[0], b_[1], b_[2]],
                      [self.cat_stage, self.num_in, self.lab_in], self._slot_dev)
    # the embedding exchange (id push, gather + NVLink push of the pooled rows; all signalling
    # folded into those kernels) runs on the side stream while the bottom MLP runs on the main
    # stream; they meet at the interaction, whose head waits for the owners' "output ready"
    eng = self.engine
    if self._side is not None:
      self._side.wait_stream(torch.cuda.current_stream())
      with torch.cuda.stream(self._side):
        eng.launch_forward()
    ops.cast_pad(self.num_in, self.x0)
    x = self.x0
    for L in self.bottom:
      x = self._linear_fwd(L, x)
    if self._side is not None:
      torch.cuda.current_stream().wait_stream(self._side)
    else:
      eng.launch_forward()
    if eng.out_needs_reduce:  # multi-hot row slices: partial pools are summed first
      eng.wait_output()
      ops.interact_fwd(x, eng.out, self.n_emb, self.z, [])
    else:
      ops.interact_fwd(x, eng.out, self.n_emb, self.z, eng.sync_out_wait())
    x = self.z
    for L in self.top:
      x = self._linear_fwd(L, x)

  def _backward(self):
    ops, eng = self.ops, self.engine
    b = self._batch
    last = self.top[-1]
    self.loss.zero_()
    H = self.head
    # final layer + loss + their backward; also the ReLU mask and bias gradient of the layer below
    ops.head_loss(last.y, H.w16.view(-1), H.b16, self.lab_in, 1.0 / b, last.dy,
                  H.gw.view(-1), H.gb, last.gb, self.loss, None)
    # top MLP backward
    for i in range(len(self.top) - 1, -1, -1):
      L = self.top[i]
      x = self.top[i - 1].y if i > 0 else self.z
      dx = self.top[i - 1].dy if i > 0 else self.dz
      self._wgrad(L, x)
      if i > 0:
        self._dgrad_relu(L, x, dx, self.top[i - 1].gb)
      else:
        torch.mm(L.dy, L.w16, out=dx)
    if self._ar_stream is not None:
      ar = self._ar_stream
      ar.wait_stream(torch.cuda.current_stream())
      if getattr(self, "_w_used", False):
        ar.wait_stream(self._wstream)
      off = self.top[0].w_off  # flat layout: bottom layers | top layers | head
      with torch.cuda.stream(ar):
        self.ctx.allreduce_(self.gsym, self.n_flat - off, torch.float32, scale=1.0 / self.world,
                            byte_offset=off * 4, max_blocks=32)
    # interaction backward: every piece of the embedding gradient is stored straight into the
    # receive buffer of the rank that owns the table (slice) - the gradient all-to-all rides on
    # the kernel's epilogue stores - and its tail signals "gradient ready" to the owners
    hb = self.bottom[-1]
    pushed = eng.streamed_push
    if pushed:
      # pieces of remote owners are staged locally; the copy kernel (own stream, a few blocks)
      # forwards every finished chunk over NVLink while the interaction backward keeps computing.
      # The producer is launched FIRST: the copy kernel spins on the producer's progress, so it
      # must never sit in front of it in a hardware queue the two streams happen to share
      # (streams alias onto CUDA_DEVICE_MAX_CONNECTIONS queues) - launched second it at worst
      # runs after the producer instead of next to it.
      eng.push_counters.zero_()
      self._push_ready.record(torch.cuda.current_stream())
      ops.interact_bwd(hb.y, eng.out, self.n_emb, self.dz, hb.dy, 0, 0, 1.0, eng.routes_stage,
                       len(eng.routes_stage_np), [], eng.push_counters, eng.push_chunk_rows)
      self._push_stream.wait_event(self._push_ready)
      with torch.cuda.stream(self._push_stream):
        eng.launch_streamed_push()
    else:
      ops.interact_bwd(hb.y, eng.out, self.n_emb, self.dz, hb.dy, 0, 0, 1.0, eng.routes_all,
                       len(eng.routes_all_np), eng.sync_grad_signal(), None, 0)
    # embedding exchange + fused table update, overlapped with the bottom MLP backward
    if self._side is not None:
      self._side.wait_stream(torch.cuda.current_stream())
      if pushed:  # the update's head waits for every rank's copy kernel: ours must be done first
        self._side.wait_stream(self._push_stream)
      with torch.cuda.stream(self._side):
        eng.backward_inplace()
    else:
      if pushed:
        torch.cuda.current_stream().wait_stream(self._push_stream)
      eng.backward_inplace()
    ops.relu_bwd_bias(hb.dy, hb.y, hb.gb)
    for i in range(len(self.bottom) - 1, -1, -1):
      L = self.bottom[i]
      x = self.bottom[i - 1].y if i > 0 else self.x0
      self._wgrad(L, x)
      if i > 0:
        self._dgrad_relu(L, x, self.bottom[i - 1].dy, self.bottom[i - 1].gb)
    # dense gradient all-reduce (one NVLink kernel, averaged) + fused SGD / re-cast / zero
    if self._dp_slots and self._side is not None:
      # the replicated tables' gradients are produced by the embedding backward on the side stream
      torch.cuda.current_stream().wait_stream(self._side)
    if getattr(self, "_w_used", False):  # join only a stream that took part in this step
      torch.cuda.current_stream().wait_stream(self._wstream)
      self._w_used = False
    if self.world > 1:
      if self._ar_stream is not None:
        # bottom-MLP bucket, behind the top bucket on the same stream (one flag channel)
        ar = self._ar_stream
        ar.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(ar):
          self.ctx.allreduce_(self.gsym, self.top[0].w_off, torch.float32, scale=1.0 / self.world)
        torch.cuda.current_stream().wait_stream(ar)
      else:
        self.ctx.allreduce_(self.gsym, self.n_flat, torch.float32, scale=1.0 / self.world)
    ops.dense_sgd(self.p32, self.p16, self.g32, self.lr_t, 1.0)
    self._refresh_transposes()
    if self._side is not None:
      torch.cuda.current_stream().wait_stream(self._side)

  def _step_impl(self):
    with nvtx.range("dlrm_forward"):
      self._forward()
    with nvtx.range("dlrm_backward_update"):
      self._backward()

  def set_lr(self, lr: float):
    self.lr = float(lr)
    self.lr_t.fill_(self.lr)  # shared with the embedding engine: one fill per schedule step
    if self.emb._fused_optimizer is not None:
      self.emb._fused_optimizer["lr"] = self.lr

  def load_batch(self, numerical, categorical, labels):
    """Copy one batch into the static input buffers (host pinned or device tensors).
    ``categorical``: ``[n_features, batch]`` tensor (feature major) or list of ``[batch]``."""
    b = int(numerical.shape[0])
    if b != self._batch:
      self._alloc(b)
    self.num_in.copy_(numerical, non_blocking=True)
    self.lab_in.copy_(labels.reshape(-1), non_blocking=True)
    if isinstance(categorical, (list, tuple)):
      for v, c in zip(self.engine.in_views, categorical):
        v.copy_(c.reshape(v.shape), non_blocking=True)
    else:
      self.cat_stage.copy_(categorical, non_blocking=True)

  def prefetch(self, numerical, categorical, labels):
    """Asynchronous input pipeline: enqueue the H2D copy of the *next* batch (pinned host
    tensors; ``categorical`` as ``[n_features, batch]`` int32) on the copy stream while the
    current step runs.  Consume with :meth:`run_prefetched` in the same order."""
    b = int(numerical.shape[0])
    if b != self._batch:
      self._alloc(b)
    if not self._use_stage:
      self._use_stage = True
      self._graph = None  # the staged schedule starts with select_copy
      self._copy_stream = torch.cuda.Stream(device=self.dev)
    slot = self._prefetched & 1
    cs = self._copy_stream
    if self._prefetched >= 2:
      cs.wait_event(self._consumed[slot])  # the step that used this slot has been enqueued & done
    with torch.cuda.stream(cs):
      st = self._stage[slot]
      st[0].copy_(categorical, non_blocking=True)
      st[1].copy_(numerical, non_blocking=True)
      st[2].copy_(labels.reshape(-1), non_blocking=True)
      self._h2d_done[slot].record(cs)
    self._prefetched += 1

  def run_prefetched(self) -> torch.Tensor:
    """Run one step on the oldest prefetched batch."""
    assert self._consumed_n < self._prefetched, "call prefetch() first"
    slot = self._consumed_n & 1
    main = torch.cuda.current_stream()
    main.wait_event(self._h2d_done[slot])
    self._slot_dev.copy_(self._slot_host[slot], non_blocking=True)
    loss = self.run()
    self._consumed[slot].record(main)
    self._consumed_n += 1
    return loss

  def run(self) -> torch.Tensor:
    """Run one step on the loaded batch; returns the (device) mean loss of the local batch."""
    if self.scheduler is not None:
      self.set_lr(self.scheduler.step())
    if not self.use_cuda_graph:
      self._step_impl()
      return self.loss
    if self._graph is None:
      # warm up on a side stream (cuBLAS handles / workspaces), then capture.  The learning rate
      # is zero while warming up so the extra passes leave the weights untouched.
      self.lr_t.zero_()
      self.engine.dry_updates(True)  # optimizer state (Adagrad / Adam) stays untouched as well
      s = torch.cuda.Stream(device=self.dev)
      s.wait_stream(torch.cuda.current_stream())
      with torch.cuda.stream(s):
        for _ in range(2):
          self._step_impl()
      torch.cuda.current_stream().wait_stream(s)
      torch.cuda.synchronize()
      self.engine.dry_updates(False)
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g):
        self._step_impl()
      self._graph = g
      self.set_lr(self.lr)
    self._graph.replay()
    return self.loss

  def step(self, numerical, categorical, labels) -> torch.Tensor:
    self.load_batch(numerical, categorical, labels)
    return self.run()

"""Hybrid-parallel training step: fused model-parallel embedding update + data-parallel dense
update, with an optional whole-step CUDA graph.

One call of :meth:`HybridTrainer.step` does what the reference's ``train_step`` does
(examples/dlrm/main.py:200-209): forward, loss, backward through the embedding exchange, dense
gradient all-reduce, optimizer apply.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
from torch import nn

from ..parallel.comm import CommContext
from ..parallel.hybrid import GradBucket
from ..utils.lr_schedule import LearningRateScheduler


class HybridTrainer:
  """Owns the dense optimizer state and the gradient bucket of a model whose embeddings are a
  :class:`DistributedEmbedding` with a fused optimizer.

  Args:
    model: module taking ``(numerical, categorical)``; must expose ``embedding``
      (DistributedEmbedding) and ``dense_parameters()``.
    lr: learning rate (dense SGD and fused embedding optimizer share it like the reference).
    embedding_optimizer: ``sgd`` | ``adagrad`` | ``rowwise_adagrad`` | ``adam``.
    scheduler: optional :class:`LearningRateScheduler`.
  """

  def __init__(self, model: nn.Module, lr: float = 24.0, embedding_optimizer: str = "sgd",
               scheduler: Optional[LearningRateScheduler] = None, momentum: float = 0.0,
               embedding_optimizer_kwargs: Optional[dict] = None,
               loss_fn: Optional[Callable] = None):
    self.model = model
    self.emb = model.embedding
    self.emb.set_optimizer(embedding_optimizer, lr=lr, **(embedding_optimizer_kwargs or {}))
    self.dense_params: List[nn.Parameter] = list(model.dense_parameters())
    self.scheduler = scheduler
    self.lr = lr
    dev = self.dense_params[0].device
    self.ctx = CommContext.default(dev) if dev.type == "cuda" else None
    self.world = self.emb.world_size
    self.bucket = GradBucket(self.dense_params, self.ctx if self.world > 1 else None)
    self.bucket.attach()
    self.opt = torch.optim.SGD(self.dense_params, lr=lr, momentum=momentum,
                               foreach=dev.type == "cuda")
    self.loss_fn = loss_fn or nn.BCEWithLogitsLoss()

  def set_lr(self, lr: float):
    self.lr = lr
    for g in self.opt.param_groups:
      g["lr"] = lr
    self.emb.set_learning_rate(lr)

  def step(self, numerical, categorical, labels, staged: bool = False) -> torch.Tensor:
    if self.scheduler is not None:
      self.set_lr(self.scheduler.step())
    self.bucket.zero_()
    logits = self.model(numerical, categorical, staged=staged) if staged else \
        self.model(numerical, categorical)
    loss = self.loss_fn(logits.float(), labels)
    loss.backward()  # embedding tables are updated inside the backward kernels
    self.bucket.gather_grads_()
    self.bucket.allreduce_(average=True)
    self.opt.step()
    return loss.detach()

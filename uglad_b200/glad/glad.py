"""glad(): the unrolled alternating-minimisation model (reference: uglad/glad/glad.py:74-150),
executed by the sm_100a kernels behind uglad_glad_forward / uglad_glad_backward."""
from __future__ import annotations

import torch.nn as nn
from torch import Tensor
from torch.optim import Adam, Optimizer

from .. import ops


def get_optimizers(model_glad: nn.Module, lr_glad: float = 0.002, use_optimizer: str = "adam",
                   capturable: bool = False) -> Optimizer:
    """glad.py:11-36.  `capturable`: keep Adam's step counter on the device so that the whole epoch can be
    captured in a CUDA graph (ops.GraphedStep)."""
    if use_optimizer == "adam":
        return Adam(model_glad.parameters(), lr=lr_glad, betas=(0.9, 0.999), eps=1e-08, capturable=capturable)
    raise ValueError("Optimizer not found! Supported optimizers: ['adam']")


def glad(Sb: Tensor, model, lambda_init: float = 1, L: int = 15, INIT_DIAG: int = 0,
         USE_CUDA: bool = True, exact_sqrt: bool = False, group=None, total_graphs=None) -> Tensor:
    """Sb [B,D,D] (or [D,D]) covariance -> theta_pred [B,D,D].  Same arguments as the
    reference; two extras: `exact_sqrt` replaces the reference's 10-step Newton-Schulz square
    root by the exact one, `group` is a torch.distributed process group when the batch is
    sharded by graph over GPUs (`total_graphs`: the graph count over all ranks, if known)."""
    if Sb.dim() == 2:
        Sb = Sb.reshape(1, Sb.shape[0], Sb.shape[1])
    return ops.GladFunction.apply(Sb, model.packed(), int(L), int(INIT_DIAG), int(model.H),
                                  float(lambda_init), bool(exact_sqrt), group, total_graphs)

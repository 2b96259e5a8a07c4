"""GladParams: the learnable hyper-parameter networks of GLAD (reference:
uglad/glad/glad_params.py:5-91).  Same attribute names, same state_dict keys, same
initialisation order -- a reference checkpoint loads unchanged -- but eta_forward /
lambda_forward run the fused CUDA kernels."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from .. import ops


class GladParams(nn.Module):
    def __init__(self, theta_init_offset: float, nF: int, H: int, USE_CUDA: bool = True,
                 device=None) -> None:
        super().__init__()
        if nF != 3:
            raise ValueError("uglad_b200 kernels implement nF=3 (theta_k1, S, theta_prev), "
                             "the only value glad() uses (glad.py:144)")
        if not 1 <= H <= 8:
            raise ValueError("H must be in [1, 8]")
        self.nF, self.H = nF, H
        # construction order == reference order, so torch.manual_seed gives the same weights
        self.theta_init_offset = nn.Parameter(torch.tensor([float(theta_init_offset)]))
        self.rho_l1 = nn.Sequential(nn.Linear(nF, H), nn.Tanh(), nn.Linear(H, H), nn.Tanh(),
                                    nn.Linear(H, 1), nn.Sigmoid())
        self.lambda_f = nn.Sequential(nn.Linear(2, H), nn.Tanh(), nn.Linear(H, 1), nn.Sigmoid())
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is not None:
            self.to(device)

    def packed(self) -> Tensor:
        """Parameters flattened in state_dict order: the layout of uglad_param_count()."""
        return torch.cat([p.reshape(-1) for p in self.parameters()])

    def eta_forward(self, X: Tensor, S: Tensor, k: int = 0, F3: Tensor = None) -> Tensor:
        """Entrywise soft threshold by the rho_l1 network (glad_params.py:56-77).  Forward
        only: inside glad() the same kernel runs with its hand-written backward."""
        if F3 is None:
            raise ValueError("uglad_b200 implements the three-feature form used by glad()")
        Z, _ = ops.z_update(X, S, F3, self.packed().detach(), self.H)
        return Z

    def lambda_forward(self, normF, prev_lambda, k: int = 0) -> Tensor:
        """glad_params.py:79-91 (tiny 2-H-1 MLP on two scalars; inputs are detached)."""
        dev = self.theta_init_offset.device
        x = torch.tensor([float(normF), float(prev_lambda)], device=dev)
        return self.lambda_f(x)

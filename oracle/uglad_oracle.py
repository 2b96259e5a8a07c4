"""CPU oracle for the uGLAD hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / reference legs of
``bench.py`` may import this module.  The product path (``uglad_b200``) never does and
fails loudly when its CUDA library is missing.

Two independent restatements of the reference algorithm live here:

* ``matmul`` formulation  (``glad_unrolled`` / ``glasso_loss`` / ``train``): the
  reference's own arithmetic -- Newton-Schulz matrix square roots by repeated matmul,
  torch autograd for everything else -- written batched instead of per-matrix.  This is
  what the parity tests compare the CUDA path against and what ``bench.py --impl
  reference`` times.  Parity pinned: ``tests/golden/make_golden.py`` runs the real
  reference (imported from /root/reference) on seeded inputs and
  ``tests/test_oracle_golden.py`` checks this module against those vectors.

* ``spectral`` formulation (``spectral_*``): float64 numpy eigendecomposition with the
  Newton-Schulz iteration collapsed to a scalar recurrence per eigenvalue and its
  backward collapsed to an F-matrix.  It is the algorithm the CUDA kernels implement, so
  agreement between the two formulations is the proof that the kernel math is the
  reference's math.

Reference map (file:line under /root/reference/uglad):
  glad/glad.py:74-150          unrolled alternating minimisation  -> glad_unrolled
  glad/glad_params.py:33-54    rho_l1 / lambda_f MLPs             -> rho_net / lambda_net
  glad/glad_params.py:56-77    entrywise soft threshold           -> eta_threshold
  glad/glad_params.py:79-91    lambda update (inputs detached)    -> lambda_net call sites
  glad/torch_sqrtm.py:12-28    Newton-Schulz sqrt forward          -> _NSSqrt.forward
  glad/torch_sqrtm.py:31-45    Newton-Schulz sqrt backward         -> _NSSqrt.backward
  main.py:289-335              glasso loss (+ log-cosh struct loss)-> glasso_loss
  main.py:338-425              direct-mode Adam loop               -> train
  main.py:673-716              consensus over imputations          -> consensus_min
  utils/prepare_data.py:328-356 covariance + eigenvalue repair     -> covariance
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

NS_ITERS = 10  # torch_sqrtm.py:13 and :33 (itr_TH)


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------
PARAM_KEYS = (
    "theta_init_offset",
    "rho_l1.0.weight", "rho_l1.0.bias", "rho_l1.2.weight", "rho_l1.2.bias",
    "rho_l1.4.weight", "rho_l1.4.bias",
    "lambda_f.0.weight", "lambda_f.0.bias", "lambda_f.2.weight", "lambda_f.2.bias",
)


def init_params(seed: int, theta_init_offset: float = 1.0, nF: int = 3, H: int = 3,
                dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Same parameter names/shapes as GladParams.state_dict() (glad_params.py:10-31),
    drawn with nn.Linear's default init in the same construction order so that
    torch.manual_seed(seed) reproduces the reference's initial weights."""
    torch.manual_seed(seed)
    l1, lH1, l2 = torch.nn.Linear(nF, H), torch.nn.Linear(H, H), torch.nn.Linear(H, 1)
    f1, f2 = torch.nn.Linear(2, H), torch.nn.Linear(H, 1)
    vals = [torch.tensor([theta_init_offset]),
            l1.weight, l1.bias, lH1.weight, lH1.bias, l2.weight, l2.bias,
            f1.weight, f1.bias, f2.weight, f2.bias]
    return {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in zip(PARAM_KEYS, vals)}


def rho_net(P, feats: torch.Tensor) -> torch.Tensor:
    """glad_params.py:33-44: Linear-tanh-Linear-tanh-Linear-sigmoid on [..., nF]."""
    h = torch.tanh(feats @ P["rho_l1.0.weight"].T + P["rho_l1.0.bias"])
    h = torch.tanh(h @ P["rho_l1.2.weight"].T + P["rho_l1.2.bias"])
    return torch.sigmoid(h @ P["rho_l1.4.weight"].T + P["rho_l1.4.bias"])


def lambda_net(P, normF: float, prev_lambda: float) -> torch.Tensor:
    """glad_params.py:46-54 + :79-91.  The reference rebuilds the feature vector with
    torch.Tensor([...]), which detaches both inputs: gradient reaches lambda_f's weights
    only."""
    dt = P["lambda_f.0.weight"].dtype
    x = torch.tensor([float(normF), float(prev_lambda)], dtype=dt)
    h = torch.tanh(P["lambda_f.0.weight"] @ x + P["lambda_f.0.bias"])
    return torch.sigmoid(P["lambda_f.2.weight"] @ h + P["lambda_f.2.bias"])  # shape [1]


def eta_threshold(P, X, S, F3):
    """glad_params.py:56-77: rho = rho_l1([X, S, F3]) per entry; sign(X)*max(0,|X|-rho)."""
    feats = torch.stack((X, S, F3), dim=-1)
    rho = rho_net(P, feats).squeeze(-1)
    return torch.sign(X) * torch.clamp_min(torch.abs(X) - rho, 0.0)


# --------------------------------------------------------------------------------------
# Newton-Schulz square root, batched (torch_sqrtm.py)
# --------------------------------------------------------------------------------------
class _NSSqrt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A):  # A: [B,D,D]
        D = A.shape[-1]
        eye = torch.eye(D, dtype=A.dtype).expand_as(A)
        nrm = torch.linalg.matrix_norm(A).reshape(-1, 1, 1)  # Frobenius, per matrix
        Y, Z = A / nrm, eye.clone()
        for _ in range(NS_ITERS):
            T = 0.5 * (3.0 * eye - Z @ Y)
            Y, Z = Y @ T, T @ Z
        R = Y * torch.sqrt(nrm)
        ctx.save_for_backward(R)
        return R

    @staticmethod
    def backward(ctx, G):
        (R,) = ctx.saved_tensors
        D = R.shape[-1]
        eye = torch.eye(D, dtype=R.dtype).expand_as(R)
        nrm = torch.linalg.matrix_norm(R).reshape(-1, 1, 1)
        A, Q = R / nrm, G / nrm
        for _ in range(NS_ITERS):
            At = A.transpose(-1, -2)
            Q = 0.5 * (Q @ (3.0 * eye - A @ A) - At @ (At @ Q - Q @ A))
            A = 0.5 * A @ (3.0 * eye - A @ A)
        return 0.5 * Q


def ns_sqrt(A: torch.Tensor) -> torch.Tensor:
    return _NSSqrt.apply(A)


# --------------------------------------------------------------------------------------
# unrolled GLAD + loss  (matmul formulation)
# --------------------------------------------------------------------------------------
def _per_matrix(fn, A: torch.Tensor) -> torch.Tensor:
    """torch 2.11's BATCHED CPU LU (inverse / logdet) is broken in this image when it runs on more than
    one thread and D > 128 (oneMKL "Parameter 6 was incorrect on entry to SLASWP", corrupt pivots or a
    hang; profiles/r02_torch_cpu_batched_lu_bug.txt).  The unbatched call is sound, so larger matrices
    go through it one by one -- the same arithmetic per matrix."""
    if A.shape[0] > 1 and A.shape[-1] > 128 and torch.get_num_threads() > 1:
        return torch.stack([fn(a) for a in A])
    return fn(A)



def glad_unrolled(S: torch.Tensor, P, L: int = 15, init_diag: int = 0,
                  lambda_init: float = 1.0, trace: Optional[dict] = None) -> torch.Tensor:
    """glad.py:74-150.  S: [B,D,D].  Returns theta_pred [B,D,D] (autograd-connected)."""
    if S.dim() == 2:
        S = S.unsqueeze(0)
    D = S.shape[-1]
    eye = torch.eye(D, dtype=S.dtype).expand_as(S)
    t0 = P["theta_init_offset"]
    if init_diag == 1:
        theta = torch.diag_embed(1.0 / (torch.diagonal(S, dim1=-2, dim2=-1) + t0))
    else:
        theta = _per_matrix(torch.linalg.inv, S + t0 * eye)
    lam = lambda_net(P, lambda_init, 0.0)
    lams, norms = [lam.detach().clone()], []
    for _ in range(L):
        b = (1.0 / lam) * S - theta
        sq = ns_sqrt(b.transpose(-1, -2) @ b + (4.0 / lam) * eye)
        x = 0.5 * (sq - b)
        new_theta = eta_threshold(P, x, S, theta)
        normF = torch.mean(torch.sum((new_theta - x) ** 2, dim=(1, 2))).item()
        lam = lambda_net(P, normF, lam.item())
        theta = new_theta
        lams.append(lam.detach().clone())
        norms.append(normF)
    if trace is not None:
        trace["lambda"] = torch.cat(lams[:-1]).numpy()  # lambda_k used by layer k
        trace["normF"] = np.asarray(norms)
    return theta


def glasso_loss(theta: torch.Tensor, S: torch.Tensor,
                struct_theta: Optional[torch.Tensor] = None) -> torch.Tensor:
    """main.py:289-335: sum_b(-logdet(theta_b) + tr(S_b theta_b)) / B  (+ log-cosh prior)."""
    B, D, _ = S.shape
    t1 = -_per_matrix(torch.logdet, theta)
    t2 = torch.einsum("bij,bji->b", S, theta)
    loss = torch.sum(t1 + t2) / B
    if struct_theta is not None:
        mask = (1 - struct_theta) - torch.eye(D, dtype=theta.dtype).expand(B, -1, -1)
        loss = loss + torch.sum(torch.log(torch.cosh(theta * mask))) / B
    return loss


def forward_loss(S, P, L=15, init_diag=0, loss_S=None, struct_theta=None):
    """main.py:252-286 forward_uGLAD."""
    theta = glad_unrolled(S, P, L=L, init_diag=init_diag)
    return theta, glasso_loss(theta, S if loss_S is None else loss_S, struct_theta)


def train(S: torch.Tensor, P, epochs: int, lr: float = 0.002, L: int = 15, init_diag: int = 0,
          loss_S=None, struct_theta=None):
    """main.py:389-414 (direct), :616-630 (missing), :766-778 (multitask): Adam
    (glad.py:28-35: betas .9/.999, eps 1e-8) on the glasso loss.  Returns the theta of the
    last forward (the reference reports the pre-step theta of the final epoch) and losses."""
    opt = torch.optim.Adam(list(P.values()), lr=lr, betas=(0.9, 0.999), eps=1e-8)
    losses, theta = [], None
    for _ in range(epochs):
        opt.zero_grad()
        theta, loss = forward_loss(S, P, L, init_diag, loss_S, struct_theta)
        if torch.isnan(loss):
            break
        loss.backward()
        opt.step()
        losses.append(float(loss))
    return theta.detach(), np.asarray(losses)


def consensus_min(theta_K: torch.Tensor) -> torch.Tensor:
    """main.py:673-716 with type='min': sign by majority (ties -> +1) times min |.|."""
    val = torch.min(torch.abs(theta_K), 0)[0]
    sgn = torch.sum(torch.sign(theta_K), 0)
    sgn = torch.where(sgn >= 0, torch.ones_like(sgn), -torch.ones_like(sgn))
    return (sgn * val).reshape(1, *theta_K.shape[1:])


def covariance(Xb, offset: float = 0.1) -> np.ndarray:
    """prepare_data.py:328-356: per matrix, biased covariance of the centred samples in
    float64 (sklearn empirical_covariance == np.cov(X.T, bias=1)); when the smallest
    eigenvalue is <= 1e-6 shift the diagonal so that it becomes `offset`."""
    out = []
    for X in Xb:
        X = np.asarray(X, dtype=np.float64)
        Xc = X - X.mean(axis=0)
        S = Xc.T @ Xc / X.shape[0]
        ev = np.linalg.eigvals(S).real
        if ev.min() <= 1e-6:
            S = S + np.eye(S.shape[-1]) * (offset - ev.min())
        out.append(S)
    return np.array(out)


# --------------------------------------------------------------------------------------
# spectral formulation (float64 numpy) -- the algorithm the CUDA kernels implement
# --------------------------------------------------------------------------------------
def ns_scalar_forward(mu: np.ndarray) -> np.ndarray:
    """Newton-Schulz forward on the eigenvalues mu (>0) of one matrix: every iterate is a
    polynomial in the input, so Y_t = V diag(y_t) V^T with the scalar recurrence below."""
    nrm = math.sqrt(float(np.sum(mu * mu)))
    y, z = mu / nrm, np.ones_like(mu)
    for _ in range(NS_ITERS):
        t = 0.5 * (3.0 - z * y)
        y, z = y * t, t * z
    return y * math.sqrt(nrm)


def ns_scalar_backward_factor(s: np.ndarray) -> np.ndarray:
    """Newton-Schulz backward in the eigenbasis of the saved root (eigenvalues s): entry
    (i,j) of Q is scaled each step by 0.5*(3 - a_i^2 - a_j^2 + a_i a_j) while a follows
    a <- 0.5 a (3 - a^2).  Returns C with grad_in = V (C * (V^T grad_out V)) V^T."""
    nrm = math.sqrt(float(np.sum(s * s)))
    a = s / nrm
    c = np.ones((s.size, s.size))
    for _ in range(NS_ITERS):
        c *= 0.5 * (3.0 - a[:, None] ** 2 - a[None, :] ** 2 + a[:, None] * a[None, :])
        a = 0.5 * a * (3.0 - a * a)
    return 0.5 * c / nrm


def _mlp_rho_np(W, x, s, f):
    """rho MLP forward on float64 arrays; also returns what the backward needs."""
    feats = np.stack((x, s, f), -1)
    h1 = np.tanh(feats @ W["rho_l1.0.weight"].T + W["rho_l1.0.bias"])
    h2 = np.tanh(h1 @ W["rho_l1.2.weight"].T + W["rho_l1.2.bias"])
    o = 1.0 / (1.0 + np.exp(-(h2 @ W["rho_l1.4.weight"].T + W["rho_l1.4.bias"])))
    return feats, h1, h2, o[..., 0]


def _mlp_lambda_np(W, nf, pl):
    x = np.array([nf, pl])
    h = np.tanh(W["lambda_f.0.weight"] @ x + W["lambda_f.0.bias"])
    o = 1.0 / (1.0 + np.exp(-(W["lambda_f.2.weight"] @ h + W["lambda_f.2.bias"])))
    return x, h, float(o[0])


def spectral_forward_backward(S, P, L=15, init_diag=0, lambda_init=1.0, exact_sqrt=False,
                              loss_S=None):
    """Forward + backward of glasso_loss(glad_unrolled(S)) in float64 via eigendecomposition.
    S: [B,D,D] array.  Returns dict(theta, loss, grads{name: array}, lambda, normF).
    exact_sqrt=True replaces the 10-step Newton-Schulz emulation by the true sqrt /
    Lyapunov solve (what the reference would compute with infinitely many iterations)."""
    S = np.asarray(S, dtype=np.float64)
    W = {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)).astype(np.float64)
         for k, v in P.items()}
    B, D, _ = S.shape
    LS = S if loss_S is None else np.asarray(loss_S, dtype=np.float64)
    t0 = float(W["theta_init_offset"][0])
    sS, VS = np.linalg.eigh(S)
    if init_diag == 1:
        dS = np.einsum("bii->bi", S)
        theta = np.stack([np.diag(1.0 / (dS[b] + t0)) for b in range(B)])
    else:
        theta = np.matmul(VS * (1.0 / (sS + t0))[:, None, :], VS.transpose(0, 2, 1))
    theta0 = theta
    lam_in, lam_h, lam = _mlp_lambda_np(W, lambda_init, 0.0)
    saved, lam_feats = [], [(lam_in, lam_h, lam)]
    for _ in range(L):
        b = S / lam - theta
        beta, V = np.linalg.eigh(b)
        mu = beta * beta + 4.0 / lam
        s = np.sqrt(mu) if exact_sqrt else np.stack([ns_scalar_forward(m) for m in mu])
        x = np.matmul(V * (0.5 * (s - beta))[:, None, :], V.transpose(0, 2, 1))
        feats, h1, h2, rho = _mlp_rho_np(W, x, S, theta)
        z = np.sign(x) * np.maximum(np.abs(x) - rho, 0.0)
        normF = float(np.mean(np.sum((z - x) ** 2, axis=(1, 2))))
        saved.append((lam, beta, V, s, x, feats, h1, h2, rho))
        lam_in, lam_h, lam = _mlp_lambda_np(W, normF, lam)
        lam_feats.append((lam_in, lam_h, lam))
        theta = z
    # loss
    ev, Vt = np.linalg.eigh(0.5 * (theta + theta.transpose(0, 2, 1)))
    sign = np.prod(np.sign(ev), axis=1)
    logdet = np.where(sign > 0, np.sum(np.log(np.abs(ev)), axis=1), np.nan)
    loss = float(np.sum(-logdet + np.einsum("bij,bji->b", LS, theta)) / B)
    # backward
    g = {k: np.zeros_like(v) for k, v in W.items()}
    G = (-np.matmul(Vt * (1.0 / ev)[:, None, :], Vt.transpose(0, 2, 1)) + LS.transpose(0, 2, 1)) / B
    g_lams = np.zeros(L)
    for k in reversed(range(L)):
        lam_k, beta, V, s, x, feats, h1, h2, rho = saved[k]
        act = (np.abs(x) - rho) > 0
        g_rho = np.where(act, -np.sign(x) * G, 0.0)
        g_x = np.where(act, G, 0.0)
        # rho MLP backward
        d3 = g_rho * rho * (1 - rho)
        g["rho_l1.4.weight"] += np.einsum("bij,bijh->h", d3, h2)[None, :]
        g["rho_l1.4.bias"] += d3.sum()
        d2 = d3[..., None] * W["rho_l1.4.weight"][0] * (1 - h2 * h2)
        g["rho_l1.2.weight"] += np.einsum("bijo,bijh->oh", d2, h1)
        g["rho_l1.2.bias"] += d2.sum(axis=(0, 1, 2))
        d1 = (d2 @ W["rho_l1.2.weight"]) * (1 - h1 * h1)
        g["rho_l1.0.weight"] += np.einsum("bijo,bijf->of", d1, feats)
        g["rho_l1.0.bias"] += d1.sum(axis=(0, 1, 2))
        gf = d1 @ W["rho_l1.0.weight"]  # grads wrt (x, S, theta_prev)
        g_x = g_x + gf[..., 0]
        g_prev = gf[..., 2]
        # spectral backward of x = V f(beta) V^T; only the symmetric part of the incoming
        # gradient can reach a parameter, so it is symmetrised here.
        g_x = 0.5 * (g_x + g_x.transpose(0, 2, 1))
        Gt = np.matmul(np.matmul(V.transpose(0, 2, 1), g_x), V)
        g_b = np.empty_like(Gt)
        for bb in range(B):
            if exact_sqrt:
                C = 0.5 / (s[bb][:, None] + s[bb][None, :])
            else:
                C = 0.5 * ns_scalar_backward_factor(s[bb])
            # grad wrt (b^T b + 4/lam I) in the eigenbasis is H = C * Gt ; d(b^T b) -> b(H+H^T)
            H = C * Gt[bb]
            g_lams[k] += -4.0 / lam_k ** 2 * np.trace(H)
            g_b[bb] = 0.5 * (beta[bb][:, None] + beta[bb][None, :]) * (2.0 * H) - 0.5 * Gt[bb]
        g_b = np.matmul(np.matmul(V, g_b), V.transpose(0, 2, 1))
        g_lams[k] += -np.sum(S * g_b) / lam_k ** 2
        G = g_prev - g_b
    # theta_init
    if init_diag == 1:
        dS = np.einsum("bii->bi", S)
        g["theta_init_offset"] += -np.sum(np.einsum("bii->bi", G) / (dS + t0) ** 2)
    else:
        Gs = 0.5 * (G + G.transpose(0, 2, 1))
        Gt = np.einsum("bki,bki->bi", VS, np.matmul(Gs, VS))
        g["theta_init_offset"] += -np.sum(Gt / (sS + t0) ** 2)
    # lambda MLP backward (inputs are constants)
    for k in range(L):
        xin, h, o = lam_feats[k]
        d2 = g_lams[k] * o * (1 - o)
        g["lambda_f.2.weight"] += d2 * h[None, :]
        g["lambda_f.2.bias"] += d2
        d1 = d2 * W["lambda_f.2.weight"][0] * (1 - h * h)
        g["lambda_f.0.weight"] += np.outer(d1, xin)
        g["lambda_f.0.bias"] += d1
    return dict(theta=theta, theta0=theta0, loss=loss, grads=g,
                lam=np.array([sv[0] for sv in saved]), g_lam=g_lams)


# --------------------------------------------------------------------------------------
# mode drivers (main.py:34-226, :338-789) restated on top of the matmul formulation; pinned by
# tests/golden/modes.npz (the real reference's uGLAD_GL / uGLAD_multitask fits)
# --------------------------------------------------------------------------------------
def clean_table_minmax(X) -> np.ndarray:
    """prepare_data.py:361-516 with fit()'s arguments (NORM='min_max', no variance / condition
    pruning): drop all-zero rows, fill NaNs with column means, drop constant columns, min-max
    normalise, drop duplicated columns (first occurrence kept)."""
    X = np.asarray(X, dtype=np.float64)
    X = X[~np.all(X == 0, axis=1)]
    col_mean = np.nanmean(X, axis=0)
    X = np.where(np.isnan(X), col_mean, X)
    keep = [j for j in range(X.shape[1]) if np.unique(X[:, j]).size > 1]
    X = X[:, keep]
    X = (X - X.min(0)) / (X.max(0) - X.min(0))
    seen, cols = set(), []
    for j in range(X.shape[1]):
        key = X[:, j].tobytes()
        if key not in seen:
            seen.add(key)
            cols.append(j)
    return X[:, cols]


def kfold_blocks(M: int, K: int):
    """sklearn KFold(n_splits=K) without shuffling: (train, test) index arrays per fold."""
    sizes = np.full(K, M // K)
    sizes[: M % K] += 1
    out, start = [], 0
    for sz in sizes:
        test = np.arange(start, start + sz)
        out.append((np.concatenate([np.arange(0, start), np.arange(start + sz, M)]), test))
        start += sz
    return out


def _f32(S):
    return torch.tensor(np.asarray(S), dtype=torch.float32)


def init_params_running(theta_init_offset: float = 1.0, nF: int = 3, H: int = 3, dtype=torch.float32):
    """init_params without re-seeding: the next GladParams drawn from torch's global generator
    (the CV mode calls init_uGLAD once per fold, main.py:483)."""
    l1, lH1, l2 = torch.nn.Linear(nF, H), torch.nn.Linear(H, H), torch.nn.Linear(H, 1)
    f1, f2 = torch.nn.Linear(2, H), torch.nn.Linear(H, 1)
    vals = [torch.tensor([theta_init_offset]), l1.weight, l1.bias, lH1.weight, lH1.bias, l2.weight, l2.bias,
            f1.weight, f1.bias, f2.weight, f2.bias]
    return {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in zip(PARAM_KEYS, vals)}


def fit_direct(X, seed, epochs, lr, L=15, true_theta=None):
    """uGLAD_GL.fit(mode='direct') (main.py:338-425): the structure prior is the true theta."""
    X = clean_table_minmax(X)
    S = _f32(covariance(X[None]))
    P = init_params(seed)
    st = None if true_theta is None else _f32(np.asarray(true_theta)[None])
    theta, _ = train(S, P, epochs, lr, L, 0, struct_theta=st)
    return theta[0].numpy()


def fit_cv(X, seed, epochs, lr, k_fold, L=15):
    """uGLAD_GL.fit(mode='cv') (main.py:428-550): per fold keep the (post-step) parameters of the
    epoch with the best held-out loss; the best fold's model is run on the full covariance."""
    X = clean_table_minmax(X)
    S = _f32(covariance(X[None]))
    torch.manual_seed(seed)
    best = (np.inf, None)
    for train_idx, test_idx in kfold_blocks(X.shape[0], k_fold):
        S_tr, S_te = _f32(covariance(X[train_idx][None])), _f32(covariance(X[test_idx][None]))
        P = init_params_running()
        opt = torch.optim.Adam(list(P.values()), lr=lr, betas=(0.9, 0.999), eps=1e-8)
        fold_best = (np.inf, None)
        for _ in range(epochs):
            opt.zero_grad()
            _, ltr = forward_loss(S_tr, P, L, 0)
            with torch.no_grad():
                _, lte = forward_loss(S_te, P, L, 0)
            ltr.backward()
            opt.step()
            if float(lte) < fold_best[0]:
                fold_best = (float(lte), {k: v.detach().clone() for k, v in P.items()})
        if fold_best[0] < best[0]:
            best = fold_best
    with torch.no_grad():
        theta, _ = forward_loss(S, best[1], L, 0)
    return theta[0].numpy()


def fit_missing(X, seed, epochs, lr, k_fold, L=15):
    """uGLAD_GL.fit(mode='missing') (main.py:553-644): K row-subsampled covariances trained jointly
    against the full-data covariance (the loss is divided by ITS batch size, 1), then the consensus."""
    X = clean_table_minmax(X)   # NaNs are mean-imputed by process_table already
    S = _f32(covariance(X[None]))
    S_K = _f32(covariance([X[tr] for tr, _ in kfold_blocks(X.shape[0], k_fold)]))
    P = init_params(seed)
    opt = torch.optim.Adam(list(P.values()), lr=lr, betas=(0.9, 0.999), eps=1e-8)
    theta = None
    for _ in range(epochs):
        opt.zero_grad()
        theta = glad_unrolled(S_K, P, L=L)
        loss = torch.sum(-_per_matrix(torch.logdet, theta) + torch.einsum("ij,bji->b", S[0], theta)) / S.shape[0]
        loss.backward()
        opt.step()
    return consensus_min(theta.detach())[0].numpy()


def fit_multitask(Xs, seed, epochs, lr, L=15):
    """uGLAD_multitask.fit (main.py:155-226, :719-789)."""
    Xs = [clean_table_minmax(X) for X in Xs]
    S = _f32(covariance(Xs))
    P = init_params(seed)
    theta, _ = train(S, P, epochs, lr, L, 0)
    return theta.numpy()

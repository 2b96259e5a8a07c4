"""The oracle port against the REAL reference run here (oracle/_ref: the reference's own package tree, vendored by
oracle/build_ref.py and imported through oracle/ref_loader.py): forward theta, loss and all eleven gradients on fresh
seeds -- next to the committed goldens this pins the oracle to the reference's behaviour on inputs the goldens do not
hold (multitask batch, consensus-style loss covariance, structure prior, INIT_DIAG=1).  Skipped when oracle/_ref has not
been built (it is git-ignored; __graft_entry__.build() creates it wherever /root/reference exists)."""
import contextlib
import io
import warnings

import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import uglad_oracle as O

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not built")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _reference_run(ref, S, seed, L, init_diag, loss_S, struct):
    torch.manual_seed(seed)
    model, _ = ref.main.init_uGLAD(lr=0.002, theta_init_offset=1.0, nF=3, H=3)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        theta, loss = ref.main.forward_uGLAD(S, model, L=L, INIT_DIAG=init_diag, loss_Sb=loss_S, struct_theta=struct)
    loss.backward()
    return theta.detach().numpy(), float(loss), {k: p.grad.numpy().copy() for k, p in model.named_parameters()}


@pytest.mark.parametrize("B,D,L,init_diag,consensus,prior", [(1, 12, 15, 0, False, False), (3, 20, 15, 0, False, False),
                                                             (4, 16, 6, 0, True, False), (1, 10, 15, 1, False, False),
                                                             (2, 14, 8, 0, False, True)])
def test_oracle_port_equals_the_vendored_reference(B, D, L, init_diag, consensus, prior):
    ref = ref_loader.load()
    rng = np.random.default_rng(100 * B + D)
    X = rng.random((B, 4 * D, D))
    S = torch.tensor(O.covariance(X), dtype=torch.float32)
    loss_S = torch.tensor(O.covariance(X[:1]), dtype=torch.float32) if consensus else None
    struct = None
    if prior:
        A = (rng.random((B, D, D)) < 0.2).astype(np.float32)
        struct = torch.tensor(np.maximum(A, A.transpose(0, 2, 1)))
    seed = 7 + D
    th_r, loss_r, g_r = _reference_run(ref, S, seed, L, init_diag, loss_S, struct)
    P = O.init_params(seed)
    th_o, loss_o = O.forward_loss(S, P, L, init_diag, loss_S=loss_S, struct_theta=struct)
    loss_o.backward()
    assert rel(th_o.detach().numpy(), th_r) < 1e-5
    assert np.array_equal(th_o.detach().numpy() != 0, th_r != 0)
    assert abs(float(loss_o) - loss_r) < 1e-5 * max(1.0, abs(loss_r))
    for k in O.PARAM_KEYS:
        assert rel(P[k].grad.numpy(), g_r[k]) < 1e-4, k


def test_vendored_tree_is_the_reference_unmodified():
    """MANIFEST.json (written by oracle/build_ref.py) lists the copied files with their hashes."""
    import hashlib
    import json
    import os
    man = json.load(open(os.path.join(ref_loader.REF_DIR, "MANIFEST.json")))
    files = man.get("files", man)
    assert files
    for relpath, digest in (files.items() if isinstance(files, dict) else []):
        p = os.path.join(ref_loader.REF_DIR, relpath)
        if os.path.isfile(p) and isinstance(digest, str) and len(digest) in (40, 64):
            h = hashlib.sha256(open(p, "rb").read()).hexdigest() if len(digest) == 64 else hashlib.sha1(open(p, "rb").read()).hexdigest()
            assert h == digest, relpath

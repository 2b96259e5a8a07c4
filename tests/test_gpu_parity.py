"""Parity of the CUDA path (through the C-ABI library) against the oracle and the golden
vectors produced by the real reference.  Run on the B200 box:  pytest -m gpu

Tolerances (BASELINE.json north_star): theta within 1e-4 relative Frobenius error of the
reference's torch result; recovered edge set identical except for entries within 1e-5 of the
threshold.  Gradients / losses are held to 1e-3 / 1e-4 relative."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import uglad_oracle as O  # the checker, never the thing under test

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
THETA_TOL = 1e-4


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def load_model(g, tag):
    from uglad_b200.glad.glad_params import GladParams
    model = GladParams(1.0, 3, 3)
    model.load_state_dict({k: torch.tensor(g[f"{tag}/{k}"]) for k in O.PARAM_KEYS})
    return model


def edge_sets_match(theta, ref, margin=1e-5):
    """Supports must agree wherever the reference entry is not within `margin` of zero."""
    clear = np.abs(ref) > margin
    same = (theta != 0) == (ref != 0)
    return bool(np.all(same | ~clear)) and bool(np.all(np.abs(theta[(ref == 0)]) <= margin))


def test_library_is_the_cuda_one():
    from uglad_b200 import _lib
    lib = _lib.load()
    assert lib.uglad_abi_version() == 1
    assert os.path.basename(_lib.LIB_PATH) == "libuglad_b200.so"


@pytest.mark.parametrize("D,B", [(1, 2), (2, 3), (3, 2), (10, 4), (33, 3), (100, 5), (129, 2), (167, 2), (200, 2), (232, 1)])
def test_eigh_indefinite(dev, D, B):
    from uglad_b200 import ops
    rng = np.random.default_rng(D)
    A = rng.standard_normal((B, D, D))
    A = (A + A.transpose(0, 2, 1)) / 2
    w, Vt, info = ops.eigh(torch.tensor(A, dtype=torch.float32, device=dev), indefinite=True)
    w, Vt = w.cpu().numpy().astype(np.float64), Vt.cpu().numpy().astype(np.float64)
    R = np.einsum("bki,bk,bkj->bij", Vt, w, Vt)
    assert rel(R, A) < 3e-5
    assert np.abs(np.einsum("bki,bli->bkl", Vt, Vt) - np.eye(D)).max() < 2e-5
    assert np.abs(np.sort(w, 1) - np.linalg.eigvalsh(A)).max() < 3e-5 * np.abs(A).sum(-1).max()


@pytest.mark.parametrize("D", [5, 64, 150])
def test_eigh_positive_definite_and_degenerate(dev, D):
    from uglad_b200 import ops
    rng = np.random.default_rng(D)
    Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    ev = np.concatenate([np.full(D // 2, 2.0), np.linspace(0.05, 5.0, D - D // 2)])  # repeated eigenvalue
    A = (Q * ev) @ Q.T
    w, Vt, _ = ops.eigh(torch.tensor(A[None], dtype=torch.float32, device=dev), indefinite=False)
    w, Vt = w.cpu().numpy().astype(np.float64), Vt.cpu().numpy().astype(np.float64)
    assert np.abs(np.sort(w[0]) - np.sort(ev)).max() < 2e-5
    assert rel(np.einsum("bki,bk,bkj->bij", Vt, w, Vt)[0], A) < 2e-5


def test_eigh_diagonal_and_zero_offdiag(dev):
    from uglad_b200 import ops
    d = np.array([3.0, -1.0, 0.5, 7.0, 2.0, 2.0])
    w, Vt, _ = ops.eigh(torch.tensor(np.diag(d)[None], dtype=torch.float32, device=dev), indefinite=True)
    assert np.abs(np.sort(w.cpu().numpy()[0]) - np.sort(d)).max() < 1e-5


@pytest.mark.parametrize("B,M,D", [(1, 3, 2), (2, 50, 7), (3, 500, 20), (1, 1000, 100), (2, 257, 65)])
def test_covariance(dev, B, M, D):
    from uglad_b200 import ops
    rng = np.random.default_rng(M)
    X = rng.random((B, M, D))
    Xd = torch.tensor(X, dtype=torch.float32, device=dev)
    S = ops.covariance(Xd).cpu().numpy()   # tcgen05 3xTF32 contraction on the centred, feature-major samples
    assert rel(S, O.covariance(X, offset=0.1)) < 5e-6
    assert np.array_equal(S, S.transpose(0, 2, 1))   # exactly symmetric
    ops.tune("use_tc", 0)                   # FP32 SIMT contraction
    try:
        S_simt = ops.covariance(Xd).cpu().numpy()
    finally:
        ops.tune("use_tc", 1)
    assert rel(S_simt, O.covariance(X, offset=0.1)) < 5e-6
    assert rel(S, S_simt) < 5e-6


def test_rank_deficient_covariance_is_repaired(dev):
    from uglad_b200.utils import prepare_data
    rng = np.random.default_rng(3)
    X = rng.standard_normal((1, 5, 12))  # fewer samples than features
    S = prepare_data.get_covariance(X, offset=0.1).cpu().numpy()
    assert rel(S, O.covariance(X, offset=0.1)) < 1e-5
    assert abs(np.linalg.eigvalsh(S[0].astype(np.float64)).min() - 0.1) < 1e-5


def test_ragged_sample_counts(dev):
    from uglad_b200.utils import prepare_data
    rng = np.random.default_rng(4)
    Xs = [rng.random((m, 9)) for m in (40, 55, 40)]
    S = prepare_data.get_covariance(Xs).cpu().numpy()
    for i, X in enumerate(Xs):
        assert rel(S[i], O.covariance([X])[0]) < 5e-6


def test_z_update_matches_oracle(dev):
    from uglad_b200 import ops
    P = O.init_params(5)
    rng = np.random.default_rng(5)
    X, S, T = (torch.tensor(rng.standard_normal((3, 17, 17)), dtype=torch.float32) for _ in range(3))
    flat = torch.cat([P[k].detach().reshape(-1) for k in O.PARAM_KEYS]).to(dev)
    Z, normf = ops.z_update(X.to(dev), S.to(dev), T.to(dev), flat, 3)
    Zo = O.eta_threshold(P, X, S, T).detach()
    assert rel(Z.cpu().numpy(), Zo.numpy()) < 1e-6
    assert abs(normf.item() - float(((Zo - X) ** 2).sum())) < 1e-4 * float(((Zo - X) ** 2).sum())


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_forward_backward_against_reference_golden(dev, path):
    from uglad_b200 import main as ug, ops
    ops.reset_warm_start()
    g = np.load(path)
    L, idg = int(g["L"]), int(g["init_diag"])
    S = torch.tensor(g["S"], device=dev)
    model = load_model(g, "p0")
    theta, loss = ug.forward_uGLAD(S, model, L=L, INIT_DIAG=idg)
    loss.backward()
    th = theta.detach().cpu().numpy()
    assert rel(th, g["theta0"]) < THETA_TOL
    assert edge_sets_match(th, g["theta0"])
    assert abs(loss.item() - float(g["loss0"])) < 1e-4 * max(1.0, abs(float(g["loss0"])))
    for k, p in model.named_parameters():
        assert rel(p.grad.cpu().numpy(), g["g0/" + k]) < 1e-3, k


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_training_loop_against_reference_golden(dev, path):
    """Same seeds, same Adam loop (main.py:389-414): the trajectory of the real reference."""
    from uglad_b200 import main as ug, ops
    ops.reset_warm_start()
    g = np.load(path)
    L, idg, E = int(g["L"]), int(g["init_diag"]), int(g["epochs"])
    S = torch.tensor(g["S"], device=dev)
    model = load_model(g, "p0")
    opt = ug.glad.get_optimizers(model, lr_glad=float(g["lr"]))
    theta, losses = ug._fit_loop(S, model, opt, E, L, idg, False)
    losses = torch.stack(losses).cpu().numpy()
    assert np.abs(losses - g["losses"]).max() < 2e-4 * max(1.0, np.abs(g["losses"]).max())
    th = theta.detach().cpu().numpy()
    assert rel(th, g["thetaT"]) < THETA_TOL
    assert edge_sets_match(th, g["thetaT"])
    if "consensus" in g.files:
        c = ug.get_final_precision_from_batch(theta.detach(), type="min").cpu().numpy()
        assert rel(c, g["consensus"]) < THETA_TOL


def test_eigensolver_path_with_simt_products(dev):
    """The FP32 SIMT products behind the "use_tc" = 0 knob give the same answer as the tcgen05 ones."""
    from uglad_b200 import main as ug, ops
    g = np.load(os.path.join(ROOT, "tests", "golden", "d100_lowrho.npz"))
    S = torch.tensor(g["S"], device=dev)
    ops.tune("use_tc", 0)
    try:
        ops.reset_warm_start()
        model = load_model(g, "p0")
        theta, loss = ug.forward_uGLAD(S, model, L=int(g["L"]), INIT_DIAG=int(g["init_diag"]))
        loss.backward()
        assert rel(theta.detach().cpu().numpy(), g["theta0"]) < THETA_TOL
        for k, p in model.named_parameters():
            assert rel(p.grad.cpu().numpy(), g["g0/" + k]) < 1e-3, k
    finally:
        ops.tune("use_tc", 1)
        ops.reset_warm_start()


@pytest.mark.parametrize("name", ["d100_lowrho.npz", "d12_raw.npz", "d20_b3_multitask.npz"])
def test_eigensolver_path_with_presplit_products(dev, name):
    """The eigenvector products on pre-split (hi, lo) pairs behind "eig_raw" = 0 (the default forms
    hi/lo inside the tcgen05 kernel from plain operands; D = 12, 20, 100 cover rows that are and are
    not 16-byte multiples)."""
    from uglad_b200 import main as ug, ops
    g = np.load(os.path.join(ROOT, "tests", "golden", name))
    S = torch.tensor(g["S"], device=dev)
    ops.tune("eig_raw", 0)
    try:
        ops.reset_warm_start()
        model = load_model(g, "p0")
        theta, loss = ug.forward_uGLAD(S, model, L=int(g["L"]), INIT_DIAG=int(g["init_diag"]))
        loss.backward()
        assert rel(theta.detach().cpu().numpy(), g["theta0"]) < THETA_TOL
        for k, p in model.named_parameters():
            assert rel(p.grad.cpu().numpy(), g["g0/" + k]) < 1e-3, k
    finally:
        ops.tune("eig_raw", 1)
        ops.reset_warm_start()


def test_warm_start_does_not_change_the_result(dev):
    from uglad_b200 import main as ug, ops
    g = np.load(CASES[0])
    S = torch.tensor(g["S"], device=dev)
    model = load_model(g, "p0")
    ops.reset_warm_start()
    with torch.no_grad():
        cold = ug.glad.glad(S, model, L=int(g["L"]), INIT_DIAG=int(g["init_diag"])).clone()
        warm = ug.glad.glad(S, model, L=int(g["L"]), INIT_DIAG=int(g["init_diag"])).clone()
    assert rel(warm.cpu().numpy(), cold.cpu().numpy()) < 2e-5


def test_oracle_parity_random_params_and_exact_sqrt(dev):
    """Fresh seeds (not the golden ones), B > 1, both square-root modes, float64 spectral oracle."""
    from uglad_b200 import main as ug, ops
    from uglad_b200.glad.glad_params import GladParams
    rng = np.random.default_rng(21)
    X = rng.random((4, 80, 12))
    S64 = O.covariance(X)
    S = torch.tensor(S64, dtype=torch.float32, device=dev)
    P = O.init_params(21)
    model = GladParams(1.0, 3, 3)
    model.load_state_dict({k: v.detach() for k, v in P.items()})
    for exact in (False, True):
        ops.reset_warm_start()
        r = O.spectral_forward_backward(S.cpu().numpy(), P, L=9, init_diag=0, exact_sqrt=exact)
        model.zero_grad()
        theta = ug.glad.glad(S, model, L=9, exact_sqrt=exact)
        loss = ug.loss_uGLAD(theta, S)
        loss.backward()
        assert rel(theta.detach().cpu().numpy(), r["theta"]) < 2e-5
        assert abs(loss.item() - r["loss"]) < 1e-4 * max(1.0, abs(r["loss"]))
        for k, p in model.named_parameters():
            assert rel(p.grad.cpu().numpy(), r["grads"][k]) < 1e-3, (exact, k)


def test_consensus_mode_loss_broadcasts_full_covariance(dev):
    """main.py:620-622: K sub-sampled covariances in, the full-data covariance in the loss."""
    from uglad_b200 import main as ug, ops
    ops.reset_warm_start()
    rng = np.random.default_rng(8)
    X = rng.random((1, 90, 10))
    Sfull = torch.tensor(O.covariance(X), dtype=torch.float32)
    SK = torch.tensor(O.covariance([X[0][:60], X[0][30:], X[0][::2]]), dtype=torch.float32)
    P = O.init_params(8)
    from uglad_b200.glad.glad_params import GladParams
    model = GladParams(1.0, 3, 3)
    model.load_state_dict({k: v.detach() for k, v in P.items()})
    th_o, loss_o = O.forward_loss(SK, P, 15, 0, loss_S=Sfull)
    loss_o.backward()
    th, loss = ug.forward_uGLAD(SK.to(dev), model, L=15, loss_Sb=Sfull.to(dev))
    loss.backward()
    assert rel(th.detach().cpu().numpy(), th_o.detach().numpy()) < THETA_TOL
    assert abs(loss.item() - loss_o.item()) < 1e-4 * max(1.0, abs(loss_o.item()))
    for k, p in model.named_parameters():
        assert rel(p.grad.cpu().numpy(), P[k].grad.numpy()) < 1e-3, k


def test_indefinite_theta_gives_nan_loss_like_torch_logdet(dev):
    from uglad_b200 import main as ug
    theta = torch.tensor(np.diag([1.0, -2.0, 3.0])[None], dtype=torch.float32, device=dev)
    S = torch.eye(3, device=dev)[None]
    assert torch.isnan(ug.loss_uGLAD(theta, S))
    assert torch.isnan(O.glasso_loss(theta.cpu(), S.cpu()))


def test_fit_api_direct_mode(dev):
    """uGLAD_GL.fit keeps the sklearn-style outputs (main.py:34-151)."""
    from uglad_b200 import main as ug
    from uglad_b200.utils import prepare_data
    rng = np.random.default_rng(9)
    Xb, theta_true = prepare_data.get_data(10, [0.2, 0.2], 300, batch_size=1, eig_offset=1.0, rng=rng)
    m = ug.uGLAD_GL()
    m.fit(Xb[0], centered=False, epochs=5, lr=0.002, L=15, verbose=False, mode="direct")
    assert m.precision_.shape == (10, 10) and m.covariance_.shape == (10, 10) and m.location_.shape == (10,)
    assert np.isfinite(m.precision_).all()
    assert len(m.node_names_) == 10


def test_size_independent_properties_at_full_size(dev):
    """BASELINE configs[2] shape (D=100, many graphs): properties that need no oracle run --
    theta symmetric, permutation of the batch permutes the output."""
    from uglad_b200 import main as ug, ops
    from uglad_b200.utils import prepare_data
    rng = np.random.default_rng(10)
    Xb, _ = prepare_data.get_data(100, [0.05, 0.05], 300, batch_size=16, eig_offset=1.0, rng=rng)
    S = prepare_data.get_covariance(Xb)
    torch.manual_seed(0)
    model, _ = ug.init_uGLAD(lr=0.002)
    ops.reset_warm_start()
    with torch.no_grad():
        th = ug.glad.glad(S, model, L=15).clone()
        perm = torch.randperm(16, device=dev)
        ops.reset_warm_start()
        th_p = ug.glad.glad(S[perm].contiguous(), model, L=15).clone()
    assert rel(th.cpu().numpy(), th.transpose(1, 2).cpu().numpy()) < 1e-5
    assert rel(th_p.cpu().numpy(), th[perm].cpu().numpy()) < 1e-5


# ---- large-D path (Newton-Schulz as dense products + blocked Cholesky) ---------------------------
@pytest.fixture(params=[(1, 1), (1, 0), (0, 1)], ids=["tcgen05-raw", "tcgen05-presplit", "simt"])
def force_large_path(request):
    """Lower the eigensolver/large-D threshold to 0 so that the large-D kernels run on the small
    golden problems too (the threshold only selects the algorithm, never the result); every
    product back-end: tcgen05 3xTF32 on plain operands split in shared memory (the default), on
    pre-split hi/lo pairs, and the FP32 SIMT kernel."""
    from uglad_b200 import ops
    ops.tune("small_d_max", 0)
    ops.tune("use_tc", request.param[0])
    ops.tune("tc_raw", request.param[1])
    ops.reset_warm_start()
    yield
    ops.tune("small_d_max", 166)
    ops.tune("use_tc", 1)
    ops.tune("tc_raw", 1)
    ops.reset_warm_start()


@pytest.mark.parametrize("M,N,K,batch,bn", [
    (128, 128, 32, 1, 0), (100, 100, 100, 3, 0), (7, 5, 3, 2, 0), (200, 200, 200, 2, 0), (130, 70, 45, 2, 64),
    (300, 260, 129, 1, 112), (256, 256, 256, 1, 128), (1000, 1000, 1000, 1, 0), (100, 100, 100, 300, 0),
    (300, 260, 132, 1, 112), (130, 70, 44, 2, 64), (129, 250, 520, 2, 128)])
@pytest.mark.parametrize("raw", [1, 2, 0], ids=["raw-tma-store", "raw-direct-store", "presplit"])
def test_tcgen05_3xtf32_gemm(dev, M, N, K, batch, bn, raw):
    """C = alpha A B^T + beta E1 + diag I on the tensor pipe against float64 numpy: FP32-class
    accuracy (3xTF32), ragged edges, batches larger than the SM count, every tile width; plain
    operands split inside the kernel (K % 4 == 0) and pre-split pairs."""
    import ctypes as C
    from uglad_b200 import _lib, ops
    lib = _lib.load()
    ops.tune("tc_bn", bn)
    ops.tune("tc_raw", 1 if raw else 0)
    ops.tune("tc_tma_store", 1 if raw == 1 else 0)
    try:
        rng = np.random.default_rng(M * 7 + N)
        A = rng.standard_normal((batch, M, K)).astype(np.float32)
        B = rng.standard_normal((batch, N, K)).astype(np.float32)
        E = rng.standard_normal((batch, M, N)).astype(np.float32)
        ref = 0.75 * np.einsum("bmk,bnk->bmn", A.astype(np.float64), B.astype(np.float64)) - 0.5 * E
        ref[:, np.arange(min(M, N)), np.arange(min(M, N))] += 2.0
        dA, dB, dE = (torch.tensor(x, device=dev) for x in (A, B, E))
        out = torch.empty(batch, M, N, device=dev)
        scratch = torch.empty(lib.uglad_tc_gemm_scratch_floats(M, N, K, batch), device=dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = lib.uglad_tc_gemm(dA.data_ptr(), dB.data_ptr(), dE.data_ptr(), out.data_ptr(), M, N, K, batch,
                               0.75, -0.5, 2.0, scratch.data_ptr(), st)
        assert rc == 0, lib.uglad_last_error().decode()
        torch.cuda.synchronize()
        got = out.cpu().numpy().astype(np.float64)
        # error model: entries are ~sqrt(K); the split drops lo*lo (2^-22) and the tensor core adds
        # each K=8 granule into the FP32 accumulator with truncation, so the error grows ~K ulp
        # (measured 1e-5 at K=32, 9e-4 at K=1000); single-pass TF32 would be ~5e-4 sqrt(K)
        tol = 1e-5 * np.sqrt(K) * (1.0 + K / 100.0)
        assert np.abs(got - ref).max() < tol, (np.abs(got - ref).max(), tol)
    finally:
        ops.tune("tc_bn", 0)
        ops.tune("tc_raw", 1)
        ops.tune("tc_tma_store", 1)


def test_tcgen05_gemm_many_tiles_cold_operands(dev):
    """Regression: ~14 short-K tiles per CTA of the raw-operand kernel (three stages, two split
    groups) with operand sets rotating through more memory than L2 holds, so that the TMA loads of
    different stages complete out of order.  A split group that skipped the other group's
    completions used to mistake a stale barrier phase for its own (hang behind the watchdog trap)."""
    import ctypes as C
    from uglad_b200 import _lib, ops
    lib = _lib.load()
    M, N, K, batch = 100, 100, 128, 2048
    g = torch.Generator(device=dev).manual_seed(5)
    sets = [torch.randn(batch, M, K, device=dev, generator=g) for _ in range(3)]
    outs = [torch.empty(batch, M, N, device=dev) for _ in range(3)]
    scratch = torch.empty(max(lib.uglad_tc_gemm_scratch_floats(M, N, K, batch), 1), device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for it in range(45):
        A, out = sets[it % 3], outs[it % 3]
        rc = lib.uglad_tc_gemm(A.data_ptr(), A.data_ptr(), None, out.data_ptr(), M, N, K, batch, 1.0, 0.0, 0.0,
                               scratch.data_ptr(), st)
        assert rc == 0, lib.uglad_last_error().decode()
    torch.cuda.synchronize()
    for A, out in zip(sets, outs):
        ref = torch.einsum("bmk,bnk->bmn", A[:64].double(), A[:64].double())
        assert (out[:64].double() - ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_large_path_forward_backward_against_reference_golden(dev, force_large_path, path):
    from uglad_b200 import main as ug
    g = np.load(path)
    L, idg = int(g["L"]), int(g["init_diag"])
    S = torch.tensor(g["S"], device=dev)
    model = load_model(g, "p0")
    theta, loss = ug.forward_uGLAD(S, model, L=L, INIT_DIAG=idg)
    loss.backward()
    th = theta.detach().cpu().numpy()
    assert rel(th, g["theta0"]) < THETA_TOL
    assert edge_sets_match(th, g["theta0"])
    assert abs(loss.item() - float(g["loss0"])) < 1e-4 * max(1.0, abs(float(g["loss0"])))
    for k, p in model.named_parameters():
        assert rel(p.grad.cpu().numpy(), g["g0/" + k]) < 1e-3, k


@pytest.mark.parametrize("D,B,L", [(233, 2, 4), (320, 1, 3)])
def test_large_path_matches_oracle_above_the_threshold(dev, D, B, L):
    """Sizes the one-CTA eigensolver cannot hold, including a D that is not a multiple of 4."""
    from uglad_b200 import main as ug, ops
    from uglad_b200.glad.glad_params import GladParams
    ops.reset_warm_start()
    rng = np.random.default_rng(D)
    X = rng.random((B, 2 * D, D))
    S = torch.tensor(O.covariance(X), dtype=torch.float32)
    P = O.init_params(D)
    model = GladParams(1.0, 3, 3)
    model.load_state_dict({k: v.detach() for k, v in P.items()})
    th_o, loss_o = O.forward_loss(S, P, L, 0)
    loss_o.backward()
    th, loss = ug.forward_uGLAD(S.to(dev), model, L=L, INIT_DIAG=0)
    loss.backward()
    assert rel(th.detach().cpu().numpy(), th_o.detach().numpy()) < THETA_TOL
    assert abs(loss.item() - loss_o.item()) < 1e-4 * max(1.0, abs(loss_o.item()))
    for k, p in model.named_parameters():
        assert rel(p.grad.cpu().numpy(), P[k].grad.numpy()) < 1e-3, k


@pytest.mark.parametrize("D,B", [(70, 3), (129, 2), (300, 2)])
def test_cholesky_loss_and_gradient(dev, force_large_path, D, B):
    from uglad_b200 import main as ug
    rng = np.random.default_rng(D + 1)
    A = rng.standard_normal((B, D, 2 * D))
    theta = torch.tensor(A @ A.transpose(0, 2, 1) / D + 0.5 * np.eye(D), dtype=torch.float32)
    S = torch.tensor(rng.standard_normal((B, D, D)), dtype=torch.float32)
    S = S + S.transpose(1, 2)
    t_o = theta.clone().requires_grad_(True)
    loss_o = O.glasso_loss(t_o, S)
    loss_o.backward()
    t_g = theta.to(dev).requires_grad_(True)
    loss = ug.loss_uGLAD(t_g, S.to(dev))
    loss.backward()
    assert abs(loss.item() - loss_o.item()) < 1e-5 * max(1.0, abs(loss_o.item()))
    assert rel(t_g.grad.cpu().numpy(), t_o.grad.numpy()) < 2e-5


def test_large_path_indefinite_theta_gives_nan_loss(dev, force_large_path):
    from uglad_b200 import main as ug
    theta = torch.tensor(np.diag([1.0, -2.0, 3.0])[None], dtype=torch.float32, device=dev)
    assert torch.isnan(ug.loss_uGLAD(theta, torch.eye(3, device=dev)[None]))


@pytest.mark.parametrize("M,D", [(5, 12), (40, 70), (300, 40)])
def test_large_path_covariance_repair(dev, force_large_path, M, D):
    """prepare_data.py:345-355 without an eigensolver: Cholesky test + bisection on the shift."""
    from uglad_b200.utils import prepare_data
    rng = np.random.default_rng(M)
    X = rng.standard_normal((2, M, D))
    S = prepare_data.get_covariance(X, offset=0.1).cpu().numpy()
    assert rel(S, O.covariance(X, offset=0.1)) < 2e-5


LARGE = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "large", "*.npz")))


@pytest.mark.parametrize("path", LARGE, ids=[os.path.basename(p)[:-4] for p in LARGE])
def test_d1000_against_reference_golden(dev, path):
    """BASELINE configs[4] (D=1000, L=15) against the real reference run on the CPU
    (tests/golden/make_golden.py large): theta on every 8th row, its Frobenius norm and support
    size, the loss and the 42 gradients."""
    from uglad_b200 import main as ug, ops
    ops.reset_warm_start()
    g = np.load(path)
    D, L, step = int(g["D"]), int(g["L"]), int(g["row_step"])
    S = np.zeros((D, D), np.float32)
    S[np.triu_indices(D)] = g["S_triu"]
    S = S + np.triu(S, 1).T
    model = load_model(g, "p0")
    theta, loss = ug.forward_uGLAD(torch.tensor(S[None], device=dev), model, L=L, INIT_DIAG=0)
    loss.backward()
    th = theta.detach().cpu().numpy()[0]
    assert rel(th[::step], g["theta0_rows"]) < THETA_TOL
    assert edge_sets_match(th[::step], g["theta0_rows"])
    assert abs(np.linalg.norm(th.astype(np.float64)) - float(g["theta0_fro"])) < THETA_TOL * float(g["theta0_fro"])
    assert abs(int((th != 0).sum()) - int(g["theta0_nnz"])) <= 2e-4 * int(g["theta0_nnz"]) + 2
    assert abs(loss.item() - float(g["loss0"])) < 1e-4 * max(1.0, abs(float(g["loss0"])))
    for k, p in model.named_parameters():
        assert rel(p.grad.cpu().numpy(), g["g0/" + k]) < 1e-3, k


# ---- callers either side of the hot path (DESIGN.md row f) ---------------------------------------
def test_fit_api_missing_mode_consensus(dev):
    """uGLAD_GL.fit(mode='missing') (main.py:553-644): mean imputation, K row-subsampled covariances
    trained jointly against the full-data covariance, consensus by majority sign / min magnitude."""
    from uglad_b200 import main as ug
    from uglad_b200.utils import prepare_data
    rng = np.random.default_rng(11)
    Xb, _ = prepare_data.get_data(12, [0.2, 0.2], 240, batch_size=1, eig_offset=1.0, rng=rng)
    Xm = prepare_data.add_noise_dropout(Xb, dropout=0.2, rng=rng)[0]
    m = ug.uGLAD_GL()
    m.fit(Xm, epochs=4, lr=0.002, L=15, verbose=False, mode="missing", k_fold=3)
    assert m.precision_.shape == (12, 12) and np.isfinite(m.precision_).all()
    assert np.allclose(m.precision_, m.precision_.T, atol=1e-5)


def test_fit_api_cv_mode(dev):
    """uGLAD_GL.fit(mode='cv') (main.py:428-550): per fold, keep the parameters with the best held-out loss."""
    from uglad_b200 import main as ug
    from uglad_b200.utils import prepare_data
    rng = np.random.default_rng(12)
    Xb, theta = prepare_data.get_data(10, [0.2, 0.2], 200, batch_size=1, eig_offset=1.0, rng=rng)
    m = ug.uGLAD_GL()
    res = m.fit(Xb[0], true_theta=theta[0], epochs=3, lr=0.002, L=15, verbose=False, mode="cv", k_fold=2)
    assert m.precision_.shape == (10, 10) and np.isfinite(m.precision_).all()
    assert isinstance(res, dict) and "auc" in {k.lower() for k in res}


def test_multitask_api_ragged_sample_counts(dev):
    """uGLAD_multitask.fit (main.py:155-226): one shared model, graphs with different sample counts."""
    from uglad_b200 import main as ug
    from uglad_b200.utils import prepare_data
    rng = np.random.default_rng(13)
    Xs = [prepare_data.get_data(9, [0.2, 0.2], m, batch_size=1, eig_offset=1.0, rng=rng)[0][0] for m in (120, 150, 120)]
    mt = ug.uGLAD_multitask()
    mt.fit(Xs, epochs=3, lr=0.002, L=15, verbose=False)
    assert mt.precision_.shape == (3, 9, 9) and np.isfinite(mt.precision_).all()
    assert mt.covariance_.shape == (3, 9, 9)


def test_covariance_prefetcher_matches_direct_path(dev):
    """The side-stream input pipeline returns exactly what get_covariance returns."""
    from uglad_b200.utils import prepare_data
    rng = np.random.default_rng(14)
    Xs = [torch.from_numpy(rng.random((3, 80, 11)).astype(np.float32)).pin_memory() for _ in range(3)]
    pf = prepare_data.CovariancePrefetcher(dev)
    pf.submit(Xs[0])
    for i in range(3):
        S = pf.get()
        if i + 1 < 3:
            pf.submit(Xs[i + 1])
        torch.cuda.synchronize()
        ref = prepare_data.get_covariance(Xs[i].to(dev))
        assert torch.equal(S, ref)


# ---- edge cases ------------------------------------------------------------------------------------
@pytest.mark.parametrize("D,B,L,idg", [(1, 2, 3, 0), (2, 1, 15, 0), (3, 5, 2, 1), (165, 1, 2, 0), (167, 2, 2, 0),
                                       (167, 1, 2, 1), (10, 600, 3, 0)])
def test_edge_shapes_match_oracle(dev, D, B, L, idg):
    """Degenerate and boundary shapes: D = 1, tiny D, either side of the eigensolver / large-D
    threshold (166), INIT_DIAG on the large-D path, more graphs than the grid has SMs."""
    from uglad_b200 import main as ug, ops
    from uglad_b200.glad.glad_params import GladParams
    ops.reset_warm_start()
    rng = np.random.default_rng(1000 * D + B)
    X = rng.random((B, max(2 * D, 8), D))
    S = torch.tensor(O.covariance(X), dtype=torch.float32)
    P = O.init_params(D + B)
    model = GladParams(1.0, 3, 3)
    model.load_state_dict({k: v.detach() for k, v in P.items()})
    th_o, loss_o = O.forward_loss(S, P, L, idg)
    loss_o.backward()
    th, loss = ug.forward_uGLAD(S.to(dev), model, L=L, INIT_DIAG=idg)
    loss.backward()
    assert rel(th.detach().cpu().numpy(), th_o.detach().numpy()) < THETA_TOL
    assert abs(loss.item() - loss_o.item()) < 1e-4 * max(1.0, abs(loss_o.item()))
    for k, p in model.named_parameters():
        assert rel(p.grad.cpu().numpy(), P[k].grad.numpy()) < 2e-3, k


def test_inputs_need_not_be_contiguous_float32(dev):
    from uglad_b200 import main as ug, ops
    g = np.load(os.path.join(ROOT, "tests", "golden", "d10_m500.npz"))
    ops.reset_warm_start()
    S64 = torch.tensor(g["S"], device=dev, dtype=torch.float64)
    St = S64.transpose(1, 2)  # symmetric: same values, non-contiguous view
    model = load_model(g, "p0")
    with torch.no_grad():
        th = ug.glad.glad(St, model, L=int(g["L"]), INIT_DIAG=int(g["init_diag"]))
    assert rel(th.cpu().numpy(), g["theta0"]) < THETA_TOL


def test_exact_sqrt_is_refused_on_the_large_path(dev):
    from uglad_b200 import main as ug, _lib
    S = torch.eye(170, device=dev)[None] * 0.5
    model = ug.init_uGLAD(lr=0.002)[0]
    with pytest.raises(_lib.UgladError):
        ug.glad.glad(S, model, L=2, exact_sqrt=True)


def test_cpu_tensors_are_refused(dev):
    from uglad_b200 import main as ug, _lib
    model = ug.init_uGLAD(lr=0.002)[0]
    with pytest.raises(_lib.UgladError):
        ug.glad.glad(torch.eye(4)[None], model, L=2)


def test_warm_started_conditioning_gives_the_same_covariance(dev):
    """A warm start (eigenvectors of a similar earlier batch) only changes the work, not the result,
    including when the rank-deficient repair triggers."""
    from uglad_b200.utils import prepare_data
    rng = np.random.default_rng(15)
    X0 = rng.standard_normal((3, 6, 14))            # fewer samples than features: repair triggers
    X1 = X0 + 1e-3 * rng.standard_normal(X0.shape)
    S0 = prepare_data.get_covariance(X0)
    cold = prepare_data.get_covariance(X1)
    warm = prepare_data.get_covariance(X1, warm=S0)
    assert rel(warm.cpu().numpy(), cold.cpu().numpy()) < 1e-6
    assert rel(warm.cpu().numpy(), O.covariance(X1, offset=0.1)) < 1e-5
    w_c, w_w = cold._uglad_eig[1].wS.cpu().numpy(), warm._uglad_eig[1].wS.cpu().numpy()
    assert np.abs(np.sort(w_c, 1) - np.sort(w_w, 1)).max() < 1e-5


def test_soft_threshold_propagates_nan_like_torch(dev):
    """glad_params.py:81 is sign(X) * max(0, |X| - rho) with torch.max / torch.sign, which propagate NaN; the
    reference's NaN stop (main.py:405-409) depends on a diverged fit turning the loss NaN."""
    from uglad_b200 import ops
    rng = np.random.default_rng(9)
    D = 12
    X = torch.tensor(rng.standard_normal((1, D, D)), dtype=torch.float32)
    X[0, 3, 4] = float("nan")
    S = torch.tensor(rng.standard_normal((1, D, D)), dtype=torch.float32)
    T = torch.tensor(rng.standard_normal((1, D, D)), dtype=torch.float32)
    P = O.init_params(2)
    flat = torch.cat([P[k].detach().reshape(-1) for k in O.PARAM_KEYS])
    Z, _ = ops.z_update(X.to(dev), S.to(dev), T.to(dev), flat.to(dev), H=3)
    Z = Z.cpu().numpy()
    assert np.isnan(Z[0, 3, 4]) and np.isfinite(np.delete(Z.reshape(-1), 3 * D + 4)).all()
    want = O.eta_threshold(P, X, S, T).detach().numpy()   # the oracle's torch.sign * torch.max form
    assert np.isnan(want[0, 3, 4])
    mask = ~np.isnan(want)
    assert np.allclose(Z[mask], want[mask], rtol=1e-5, atol=1e-6)

"""Generate golden vectors by running the REAL reference (imported from /root/reference).

Run here (the container that has /root/reference); the GPU box only sees the committed
.npz files.   python tests/golden/make_golden.py

matplotlib and pyvis (plotting only, absent from this image) are stubbed so that
uglad.main imports; nothing on the numeric path is touched.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    _stub("matplotlib", colors=_stub("matplotlib.colors"), pyplot=_stub(
        "matplotlib.pyplot", figure=lambda *a, **k: None, plot=lambda *a, **k: None,
        xlabel=lambda *a, **k: None, ylabel=lambda *a, **k: None, title=lambda *a, **k: None,
        legend=lambda *a, **k: None, grid=lambda *a, **k: None, savefig=lambda *a, **k: None,
        show=lambda *a, **k: None))
    _stub("pyvis", network=_stub("pyvis.network", Network=object))
    sys.path.insert(0, REF)
    import uglad.main as ref_main
    from uglad.glad import glad as ref_glad
    from uglad.glad.glad_params import GladParams
    from uglad.utils import prepare_data
    return ref_main, ref_glad, GladParams, prepare_data


def state_np(model):
    return {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}


def grads_np(model):
    return {k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()}


def make_case(name, ref_main, GladParams, prepare_data, D, M, B, L, init_diag, seed,
              epochs, lr, sparsity=(0.2, 0.2), dropout=0.0, minmax=True, store_X=True, rho_bias=None):
    np.random.seed(seed)
    Xb, true_theta = prepare_data.get_data(num_nodes=D, sparsity=list(sparsity), num_samples=M,
                                           batch_size=B, eig_offset=1, w_min=0.5, w_max=1)
    if minmax:  # what uGLAD_GL.fit feeds the solver (process_table NORM="min_max")
        Xb = np.array([(X - X.min(0)) / (X.max(0) - X.min(0)) for X in Xb])
    Sb64 = prepare_data.get_covariance(Xb, offset=0.1)
    Sb = prepare_data.convert_to_torch(Sb64, req_grad=False)
    torch.manual_seed(seed)
    model = GladParams(theta_init_offset=1.0, nF=3, H=3)
    if rho_bias is not None:  # a low threshold, as after long training: theta keeps off-diagonal support
        with torch.no_grad():
            model.rho_l1[4].bias.fill_(rho_bias)
    out = {"S": Sb.numpy(), "L": L, "init_diag": init_diag, "seed": seed, "lr": lr, "epochs": epochs}
    if store_X:
        out.update({"X": Xb.astype(np.float64), "S64": Sb64, "true_theta": true_theta})
    for k, v in state_np(model).items():
        out["p0/" + k] = v
    # one forward/backward at the initial parameters
    theta, loss = ref_main.forward_uGLAD(Sb, model, L=L, INIT_DIAG=init_diag)
    loss.backward()
    out["theta0"] = theta.detach().numpy()
    out["loss0"] = np.float64(loss.item())
    for k, v in grads_np(model).items():
        out["g0/" + k] = v
    # training trajectory (main.py:389-414)
    opt = ref_main.glad.get_optimizers(model, lr_glad=lr)
    losses = []
    for _ in range(epochs):
        opt.zero_grad()
        theta, loss = ref_main.forward_uGLAD(Sb, model, L=L, INIT_DIAG=init_diag)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    out["losses"] = np.asarray(losses)
    out["thetaT"] = theta.detach().numpy()
    for k, v in state_np(model).items():
        out["pT/" + k] = v
    if B > 1:
        out["consensus"] = ref_main.get_final_precision_from_batch(theta.detach(), type="min").numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: D={D} B={B} L={L} loss0={out['loss0']:.6f} lossT={losses[-1]:.6f}")


def make_large_case(name, ref_main, GladParams, prepare_data, D, M, L, seed, row_step=8, rho_bias=-3.5):
    """BASELINE configs[4] shape (D=1000, M=10000, L=15).  To keep the fixture small only the
    upper triangle of S (the input) and every `row_step`-th row of the reference's theta are
    stored, plus its Frobenius norm, loss, lambda-free scalars and the 42 gradients."""
    np.random.seed(seed)
    Xb, _ = prepare_data.get_data(num_nodes=D, sparsity=[0.01, 0.01], num_samples=M, batch_size=1,
                                  eig_offset=1, w_min=0.5, w_max=1)
    Xb = np.array([(X - X.min(0)) / (X.max(0) - X.min(0)) for X in Xb])
    Sb = prepare_data.convert_to_torch(prepare_data.get_covariance(Xb, offset=0.1), req_grad=False)
    torch.manual_seed(seed)
    model = GladParams(theta_init_offset=1.0, nF=3, H=3)
    with torch.no_grad():  # a low threshold (rho ~ 0.03) so that theta keeps a non-trivial support
        model.rho_l1[4].bias.fill_(rho_bias)
    out = {"S_triu": Sb.numpy()[0][np.triu_indices(D)], "D": D, "L": L, "init_diag": 0, "seed": seed,
           "row_step": row_step}
    for k, v in state_np(model).items():
        out["p0/" + k] = v
    theta, loss = ref_main.forward_uGLAD(Sb, model, L=L, INIT_DIAG=0)
    loss.backward()
    th = theta.detach().numpy()[0]
    out["theta0_rows"] = th[::row_step].copy()
    out["theta0_fro"] = np.float64(np.linalg.norm(th.astype(np.float64)))
    out["theta0_nnz"] = np.int64((th != 0).sum())
    out["loss0"] = np.float64(loss.item())
    for k, v in grads_np(model).items():
        out["g0/" + k] = v
    os.makedirs(os.path.join(HERE, "large"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "large", name + ".npz"), **out)
    print(f"{name}: D={D} L={L} loss0={out['loss0']:.6f} nnz={int(out['theta0_nnz'])}")


def _quiet(fn, *a, **k):
    """Run a reference driver with its progress prints swallowed."""
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def make_mode_cases(ref_main, prepare_data):
    """The reference's own mode drivers through its sklearn-style wrappers (main.py:34-226):
    direct with a structure prior, cv, missing (consensus) and multitask -- small D, few epochs.
    torch.manual_seed(seed) right before fit() fixes the initial GladParams (init_uGLAD runs inside)."""
    from uglad.utils import metrics as ref_metrics
    out = {}
    # --- direct mode with true_theta: the log-cosh structure prior enters the loss (main.py:325-334, :397-403)
    np.random.seed(31)
    Xb, theta = prepare_data.get_data(num_nodes=10, sparsity=[0.2, 0.2], num_samples=300, batch_size=1,
                                      eig_offset=1, w_min=0.5, w_max=1)
    torch.manual_seed(31)
    m = ref_main.uGLAD_GL()
    cmp_ = _quiet(m.fit, Xb[0].copy(), true_theta=theta[0], epochs=10, lr=0.01, L=15, verbose=False, mode="direct")
    out.update({"direct/X": Xb[0], "direct/true_theta": theta[0], "direct/precision": m.precision_,
                "direct/covariance": m.covariance_, "direct/location": m.location_,
                "direct/metric_keys": np.array(sorted(cmp_)), "direct/metric_vals": np.array([cmp_[k] for k in sorted(cmp_)])})
    # --- cv mode (main.py:428-550), 3 folds
    np.random.seed(32)
    Xb, theta = prepare_data.get_data(num_nodes=9, sparsity=[0.2, 0.2], num_samples=210, batch_size=1,
                                      eig_offset=1, w_min=0.5, w_max=1)
    torch.manual_seed(32)
    m = ref_main.uGLAD_GL()
    _quiet(m.fit, Xb[0].copy(), epochs=10, lr=0.01, L=15, verbose=False, mode="cv", k_fold=3)
    out.update({"cv/X": Xb[0], "cv/precision": m.precision_})
    # --- missing mode (main.py:553-644): 20 % NaN dropout, K = 4 row-subsampled batches, consensus
    np.random.seed(33)
    Xb, theta = prepare_data.get_data(num_nodes=12, sparsity=[0.2, 0.2], num_samples=242, batch_size=1,
                                      eig_offset=1, w_min=0.5, w_max=1)
    Xm = prepare_data.add_noise_dropout(Xb, dropout=0.2)[0]
    torch.manual_seed(33)
    m = ref_main.uGLAD_GL()
    _quiet(m.fit, Xm.copy(), epochs=10, lr=0.01, L=15, verbose=False, mode="missing", k_fold=4)
    out.update({"missing/X": Xm, "missing/precision": m.precision_, "missing/covariance": m.covariance_})
    # --- multitask (main.py:155-226, :719-789): ragged sample counts, one shared model
    np.random.seed(34)
    Xs = [prepare_data.get_data(num_nodes=8, sparsity=[0.2, 0.3], num_samples=mm, batch_size=1, eig_offset=1,
                                w_min=0.5, w_max=1)[0][0] for mm in (120, 150, 120)]
    torch.manual_seed(34)
    mt = ref_main.uGLAD_multitask()
    _quiet(mt.fit, [x.copy() for x in Xs], epochs=10, lr=0.01, L=15, verbose=False)
    for i, x in enumerate(Xs):
        out[f"multitask/X{i}"] = x
    out.update({"multitask/precision": mt.precision_, "multitask/covariance": mt.covariance_})
    # --- metrics.report_metrics_all (metrics.py:25-108) on random sparse symmetric pairs
    rng = np.random.default_rng(35)
    for i in range(4):
        d = 12 + 3 * i
        tg = np.triu((rng.random((d, d)) < 0.25) * rng.uniform(0.5, 1.0, (d, d)), 1)
        pg = np.triu((rng.random((d, d)) < 0.3) * rng.standard_normal((d, d)), 1)
        pg = np.where(rng.random((d, d)) < 0.5, pg, np.triu(tg * rng.standard_normal((d, d)), 1))
        tg, pg = tg + tg.T + np.eye(d), pg + pg.T + np.eye(d)
        r = ref_metrics.report_metrics_all(tg, pg)
        out[f"metrics/{i}/true"], out[f"metrics/{i}/pred"] = tg, pg
        out[f"metrics/{i}/keys"] = np.array(sorted(r))
        out[f"metrics/{i}/vals"] = np.array([r[k] for k in sorted(r)])
    # --- process_table (prepare_data.py:361-516): zero rows, NaNs, a constant and a duplicated column; and the
    #     condition-number pruning loop on nearly collinear columns
    import pandas as pd
    rng = np.random.default_rng(36)
    T = rng.random((40, 8))
    T[3] = 0.0
    T[17] = 0.0
    T[5, 2] = np.nan
    T[9, 6] = np.nan
    T[:, 4] = 0.7
    T[:, 7] = T[:, 1]
    out["table/raw"] = T
    out["table/minmax"] = np.array(_quiet(prepare_data.process_table, pd.DataFrame(T.copy()), NORM="min_max", VERBOSE=False))
    out["table/mean"] = np.array(_quiet(prepare_data.process_table, pd.DataFrame(T.copy()), NORM="mean", VERBOSE=False))
    Z = rng.standard_normal((60, 9))
    Z[:, 6] = Z[:, 0] + 1e-3 * rng.standard_normal(60)
    Z[:, 7] = Z[:, 1] - Z[:, 2] + 1e-3 * rng.standard_normal(60)
    Z[:, 8] = Z[:, 3] + Z[:, 4] + 1e-2 * rng.standard_normal(60)
    out["table/collinear"] = Z
    pruned = _quiet(prepare_data.process_table, pd.DataFrame(Z.copy()), NORM="min_max", COND_NUM=200.0, eigval_th=1e-3, VERBOSE=False)
    out["table/collinear_kept"] = np.array(list(pruned.columns), dtype=np.int64)
    out["table/collinear_out"] = np.array(pruned)
    os.makedirs(os.path.join(HERE, "modes"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "modes", "modes.npz"), **out)
    print("modes: direct(struct prior) / cv / missing / multitask precision_, metrics, process_table;",
          "kept columns under COND_NUM=200:", out["table/collinear_kept"].tolist())


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "modes":
        torch.set_num_threads(1)
        ref_main, ref_glad, GladParams, prepare_data = import_reference()
        make_mode_cases(ref_main, prepare_data)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "large":
        ref_main, ref_glad, GladParams, prepare_data = import_reference()
        make_large_case("d1000_m10000", ref_main, GladParams, prepare_data, D=1000, M=10000, L=15, seed=17,
                        rho_bias=float(sys.argv[2]) if len(sys.argv) > 2 else -7.0)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "lowrho":
        torch.set_num_threads(1)
        ref_main, ref_glad, GladParams, prepare_data = import_reference()
        # thresholds low enough that the soft-threshold keeps off-diagonal entries (the edge set is
        # non-trivial and moves during training)
        make_case("d20_lowrho", ref_main, GladParams, prepare_data, D=20, M=400, B=2, L=15, init_diag=0,
                  seed=18, epochs=12, lr=0.01, sparsity=(0.15, 0.15), store_X=False, rho_bias=float(sys.argv[2]))
        make_case("d100_lowrho", ref_main, GladParams, prepare_data, D=100, M=1000, B=1, L=15, init_diag=0,
                  seed=19, epochs=4, lr=0.005, sparsity=(0.05, 0.05), store_X=False, rho_bias=float(sys.argv[2]))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "d256":
        torch.set_num_threads(1)
        ref_main, ref_glad, GladParams, prepare_data = import_reference()
        # above the one-CTA eigensolver's range: exercises the Newton-Schulz GEMM path natively
        make_case("d256_b2", ref_main, GladParams, prepare_data, D=256, M=600, B=2, L=15,
                  init_diag=0, seed=16, epochs=3, lr=0.002, sparsity=(0.03, 0.03), store_X=False)
        return
    torch.set_num_threads(1)  # deterministic reduction order
    ref_main, ref_glad, GladParams, prepare_data = import_reference()
    # configs[0]: demo-notebook scale
    make_case("d10_m500", ref_main, GladParams, prepare_data, D=10, M=500, B=1, L=15,
              init_diag=0, seed=11, epochs=40, lr=0.002)
    make_case("d20_b3_multitask", ref_main, GladParams, prepare_data, D=20, M=500, B=3, L=15,
              init_diag=0, seed=12, epochs=30, lr=0.01, sparsity=(0.1, 0.2))
    make_case("d16_initdiag1", ref_main, GladParams, prepare_data, D=16, M=200, B=2, L=7,
              init_diag=1, seed=13, epochs=10, lr=0.002)
    # configs[1] shape (single graph D=100, M=1000, L=15); short trajectory to bound size
    make_case("d100_m1000", ref_main, GladParams, prepare_data, D=100, M=1000, B=1, L=15,
              init_diag=0, seed=14, epochs=6, lr=0.002, sparsity=(0.05, 0.05))
    # raw (un-normalised) data exercises larger-magnitude S
    make_case("d12_raw", ref_main, GladParams, prepare_data, D=12, M=300, B=1, L=15,
              init_diag=0, seed=15, epochs=10, lr=0.002, minmax=False)


if __name__ == "__main__":
    main()

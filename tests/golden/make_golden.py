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


def main():
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

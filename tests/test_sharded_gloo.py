"""The graph-sharded protocol of uglad_b200.ops on two CPU processes (gloo): every rank holds half
of the graphs; the only exchanged values are the graph count, one scalar per layer (the local sum
of ||Z - X||_F^2) and the packed parameter gradient.  The per-shard layer arithmetic is supplied
by the oracle here (the CUDA kernels need a GPU); what is under test is the host-side protocol that
bench.py --gpus N and uGLAD_multitask.fit(group=...) run on NCCL: with it, two shards must
reproduce the single-process result on all graphs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import uglad_oracle as O

L_LAYERS = 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, S_all, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uglad_b200 import ops
    torch.set_num_threads(1)
    group = dist.group.WORLD
    shard = np.array_split(np.arange(S_all.shape[0]), world)[rank]
    S = S_all[shard]
    P = O.init_params(3)
    B_total = ops.global_graph_count(S.shape[0], torch.device("cpu"), group)
    assert B_total == S_all.shape[0]
    eye = torch.eye(S.shape[-1]).expand_as(S)
    state = {"theta": torch.linalg.inv(S + P["theta_init_offset"] * eye), "lam": O.lambda_net(P, 1.0, 0.0)}
    normf = torch.zeros(L_LAYERS)

    def layer(k):
        if k > 0:  # lambda_k from the all-reduced sum of the previous layer (glad.py:146-150)
            state["lam"] = O.lambda_net(P, float(normf[k - 1]) / B_total, state["lam"].item())
        lam, theta = state["lam"], state["theta"]
        b = (1.0 / lam) * S - theta
        x = 0.5 * (O.ns_sqrt(b.transpose(-1, -2) @ b + (4.0 / lam) * eye) - b)
        state["theta"] = O.eta_threshold(P, x, S, theta)
        normf[k] = float(torch.sum((state["theta"] - x) ** 2))

    ops.run_sharded_layers(L_LAYERS, layer, normf, group)
    theta = state["theta"]
    loss = O.glasso_loss(theta, S) * S.shape[0] / B_total          # local sum / global count
    loss.backward()
    gp = torch.cat([P[k].grad.reshape(-1) for k in O.PARAM_KEYS])
    ops.allreduce_shared_gradients(gp, group)
    tot = loss.detach().clone().reshape(1)
    dist.all_reduce(tot, group=group)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), theta=theta.detach().numpy(), shard=shard, grad=gp.numpy(),
             loss=tot.numpy())
    dist.destroy_process_group()


def test_two_shards_reproduce_the_single_process_result(tmp_path):
    rng = np.random.default_rng(0)
    X = rng.random((5, 60, 7))  # 5 graphs: an uneven 3 + 2 split
    S_all = torch.tensor(O.covariance(X), dtype=torch.float32)
    mp.spawn(_worker, args=(2, _free_port(), S_all, str(tmp_path)), nprocs=2, join=True)
    P = O.init_params(3)
    theta, loss = O.forward_loss(S_all, P, L_LAYERS, 0)
    loss.backward()
    grad = torch.cat([P[k].grad.reshape(-1) for k in O.PARAM_KEYS]).numpy()
    got = np.zeros_like(theta.detach().numpy())
    for r in range(2):
        g = np.load(tmp_path / f"rank{r}.npz")
        got[g["shard"]] = g["theta"]
        assert np.allclose(g["grad"], grad, rtol=2e-4, atol=1e-6)
        assert abs(float(g["loss"][0]) - loss.item()) < 1e-5 * max(1.0, abs(loss.item()))
    assert np.allclose(got, theta.detach().numpy(), rtol=1e-5, atol=1e-6)


def _consensus_worker(rank, world, port, theta_all, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uglad_b200 import main as ug
    shard = np.array_split(np.arange(theta_all.shape[0]), world)[rank]
    out = ug.get_final_precision_from_batch(theta_all[shard].clone(), type="min", group=dist.group.WORLD)
    np.save(os.path.join(out_dir, f"cons{rank}.npy"), out.numpy())
    dist.destroy_process_group()


def test_sharded_consensus_is_bit_identical_to_the_single_process_result(tmp_path):
    """main.py:673-716 over ranks: K = 5 precision matrices split 3 + 2; all-reduce(MIN) of |theta| and
    all-reduce(SUM) of sign(theta) are exact, so every rank holds the single-process consensus."""
    rng = np.random.default_rng(4)
    theta_all = torch.tensor(rng.standard_normal((5, 9, 9)), dtype=torch.float32)
    theta_all[:, 2, 3] = torch.tensor([0.5, -0.5, 0.0, 0.25, -0.25])     # a sign tie resolves to +
    theta_all[:, 4, 4] = 0.0
    mp.spawn(_consensus_worker, args=(2, _free_port(), theta_all, str(tmp_path)), nprocs=2, join=True)
    want = O.consensus_min(theta_all).numpy()
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"cons{r}.npy"), want)

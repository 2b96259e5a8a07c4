"""Host-side invariants that need no GPU."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_graphed_step_epoch_closure_does_not_reference_the_instance():
    """ops.GraphedStep stores its epoch closure on the instance; if the closure referenced `self` the instance would
    sit in a reference cycle and its CUDA graphs -- which capture the NCCL gradient all-reduce of a sharded epoch --
    would stay alive until the cyclic collector runs.  An 8-rank bench hung for six minutes in
    dist.destroy_process_group() that way: the communicator waits for every graph that captured it."""
    tree = ast.parse(open(os.path.join(ROOT, "uglad_b200", "ops.py")).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "GraphedStep")
    init = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "__init__")
    epoch = next(n for n in ast.walk(init) if isinstance(n, ast.FunctionDef) and n.name == "epoch")
    names = {n.id for n in ast.walk(epoch) if isinstance(n, ast.Name)}
    assert "self" not in names
    assert any(isinstance(n, ast.FunctionDef) and n.name == "close" for n in cls.body)


def test_bench_tears_down_in_order_behind_a_watchdog():
    src = open(os.path.join(ROOT, "bench.py")).read()
    tail = src[src.rindex("if world > 1:"):]
    assert tail.index("threading.Timer") < tail.index("gc.collect()") < tail.index("dist.barrier") < tail.index("destroy_process_group")
    assert "gs.close()" in src


def test_eig_path_policy():
    """uglad_eig_path(B, D): the eigensolver path up to small_d_max (166), and up to D = 200 (D % 4 == 0) while every
    graph's warm solves fit a resident cluster of 4 CTAs (4 B <= 132 on a 148-SM part; without a device the library
    assumes 148 SMs).  A pure predicate: callable without a GPU."""
    from uglad_b200 import _lib
    lib = _lib.load()
    assert lib.uglad_small_d_max() == 166
    for B, D, want in [(1, 10, 1), (256, 100, 1), (256, 166, 1), (1, 167, 0), (32, 200, 1), (33, 200, 1), (34, 200, 0),
                       (256, 200, 0), (4, 198, 0), (4, 204, 0), (1, 1000, 0), (1, 168, 1)]:
        assert lib.uglad_eig_path(B, D) == want, (B, D)

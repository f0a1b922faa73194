"""The shipped examples run end to end on the CPU backend (BASELINE.json config 1:
"simple_linear_regression.py Allreduce(MPI_SUM) world_size=2 on CPU")."""
import re

from conftest import run_spmd


def test_linear_regression_example_world_size_2_cpu():
    res = run_spmd(2, ["examples/simple_linear_regression.py"], device="cpu", timeout=300)
    assert res.returncode == 0, res.stderr[-4000:]
    m = re.search(r"Final parameters: \[([^\]]+)\]", res.stdout)
    assert m, res.stdout[-2000:]
    params = [float(v) for v in m.group(1).split(",")]
    for got, want in zip(params, [0.1, 1.0, -2.0]):
        assert abs(got - want) < 1e-5, params
    assert "10 closure evaluations" in res.stdout  # same optimiser trajectory as a single process


def test_isend_recv_wait_example_gradient_is_two():
    res = run_spmd(3, ["examples/isend_recv_wait.py"], device="cpu", timeout=300)
    assert res.returncode == 0, res.stderr[-4000:]
    assert res.stdout.count("a.grad = [2.0]") == 3


def test_dp_linear_example_loss_decreases_cpu():
    res = run_spmd(2, ["examples/dp_linear_layer.py", "--device", "cpu", "--features", "64", "--batch", "128", "--steps", "8"],
                   device="cpu", timeout=300)
    assert res.returncode == 0, res.stderr[-4000:]
    losses = [float(v) for v in re.findall(r"loss ([0-9.]+)", res.stdout)]
    assert len(losses) == 8 and losses[-1] < losses[0]


def test_tensor_parallel_mlp_example_2d_parallelism_cpu():
    res = run_spmd(4, ["examples/tensor_parallel_mlp.py", "--tp", "2", "--device", "cpu", "--steps", "30"], device="cpu",
                   timeout=300)
    assert res.returncode == 0, res.stderr[-4000:]
    losses = [float(v) for v in re.findall(r"loss ([0-9.]+)", res.stdout)]
    assert len(losses) >= 3 and losses[-1] < 0.7 * losses[0], losses
    assert "(tp=2, dp=2)" in res.stdout


def test_pipeline_example_trains_across_three_stages_cpu():
    res = run_spmd(3, ["examples/pipeline_mlp.py", "--device", "cpu", "--steps", "40"], device="cpu", timeout=300)
    assert res.returncode == 0, res.stderr[-4000:]
    losses = [float(v) for v in re.findall(r"loss ([0-9.]+)", res.stdout)]
    assert len(losses) >= 3 and losses[-1] < 0.6 * losses[0], losses
    assert "(3 stages, 4 micro-batches)" in res.stdout


def test_sharded_optimizer_example_cpu():
    res = run_spmd(3, ["examples/sharded_optimizer.py", "--device", "cpu", "--steps", "30"], device="cpu", timeout=300)
    assert res.returncode == 0, res.stderr[-4000:]
    losses = [float(v) for v in re.findall(r"loss ([0-9.]+)", res.stdout)]
    assert len(losses) >= 3 and losses[-1] < 0.5 * losses[0], losses
    assert "optimizer state per rank" in res.stdout


def test_multinode_two_level_example_on_simulated_nodes():
    res = run_spmd(4, ["examples/multinode_two_level.py", "--device", "cpu", "--steps", "21"], device="cpu", timeout=300,
                   extra_env={"M4T_NET": "1", "M4T_NET_LOCAL_SIZE": "2"})
    assert res.returncode == 0, res.stderr[-4000:]
    assert "4 ranks = 2 node(s) x 2; node: cpu: posix-shm" in res.stdout
    losses = [float(v) for v in re.findall(r"loss ([0-9.]+)", res.stdout)]
    assert len(losses) == 3 and losses[-1] < 0.5 * losses[0], losses

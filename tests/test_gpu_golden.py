"""GPU path against the committed golden fixtures: outputs of the UNMODIFIED reference
(tests/golden/make_golden.py).  No oracle in between for the env dynamics."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-5


def close(got, want, scale, rtol=RTOL):
    np.testing.assert_allclose(np.asarray(got, np.float64), want, rtol=rtol, atol=rtol * scale + 1e-30)


def kernel_layout(a, ld):
    T, E, R = a.shape
    out = torch.zeros(T, R, ld, dtype=torch.as_tensor(a).dtype, device="cuda")
    out[:, :, :E] = torch.as_tensor(a, device="cuda").permute(0, 2, 1)
    return out


def check_accounting(g, out, scale):
    assert np.array_equal(out["C"].cpu().numpy(), g["C"].astype(np.int64))
    close(out["R"].cpu().numpy(), g["R"], scale)
    close(out["modR"].cpu().numpy(), g["modR"], scale)
    close(out["G"].cpu().numpy(), g["G"], scale)


@pytest.mark.parametrize("name", ["coverage_c1.npz", "coverage_c4.npz"])
def test_coverage_vs_reference_outputs(name):
    import safe_multiagent_rl_b200 as s
    g = np.load(os.path.join(GOLD, name))
    size, A = int(g["size"]), int(g["n_agents"])
    T, E = g["actions"].shape[:2]
    env = s.BatchedCoverageDiscrete(size, A, n_envs=E, weights=g["weights"].tolist(), starts=g["starts"])
    assert env.fieldview_size == float(g["fieldview"])
    lam = torch.as_tensor(g["lambdas"], dtype=torch.float64, device="cuda")
    buf = env.new_rollout_buffer(T)
    env.reset()
    for t in range(T):
        obs, r, c, d = env.step(g["actions"][t].astype(np.uint8), lambdas=lam, out=(buf, t))
        assert np.array_equal(env.state().cpu().numpy(), g["pos"][t])
        assert np.array_equal(c.cpu().numpy(), g["cost"][t]) and not d.any()
        close(r.cpu().numpy(), g["reward"][t], np.abs(g["reward"][t]).max())
    scale = np.abs(g["modR"]).max()
    check_accounting(g, buf.finish(float(g["gamma"]), g["thresholds"]), scale)
    fused = env.rollout(kernel_layout(g["actions"].astype(np.uint8), env.ld), lambdas=lam, gamma=float(g["gamma"]),
                        thresholds=g["thresholds"], g_mode=1)
    check_accounting(g, fused, scale)
    assert np.array_equal(env.state().cpu().numpy(), g["pos"][-1])
    # disc terms (AbstractAgent.compute_returns)
    fused2 = env.rollout(kernel_layout(g["actions"].astype(np.uint8), env.ld), lambdas=lam, gamma=float(g["gamma"]), g_mode=2)
    close(fused2["G"].cpu().numpy(), g["disc"], scale)
    # lambda update from the reference's own MetaAgent on env 0's episode (one episode recorded)
    meta = s.BatchedMetaAgent([1] * A, float(g["gamma"]), float(g["meta_lr"]), g["thresholds"], start_learning_cycle=0,
                              n_agents=A)
    meta.lambdas.copy_(lam)
    one = s.BatchedCoverageDiscrete(size, A, n_envs=1, weights=g["weights"].tolist(), starts=g["starts"][:1])
    o1 = one.rollout(kernel_layout(g["actions"][:, :1].astype(np.uint8), one.ld), lambdas=lam, gamma=float(g["gamma"]),
                     thresholds=g["thresholds"])
    meta.step(o1["stats"])
    meta.update()
    np.testing.assert_allclose(meta.lambdas.cpu().numpy(), g["lam_after"][0], rtol=1e-14)


@pytest.mark.parametrize("name", ["congestion_paper.npz", "congestion_c3.npz"])
def test_congestion_vs_reference_outputs(name):
    import safe_multiagent_rl_b200 as s
    g = np.load(os.path.join(GOLD, name))
    size, A = int(g["size"]), int(g["n_agents"])
    T, E = g["actions"].shape[:2]
    env = s.BatchedCongestion(size, A, n_envs=E, noise=float(g["noise"]), starts=g["starts"], demand_rate=g["demand"],
                              seed=int(g["philox_seed"]), env_offset=int(g["env_offset"]))
    lam = torch.as_tensor(g["lambdas"], dtype=torch.float64, device="cuda")
    buf = env.new_rollout_buffer(T)
    env.reset()
    for t in range(T):
        obs, r, c, d = env.step(g["actions"][t].astype(np.uint8), lambdas=lam, out=(buf, t))   # on-device Philox noise
        assert np.array_equal(env.state().cpu().numpy(), g["pos"][t])
        assert np.array_equal(c.cpu().numpy(), g["cost"][t])
        assert np.array_equal(r.cpu().numpy(), g["reward"][t].astype(np.float32))               # f64-exact, rounded once
    scale = np.abs(g["modR"]).max()
    check_accounting(g, buf.finish(float(g["gamma"]), g["thresholds"]), scale)
    env.set_noise_episode(0)            # the fixture holds episode 0's uniforms; a second episode would draw fresh noise
    fused = env.rollout(kernel_layout(g["actions"].astype(np.uint8), env.ld), lambdas=lam, gamma=float(g["gamma"]),
                        thresholds=g["thresholds"], g_mode=1)
    check_accounting(g, fused, scale)
    assert np.array_equal(env.state().cpu().numpy(), g["pos"][-1])


@pytest.mark.parametrize("name", ["collision_c2.npz", "collision_paper.npz"])
def test_collision_vs_reference_outputs(name):
    import safe_multiagent_rl_b200 as s
    g = np.load(os.path.join(GOLD, name))
    size, A, L = int(g["size"]), int(g["n_agents"]), int(g["n_landmarks"])
    T, E = g["actions"].shape[:2]
    env = s.BatchedCollisionAvoidance(size, A, n_envs=E, n_landmarks=L, starts=g["starts"], landmarks=g["landmarks"])
    lam = torch.as_tensor(g["lambdas"], dtype=torch.float64, device="cuda")
    buf = env.new_rollout_buffer(T)
    env.reset()
    for t in range(T):
        obs, r, c, d = env.step(g["actions"][t], lambdas=lam, out=(buf, t))
        assert np.array_equal(env.state().cpu().numpy(), g["pos"][t])                           # bit-exact float64
        assert np.array_equal(d.cpu().numpy().astype(bool), g["done"][t])
        assert np.array_equal(c.cpu().numpy(), g["cost"][t].astype(np.int64))
        assert np.array_equal(r.cpu().numpy(), g["reward"][t].astype(np.float32))
    scale = np.abs(g["modR"]).max()
    check_accounting(g, buf.finish(float(g["gamma"]), g["thresholds"]), scale)
    fused = env.rollout(kernel_layout(g["actions"].reshape(T, E, 2 * A), env.ld), lambdas=lam, gamma=float(g["gamma"]),
                        thresholds=g["thresholds"], g_mode=1)
    check_accounting(g, fused, scale)
    assert np.array_equal(fused["n_active"].cpu().numpy(), g["active"].sum(0))
    assert np.array_equal(env.state().cpu().numpy(), g["pos"][-1])


@pytest.mark.parametrize("name", ["coverage_continuous_paper.npz", "coverage_discretized.npz"])
def test_coverage_float_vs_reference_outputs(name):
    import safe_multiagent_rl_b200 as s
    g = np.load(os.path.join(GOLD, name))
    size, A, coarse = int(g["size"]), int(g["n_agents"]), int(g["coarseness"])
    T, E = g["actions"].shape[:2]
    if str(g["env"]) == "continuous":
        env = s.BatchedCoverageContinuous(size, A, n_envs=E, weights=g["weights"].tolist(), coarseness=coarse,
                                          starts=g["starts"])
    else:
        env = s.BatchedCoverageDiscretized(size, A, n_envs=E, coarseness=coarse, weights=g["weights"].tolist(),
                                           starts=g["starts"])
    lam = torch.as_tensor(g["lambdas"], dtype=torch.float64, device="cuda")
    buf = env.new_rollout_buffer(T)
    env.reset()
    for t in range(T):
        a = g["actions"][t] if str(g["env"]) == "continuous" else g["actions"][t].astype(np.uint8)
        obs, r, c, d = env.step(a, lambdas=lam, out=(buf, t))
        assert np.array_equal(env.state().cpu().numpy(), g["pos"][t])                      # bit-exact float64
        assert np.array_equal(c.cpu().numpy(), g["cost"][t].astype(np.float32))
        close(r.cpu().numpy(), g["reward"][t], np.abs(g["reward"][t]).max() + 1e-9, rtol=2e-7)
    out = buf.finish(float(g["gamma"]), g["thresholds"])
    scale = np.abs(g["modR"]).max()
    close(out["R"].cpu().numpy(), g["R"], scale)
    close(out["modR"].cpu().numpy(), g["modR"], scale)
    close(out["G"].cpu().numpy(), g["G"], scale)
    close(out["C"].cpu().numpy(), g["C"], np.abs(g["C"]).max())

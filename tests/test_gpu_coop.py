"""Lane-cooperative kernels (one env split over 2 / 4 lanes of a warp, csrc/*_coop.cu) against the
oracle: the parity checks of test_gpu_parity.py re-run with every thread mapping forced through
smarl_set_kernel_variant, over agent counts that do and do not divide by the lane count (padding
agents), ragged env counts and both step-API and fused-rollout kernels.  Same bars as the base tests:
integer / f64 state bit-exact, f32 sums within 1e-5."""
import numpy as np
import pytest
import torch

import test_gpu_parity as tp
from safe_multiagent_rl_b200 import _lib

pytestmark = pytest.mark.gpu

AGENTS = [9, 12, 13, 16, 17, 22, 24, 27, 30, 31, 32]


@pytest.mark.parametrize("lanes", [2, 4])
@pytest.mark.parametrize("A", AGENTS)
def test_collision_coop_step(lanes, A):
    with _lib.kernel_variant(_lib.ENV_COLLISION, lanes):
        tp.test_collision_step_matches_oracle(3 + A % 4, A, 1 + A % 3, 70 + A, 10, 100 + A)


@pytest.mark.parametrize("lanes", [2, 4])
@pytest.mark.parametrize("A", AGENTS)
@pytest.mark.parametrize("g_mode", [0, 1, 2])
def test_collision_coop_rollout(lanes, A, g_mode):
    with _lib.kernel_variant(_lib.ENV_COLLISION, lanes):
        tp.test_collision_fused_rollout(3 + A % 4, A, 1 + A % 3, 70 + A, 12, 200 + A, g_mode)


@pytest.mark.parametrize("lanes", [0, 2, 4])
def test_collision_variants_agree_on_large_fields(lanes):
    """Coordinates up to 200 through the f32 screen, tiny agents (the screen margin scales with the field)."""
    s = tp.smarl()
    size, A, L, E, T = 200, 16, 2, 333, 6
    rng = np.random.default_rng(7)
    starts = rng.random((E, A, 2)) * size
    starts[:, 1::2] = starts[:, 0::2] + rng.normal(0, 0.004, size=(E, A // 2, 2))      # pairs right at the collision distance
    starts = np.clip(starts, 0, size)
    landmarks = rng.random((E, L, 2)) * size
    actions = rng.normal(0, 0.002, size=(T, E, A, 2)).astype(np.float32)
    with _lib.kernel_variant(_lib.ENV_COLLISION, lanes):
        env = s.BatchedCollisionAvoidance(size, A, n_envs=E, n_landmarks=L, starts=starts, landmarks=landmarks,
                                          agents_size=0.002)
        env.reset()
        pos, done = starts.copy(), np.zeros((E, A), dtype=bool)
        total = 0
        for t in range(T):
            _, r, c, d = env.step(actions[t])
            pos, r_o, c_o, done, _ = tp.no.collision_step(pos, done, actions[t].astype(np.float64), landmarks, size,
                                                          agents_size=0.002)
            assert np.array_equal(env.state().cpu().numpy(), pos)
            assert np.array_equal(c.cpu().numpy(), c_o.astype(np.int64))
            assert np.array_equal(r.cpu().numpy(), r_o.astype(np.float32))
            total += int(c_o.sum())
        assert total > 0


# ----------------------------------------------------------------------------- Coverage
@pytest.mark.parametrize("lanes", [2, 4])
@pytest.mark.parametrize("A", AGENTS)
def test_coverage_coop_step(lanes, A):
    with _lib.kernel_variant(_lib.ENV_COVERAGE, lanes):
        tp.test_coverage_step_matches_oracle(2 * A, A, 150 + A, 6, None, 300 + A)
        tp.test_coverage_step_matches_oracle(3, A, 37, 5, 10.0, 400 + A)          # every pair overlaps, ragged E


@pytest.mark.parametrize("lanes", [2, 4])
@pytest.mark.parametrize("A", [12, 17, 32])
def test_coverage_coop_closed_loop_and_lean_buffer(lanes, A):
    with _lib.kernel_variant(_lib.ENV_COVERAGE, lanes):
        tp.test_coverage_rollouts_match_oracle(2 * A, A, 120, 8, None, 500 + A, 1)   # closed loop vs fused vs oracle
        tp.test_coverage_lean_rollout_buffer(2 * A, A, 120, 8, None, 600 + A, 1)     # reward_rows = 1 path


# ----------------------------------------------------------------------------- Congestion
@pytest.mark.parametrize("lanes", [2, 4])
@pytest.mark.parametrize("A", AGENTS)
@pytest.mark.parametrize("mode", ["philox", "recorded"])
def test_congestion_coop_step(lanes, A, mode):
    with _lib.kernel_variant(_lib.ENV_CONGESTION, lanes):
        tp.test_congestion_step_matches_oracle(2 + A % 5, A, 90 + A, 8, 0.3, 700 + A, mode)   # tiny grids: heavy edge sharing


@pytest.mark.parametrize("lanes", [2, 4])
def test_congestion_coop_no_noise_and_closed_loop(lanes):
    with _lib.kernel_variant(_lib.ENV_CONGESTION, lanes):
        tp.test_congestion_step_matches_oracle(3, 16, 64, 10, 0.0, 0, "philox")    # noise 0: mode 0 kernel
        tp.test_congestion_closed_loop_returns()


@pytest.mark.parametrize("A", AGENTS)
@pytest.mark.parametrize("g_mode", [0, 1, 2])
@pytest.mark.parametrize("mode", ["philox", "recorded"])
def test_congestion_coop_rollout(A, g_mode, mode):
    """The lane-cooperative fused rollout (congestion_coop_rollout_kernel) against the oracle: tiny grids (heavy edge
    sharing), ragged env counts, agent counts with and without padding agents."""
    with _lib.kernel_variant(_lib.ENV_CONGESTION, 4):
        tp.test_congestion_fused_rollout(2 + A % 5, A, 90 + A, 9, 0.3, 800 + A, g_mode, mode)


@pytest.mark.parametrize("size,A,E", [(10, 12, 1000), (64, 32, 777), (100, 20, 130), (200, 27, 517)])
def test_congestion_coop_rollout_is_bit_identical_to_the_one_thread_rollout(size, A, E):
    """Same per-(agent, env) accumulation order in both mappings: R, modR, C, G and the final positions agree bit for
    bit; the stats vector (a sum over envs in a different order) within f64 rounding.  Covers all four edge-key widths."""
    s = tp.smarl()
    rng = np.random.default_rng(size + A)
    T = 15
    st = rng.integers(0, size + 1, (E, A, 2)); st[:, 0] = 0
    demand = rng.random((size + 1, size + 1)) * 8 + 2
    g = torch.Generator(device="cuda"); g.manual_seed(A)
    lam = torch.full((1,), 0.3, dtype=torch.float64, device="cuda")
    actions, outs = None, []
    for lanes in (0, 4):
        with _lib.kernel_variant(_lib.ENV_CONGESTION, lanes):
            env = s.BatchedCongestion(size, A, n_envs=E, noise=0.2, starts=st, demand_rate=demand, seed=9)
            if actions is None:
                actions = torch.randint(0, 5, (T, A, env.ld), generator=g, device="cuda", dtype=torch.uint8)
            out = env.rollout(actions, lambdas=lam, gamma=0.97, thresholds=[3.0], g_mode=1)
            outs.append(({k: out[k].clone() for k in ("R", "modR", "C", "G")}, env.state().clone(), out["stats"].vec.clone()))
    a, b = outs
    assert torch.equal(a[1], b[1])
    for k in a[0]:
        assert torch.equal(a[0][k], b[0][k]), k
    assert torch.allclose(a[2], b[2], rtol=1e-12, atol=1e-9)

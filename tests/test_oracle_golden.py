"""The numpy oracle against the committed golden fixtures (reference outputs captured by
tests/golden/make_golden.py).  Runs everywhere, including the GPU box where /root/reference
does not exist.  f64 env outputs are compared bit-exactly."""
import os

import numpy as np
import pytest

from oracle import numpy_oracle as no
from oracle import philox

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def check_accounting(g, rew, cost):
    lam, gamma, thr = g["lambdas"], float(g["gamma"]), g["thresholds"]
    mod = no.modified_reward(rew, cost, lam)
    np.testing.assert_array_equal(no.episode_returns(rew, gamma), g["R"])       # same summation order
    np.testing.assert_allclose(no.episode_returns(mod, gamma), g["modR"], rtol=1e-12)
    np.testing.assert_array_equal(no.episode_cost_sums(cost), g["C"])
    np.testing.assert_allclose(no.reward_to_go(mod, gamma), g["G"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(no.discounted_terms(mod, gamma), g["disc"], rtol=1e-12, atol=1e-12)
    for e in range(rew.shape[1]):
        np.testing.assert_allclose(no.lambda_update(lam, g["C"][e], thr, float(g["meta_lr"])),
                                   g["lam_after"][e], rtol=1e-14)


@pytest.mark.parametrize("name", ["coverage_c1.npz", "coverage_c4.npz"])
def test_coverage_golden(name):
    g = load(name)
    size, A = int(g["size"]), int(g["n_agents"])
    fv = no.coverage_fieldview(size, A)
    assert fv == float(g["fieldview"])
    lut = no.coverage_penalty_lut(size, fv)
    pos = g["starts"].astype(np.int64)
    T = g["actions"].shape[0]
    rew, cost = [], []
    for t in range(T):
        pos, r, c, d = no.coverage_discrete_step(pos, g["actions"][t], size, lut, g["weights"])
        np.testing.assert_array_equal(pos, g["pos"][t])
        np.testing.assert_array_equal(r, g["reward"][t])
        np.testing.assert_array_equal(c, g["cost"][t])
        assert not d.any()
        rew.append(r); cost.append(c)
    check_accounting(g, np.stack(rew), np.stack(cost).astype(np.float64))


@pytest.mark.parametrize("name", ["congestion_paper.npz", "congestion_c3.npz"])
def test_congestion_golden(name):
    g = load(name)
    size, A, noise = int(g["size"]), int(g["n_agents"]), float(g["noise"])
    pos = g["starts"].astype(np.int64)
    T, E = g["actions"].shape[:2]
    ids = np.arange(int(g["env_offset"]), int(g["env_offset"]) + E)
    rew, cost = [], []
    for t in range(T):
        u1, u2 = philox.congestion_uniforms(int(g["philox_seed"]), ids, t, A)
        moves = no.congestion_noise_moves(g["actions"][t], u1, u2, noise)
        pos, r, c, d, con = no.congestion_step(pos, g["actions"][t], moves, size, g["demand"])
        np.testing.assert_array_equal(pos, g["pos"][t])
        np.testing.assert_array_equal(con, g["congestions"][t])
        np.testing.assert_array_equal(r, g["reward"][t])
        np.testing.assert_array_equal(c, g["cost"][t])
        rew.append(r); cost.append(c)
    check_accounting(g, np.stack(rew), np.stack(cost).astype(np.float64))


@pytest.mark.parametrize("name", ["collision_c2.npz", "collision_paper.npz"])
def test_collision_golden(name):
    g = load(name)
    size = int(g["size"])
    pos = g["starts"].copy()
    T, E, A = g["actions"].shape[:3]
    done = np.zeros((E, A), dtype=bool)
    rew, cost = [], []
    for t in range(T):
        pos, r, c, done, active = no.collision_step(pos, done, g["actions"][t].astype(np.float64),
                                                    g["landmarks"], size)
        np.testing.assert_array_equal(active, g["active"][t])
        np.testing.assert_array_equal(pos, g["pos"][t])
        np.testing.assert_array_equal(done, g["done"][t])
        np.testing.assert_array_equal(r, g["reward"][t])
        np.testing.assert_array_equal(c, g["cost"][t])
        rew.append(r); cost.append(c)
    assert (~g["active"]).any() and g["cost"].sum() > 0      # fixtures cover early ends and collisions
    check_accounting(g, np.stack(rew), np.stack(cost))


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = philox.philox4x32_10(*[np.array([v]) for v in c], *k)
        assert tuple(int(x[0]) for x in got) == want


def test_keep_threshold_is_exact():
    rng = np.random.default_rng(0)
    for noise in [0.0, 0.1, 0.25, 0.5, 0.9, 1.0, 1e-9, 0.3]:
        thr = philox.keep_threshold(noise)
        w = np.concatenate([rng.integers(0, 2 ** 32, 1000), [0, 2 ** 32 - 1, max(thr - 1, 0), min(thr, 2 ** 32 - 1)]])
        assert np.array_equal(w * 2.0 ** -32 < 1 - noise, w < thr)


def test_single_word_noise_stream_matches_the_two_draw_model():
    """One Philox word per agent: u1 = w 2^-32 decides keep, int(u2 * 5) = w mod 5 replaces.  The pair
    (u1, u2) fed to congestion.py:64-67 must reproduce the integer rule, and the moves must follow the
    reference's two-draw distribution (keep with prob 1 - noise, else uniform on 0..4)."""
    A, E = 6, 50000
    w = philox.congestion_words(9, np.arange(E), 3, A)
    u1, u2 = philox.congestion_uniforms(9, np.arange(E), 3, A)
    assert w.shape == (E, A) and np.array_equal((u2 * 5).astype(np.int64), w % 5)
    # counter layout: agents 4q..4q+3 share call q, agent a reads output word a & 3
    o = philox.philox4x32_10(np.arange(E), 0, 3, 1, 9, 0)
    assert np.array_equal(w[:, 4], o[0]) and np.array_equal(w[:, 5], o[1])
    actions = np.full((E, A), 7)
    for noise in (0.1, 0.5):
        thr = philox.keep_threshold(noise)
        moves = no.congestion_noise_moves(actions, u1, u2, noise)
        assert np.array_equal(moves, np.where(w < thr, 7, w % 5))
        n = E * A
        replaced = moves != 7
        assert abs(replaced.mean() - noise) < 5 * np.sqrt(noise * (1 - noise) / n)
        counts = np.bincount(moves[replaced], minlength=5)
        expect = replaced.sum() / 5.0
        assert ((counts - expect) ** 2 / expect).sum() < 25.0          # chi-square, 4 dof (p ~ 5e-5)
        # replacement value is independent of how far above the threshold the word landed
        hi = (w >= thr + (2 ** 32 - thr) // 2) & replaced
        c_hi = np.bincount(moves[hi], minlength=5)
        assert ((c_hi - hi.sum() / 5.0) ** 2 / (hi.sum() / 5.0)).sum() < 25.0


@pytest.mark.parametrize("name", ["coverage_continuous_paper.npz", "coverage_discretized.npz"])
def test_coverage_float_golden(name):
    g = load(name)
    size, A, coarse = int(g["size"]), int(g["n_agents"]), int(g["coarseness"])
    fv = no.coverage_fieldview(size, A)
    assert fv == float(g["fieldview"])
    pos = g["starts"].copy()
    T = g["actions"].shape[0]
    rew, cost = [], []
    for t in range(T):
        if str(g["env"]) == "continuous":
            pos, r, c, _ = no.coverage_continuous_step(pos, g["actions"][t].astype(np.float64), size, fv, g["weights"], coarse)
        else:
            pos, r, c, _ = no.coverage_discretized_step(pos, g["actions"][t], size, coarse, fv, g["weights"])
        np.testing.assert_array_equal(pos, g["pos"][t])
        np.testing.assert_array_equal(r, g["reward"][t])
        np.testing.assert_array_equal(c, g["cost"][t])
        rew.append(r); cost.append(c)
    lam, gamma = g["lambdas"], float(g["gamma"])
    mod = no.modified_reward(np.stack(rew), np.stack(cost), lam)
    np.testing.assert_array_equal(no.episode_returns(np.stack(rew), gamma), g["R"])
    np.testing.assert_allclose(no.episode_returns(mod, gamma), g["modR"], rtol=1e-12)
    np.testing.assert_allclose(no.episode_cost_sums(np.stack(cost)), g["C"], rtol=1e-14)
    np.testing.assert_allclose(no.reward_to_go(mod, gamma), g["G"], rtol=1e-12, atol=1e-12)

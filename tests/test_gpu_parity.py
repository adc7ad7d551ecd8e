"""GPU parity: the CUDA path (through the C ABI) against the numpy oracle on the same seeded
inputs.  Bars: integer dynamics (positions, costs, done, congestions via rewards) bit-exact;
Collision float64 positions bit-exact; float32 rewards / returns within 1e-5 relative
(BASELINE.json north_star), written here as rtol=1e-5 with an absolute floor of 1e-5 x the
magnitude scale of the summed terms."""
import numpy as np
import pytest
import torch

from oracle import numpy_oracle as no
from oracle import philox

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def smarl():
    import safe_multiagent_rl_b200 as s
    return s


def close(got, want, scale=None, rtol=RTOL):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    atol = rtol * (np.max(np.abs(want)) if scale is None else scale) + 1e-30
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol)


# ----------------------------------------------------------------------------- Coverage
COVERAGE_CASES = [
    # size, A, E, T, fieldview, seed
    (5, 3, 50, 50, None, 0),          # BASELINE config 1
    (32, 16, 1000, 12, None, 1),      # config 4 shape (fv = 8.0)
    (8, 5, 333, 20, None, 2),         # ragged E (not a multiple of 4 / 16)
    (64, 32, 130, 4, None, 3),        # maximum agents
    (3, 2, 17, 30, 10.0, 4),          # field of view larger than the grid: every pair overlaps
    (5, 1, 40, 5, None, 5),           # single agent: no pairs
    (127, 4, 64, 6, 40.0, 6),         # maximum grid size
]


def coverage_setup(size, A, E, T, fv, seed):
    rng = np.random.default_rng(seed)
    starts = np.floor(rng.random((E, A, 2)) * size).astype(np.int64)
    starts[0] = size - 1 if size > 1 else 0
    starts[-1] = 0
    actions = rng.integers(0, 5, size=(T, E, A))
    weights = (1.0 + (np.arange(A) % 3)).tolist()
    fvv = no.coverage_fieldview(size, A, fv)
    lut = no.coverage_penalty_lut(size, fvv)
    return starts, actions, weights, lut


@pytest.mark.parametrize("size,A,E,T,fv,seed", COVERAGE_CASES)
def test_coverage_step_matches_oracle(size, A, E, T, fv, seed):
    s = smarl()
    starts, actions, weights, lut = coverage_setup(size, A, E, T, fv, seed)
    env = s.BatchedCoverageDiscrete(size, A, n_envs=E, weights=weights, fieldview_size=fv, starts=starts)
    lam = torch.as_tensor(np.linspace(0.1, 0.5, A), dtype=torch.float64, device="cuda")
    obs = env.reset()
    assert np.array_equal(obs.cpu().numpy().reshape(E, A, 2), starts.astype(np.float32))
    pos = starts.copy()
    for t in range(T):
        obs, r, c, d = env.step(actions[t].astype(np.uint8), lambdas=lam)
        pos, r_o, c_o, d_o = no.coverage_discrete_step(pos, actions[t], size, lut, weights)
        assert np.array_equal(env.state().cpu().numpy(), pos)                       # bit-exact
        assert np.array_equal(obs.cpu().numpy().reshape(E, A, 2), pos.astype(np.float32))
        assert np.array_equal(c.cpu().numpy(), c_o)
        assert np.array_equal(d.cpu().numpy().astype(bool), d_o)
        close(r.cpu().numpy(), r_o)
        close(env.penalty[:E].cpu().numpy(), c_o @ lam.cpu().numpy())


@pytest.mark.parametrize("size,A,E,T,fv,seed", COVERAGE_CASES)
@pytest.mark.parametrize("g_mode", [0, 1, 2])
def test_coverage_rollouts_match_oracle(size, A, E, T, fv, seed, g_mode):
    """Closed-loop (T step launches into a RolloutBuffer + returns kernel) and fused open-loop
    rollouts against the oracle's accounting, and against each other."""
    s = smarl()
    starts, actions, weights, lut = coverage_setup(size, A, E, T, fv, seed)
    gamma = 0.999
    lam_np = np.linspace(0.1, 0.5, A)
    thr = np.full(A, 0.4 * T)
    lam = torch.as_tensor(lam_np, dtype=torch.float64, device="cuda")
    env = s.BatchedCoverageDiscrete(size, A, n_envs=E, weights=weights, fieldview_size=fv, starts=starts)

    pos = starts.copy()
    def step_fn(t):
        nonlocal pos
        pos, r, c, _ = no.coverage_discrete_step(pos, actions[t], size, lut, weights)
        return r, c
    want = no.rollout(step_fn, T, gamma, lam_np)
    want_G = {0: None, 1: want["G"], 2: no.discounted_terms(want["mod_reward"], gamma)}[g_mode]
    scale_R = np.max(np.abs(want["modR"])) + 1e-12

    # closed loop
    act_dev = torch.as_tensor(actions.astype(np.uint8), device="cuda")
    out = env.rollout_closed_loop(lambda obs, t: act_dev[t], T, lam, gamma, thresholds=thr, g_mode=g_mode)
    assert np.array_equal(env.state().cpu().numpy(), pos)
    assert np.array_equal(out["C"].cpu().numpy(), want["C"])
    close(out["R"].cpu().numpy(), want["R"], scale_R)
    close(out["modR"].cpu().numpy(), want["modR"], scale_R)
    close(out["buffer"].modified_rewards().cpu().numpy(), want["mod_reward"])
    if g_mode:
        close(out["G"].cpu().numpy(), want_G, scale_R)
    st = out["stats"]
    assert np.array_equal(st.cost_sum.cpu().numpy(), want["C"].sum(0))               # exact integers
    assert np.array_equal(st.violations.cpu().numpy(), (want["C"] > thr[None]).sum(0))
    close(st.return_sum.cpu().numpy(), want["R"].sum(0), scale_R * E)
    close(st.modified_return_sum.cpu().numpy(), want["modR"].sum(0), scale_R * E)
    assert float(st.count) == E
    closed = {k: out[k].cpu().numpy().copy() for k in ("R", "modR", "C")}

    # fused open loop
    act_k = torch.zeros(T, A, env.ld, dtype=torch.uint8, device="cuda")
    act_k[:, :, :E] = act_dev.permute(0, 2, 1)
    fo = env.rollout(act_k, lambdas=lam, gamma=gamma, thresholds=thr, g_mode=g_mode)
    assert np.array_equal(env.state().cpu().numpy(), pos)
    assert np.array_equal(fo["C"].cpu().numpy(), want["C"])
    close(fo["R"].cpu().numpy(), want["R"], scale_R)
    close(fo["modR"].cpu().numpy(), want["modR"], scale_R)
    if g_mode:
        close(fo["G"].cpu().numpy(), want_G, scale_R)
    st = fo["stats"]
    assert np.array_equal(st.cost_sum.cpu().numpy(), want["C"].sum(0))
    assert np.array_equal(st.violations.cpu().numpy(), (want["C"] > thr[None]).sum(0))
    close(st.return_sum.cpu().numpy(), want["R"].sum(0), scale_R * E)
    close(st.modified_return_sum.cpu().numpy(), want["modR"].sum(0), scale_R * E)
    assert float(st.count) == E
    close(fo["R"].cpu().numpy(), closed["R"], scale_R, rtol=2e-6)                    # two modes agree
    assert np.array_equal(fo["C"].cpu().numpy(), closed["C"])


# ----------------------------------------------------------------------------- Congestion
CONGESTION_CASES = [
    # size, A, E, T, noise, seed
    (3, 3, 64, 10, 0.0, 0),           # paper config, no noise
    (3, 8, 200, 30, 0.1, 1),
    (10, 8, 515, 25, 0.1, 2),         # config 3 shape, ragged E
    (2, 6, 33, 40, 0.5, 3),           # tiny grid: heavy edge sharing
    (5, 32, 40, 6, 0.3, 4),           # maximum agents
    (1, 5, 21, 20, 1.0, 5),           # all moves replaced by noise
]


def congestion_setup(size, A, E, T, seed):
    rng = np.random.default_rng(seed)
    demand = rng.random((size + 1, size + 1)) * 8 + 2
    if size == 3:
        demand = np.array([[2, 2, 4, 4], [3, 6, 10, 5], [3, 8, 3, 4], [4, 6, 7, 8]], dtype=np.float64)
    starts = np.floor(rng.random((E, A, 2)) * size).astype(np.int64)
    starts[:, 0] = 0
    actions = rng.integers(0, 5, size=(T, E, A))
    return demand, starts, actions


@pytest.mark.parametrize("size,A,E,T,noise,seed", CONGESTION_CASES)
@pytest.mark.parametrize("mode", ["philox", "recorded"])
def test_congestion_step_matches_oracle(size, A, E, T, noise, seed, mode):
    s = smarl()
    demand, starts, actions = congestion_setup(size, A, E, T, seed)
    offset = 1000 * seed + 7
    env = s.BatchedCongestion(size, A, n_envs=E, noise=noise, starts=starts, demand_rate=demand,
                              seed=12345 + seed, env_offset=offset)
    lam = torch.as_tensor([0.7], dtype=torch.float64, device="cuda")
    env.reset()
    pos = starts.copy()
    ids = np.arange(offset, offset + E)
    for t in range(T):
        u1, u2 = philox.congestion_uniforms(12345 + seed, ids, t, A)
        moves = no.congestion_noise_moves(actions[t], u1, u2, noise)
        if mode == "recorded":
            obs, r, c, d = env.step(actions[t].astype(np.uint8), lambdas=lam, moves=moves.astype(np.uint8))
        else:
            obs, r, c, d = env.step(actions[t].astype(np.uint8), lambdas=lam)
            assert np.array_equal(env.moves[:, :E].t().cpu().numpy(), moves)        # the Philox stream itself
        pos, r_o, c_o, d_o, con = no.congestion_step(pos, actions[t], moves, size, demand)
        assert np.array_equal(env.state().cpu().numpy(), pos)
        assert np.array_equal(obs.cpu().numpy().reshape(E, A, 2), pos.astype(np.float32))
        assert np.array_equal(c.cpu().numpy(), c_o)
        assert not d.any()
        # rewards are computed in f64 in the reference's operation order and rounded once to f32
        assert np.array_equal(r.cpu().numpy(), r_o.astype(np.float32))
        close(env.penalty[:E].cpu().numpy(), 0.7 * c_o[:, 0])


def test_congestion_closed_loop_returns():
    s = smarl()
    size, A, E, T, noise, seed = 10, 8, 300, 100, 0.1, 9
    demand, starts, actions = congestion_setup(size, A, E, T, seed)
    gamma, lam_np, thr = 0.9, np.array([0.35]), np.array([1.5])
    env = s.BatchedCongestion(size, A, n_envs=E, noise=noise, starts=starts, demand_rate=demand, seed=3)
    pos = starts.copy()
    def step_fn(t):
        nonlocal pos
        u1, u2 = philox.congestion_uniforms(3, np.arange(E), t, A)
        moves = no.congestion_noise_moves(actions[t], u1, u2, noise)
        pos, r, c, _, _ = no.congestion_step(pos, actions[t], moves, size, demand)
        return r, c
    want = no.rollout(step_fn, T, gamma, lam_np)
    act_dev = torch.as_tensor(actions.astype(np.uint8), device="cuda")
    lam = torch.as_tensor(lam_np, device="cuda")
    out = env.rollout_closed_loop(lambda obs, t: act_dev[t], T, lam, gamma, thresholds=thr)
    scale = np.max(np.abs(want["modR"]))
    assert np.array_equal(out["C"].cpu().numpy(), want["C"])
    close(out["R"].cpu().numpy(), want["R"], scale)
    close(out["modR"].cpu().numpy(), want["modR"], scale)
    close(out["G"].cpu().numpy(), want["G"], scale)
    assert np.array_equal(out["stats"].cost_sum.cpu().numpy(), want["C"].sum(0))


# ----------------------------------------------------------------------------- Collision
COLLISION_CASES = [
    # size, A, L, E, T, seed
    (5, 3, 1, 1000, 50, 0),           # config 2 shape
    (2, 5, 1, 257, 20, 1),            # paper config
    (5, 8, 3, 100, 30, 2),
    (3, 12, 2, 64, 25, 3),
    (4, 32, 1, 48, 6, 4),             # maximum agents
    (64, 16, 1, 300, 12, 5),          # large coordinates through the f32-screened pair loop (A >= 12)
    (200, 13, 2, 200, 8, 6),
]


def collision_setup(size, A, L, E, T, seed):
    rng = np.random.default_rng(seed)
    starts = rng.random((E, A, 2)) * size
    landmarks = rng.random((E, L, 2)) * size
    actions = rng.normal(0, 0.5, size=(T, E, A, 2)).astype(np.float32)
    q = E // 4
    # a quarter of the envs: every agent walks straight to landmark 0 (agents finish, episodes end early)
    actions[:, :q] = ((landmarks[:q, :1] - starts[:q]) / 6).astype(np.float32)[None]
    # another quarter: agents crowded around the landmark (collisions)
    starts[q:2 * q] = np.clip(landmarks[q:2 * q, :1] + rng.normal(0, 0.4, size=(q, A, 2)), 0, size)
    return starts, landmarks, actions


@pytest.mark.parametrize("size,A,L,E,T,seed", COLLISION_CASES)
def test_collision_step_matches_oracle(size, A, L, E, T, seed):
    s = smarl()
    starts, landmarks, actions = collision_setup(size, A, L, E, T, seed)
    env = s.BatchedCollisionAvoidance(size, A, n_envs=E, n_landmarks=L, starts=starts, landmarks=landmarks)
    lam = torch.as_tensor([0.5], dtype=torch.float64, device="cuda")
    env.reset()
    pos, done = starts.copy(), np.zeros((E, A), dtype=bool)
    n_done_envs = 0
    n_coll = 0
    for t in range(T):
        obs, r, c, d = env.step(actions[t], lambdas=lam)
        pos, r_o, c_o, done, active = no.collision_step(pos, done, actions[t].astype(np.float64), landmarks, size)
        assert np.array_equal(env.state().cpu().numpy(), pos), t                      # bit-exact float64
        assert np.array_equal(d.cpu().numpy().astype(bool), done)
        assert np.array_equal(c.cpu().numpy(), c_o.astype(np.int64))
        assert np.array_equal(r.cpu().numpy(), r_o.astype(np.float32))                # f64-exact, rounded once
        close(env.penalty[:E].cpu().numpy(), 0.5 * c_o[:, 0])
        n_coll += int(c_o.sum())
        n_done_envs = int((~active).sum())
    assert n_coll > 0, "test inputs should provoke collisions"
    if A <= 12:
        assert n_done_envs > 0, "test inputs should end some episodes early"


def test_collision_closed_loop_returns():
    s = smarl()
    size, A, L, E, T, seed = 5, 3, 1, 600, 50, 11
    starts, landmarks, actions = collision_setup(size, A, L, E, T, seed)
    gamma, lam_np, thr = 0.99, np.array([0.5]), np.array([1.0])
    env = s.BatchedCollisionAvoidance(size, A, n_envs=E, n_landmarks=L, starts=starts, landmarks=landmarks)
    pos, done = starts.copy(), np.zeros((E, A), dtype=bool)
    def step_fn(t):
        nonlocal pos, done
        pos, r, c, done, _ = no.collision_step(pos, done, actions[t].astype(np.float64), landmarks, size)
        return r, c
    want = no.rollout(step_fn, T, gamma, lam_np)
    act_dev = torch.as_tensor(actions, device="cuda")
    lam = torch.as_tensor(lam_np, device="cuda")
    out = env.rollout_closed_loop(lambda obs, t: act_dev[t], T, lam, gamma, thresholds=thr)
    scale = np.max(np.abs(want["modR"]))
    assert np.array_equal(out["C"].cpu().numpy(), want["C"].astype(np.int64))
    close(out["R"].cpu().numpy(), want["R"], scale)
    close(out["modR"].cpu().numpy(), want["modR"], scale)
    close(out["G"].cpu().numpy(), want["G"], scale)


# ----------------------------------------------------------------------------- lambda update
def test_lambda_update_and_meta_agent():
    s = smarl()
    A = K = 3
    meta = s.BatchedMetaAgent([1] * K, 0.999, 0.05, [25, 25, 25], start_learning_cycle=0, lambda_0=0.2,
                              n_agents=A)
    lib_len = 2 * K + 2 * A + 1
    vec = torch.zeros(lib_len, dtype=torch.float64, device="cuda")
    vec[:K] = torch.tensor([3000.0, 2000.0, 100.0])
    vec[-1] = 100.0
    meta.step(vec)
    meta.step(vec)
    meta.update()
    want = no.lambda_update([0.2] * 3, np.array([30.0, 20.0, 1.0]), [25, 25, 25], 0.05)
    np.testing.assert_allclose(meta.lambdas.cpu().numpy(), want, rtol=1e-15)
    assert meta.learning_cycle == 0


# ----------------------------------------------------------------------------- fused rollouts (Congestion, Collision)
def kernel_layout(a, ld):
    """[T, E, rows] numpy -> [T, rows, ld] device tensor."""
    T, E, R = a.shape
    out = torch.zeros(T, R, ld, dtype=torch.as_tensor(a).dtype, device="cuda")
    out[:, :, :E] = torch.as_tensor(a, device="cuda").permute(0, 2, 1)
    return out


def check_products(out, want, want_G, thr, E, scale, exact_C=True):
    assert np.array_equal(out["C"].cpu().numpy(), want["C"].astype(np.int64))
    close(out["R"].cpu().numpy(), want["R"], scale)
    close(out["modR"].cpu().numpy(), want["modR"], scale)
    if want_G is not None:
        close(out["G"].cpu().numpy(), want_G, scale)
    st = out["stats"]
    assert np.array_equal(st.cost_sum.cpu().numpy(), want["C"].sum(0))
    assert np.array_equal(st.violations.cpu().numpy(), (want["C"] > np.asarray(thr)[None]).sum(0))
    close(st.return_sum.cpu().numpy(), want["R"].sum(0), scale * E)
    close(st.modified_return_sum.cpu().numpy(), want["modR"].sum(0), scale * E)
    assert float(st.count) == E


@pytest.mark.parametrize("size,A,E,T,noise,seed", CONGESTION_CASES + [(10, 8, 700, 100, 0.1, 6), (4, 16, 90, 12, 0.2, 7)])
@pytest.mark.parametrize("g_mode", [0, 1, 2])
@pytest.mark.parametrize("mode", ["philox", "recorded"])
def test_congestion_fused_rollout(size, A, E, T, noise, seed, g_mode, mode):
    s = smarl()
    demand, starts, actions = congestion_setup(size, A, E, T, seed)
    gamma, lam_np, thr = 0.9, np.array([0.35]), np.array([0.3 * T])
    offset = 77 + seed
    env = s.BatchedCongestion(size, A, n_envs=E, noise=noise, starts=starts, demand_rate=demand, seed=5 + seed,
                              env_offset=offset)
    ids = np.arange(offset, offset + E)
    pos = starts.copy()
    all_moves = []

    def step_fn(t):
        nonlocal pos
        u1, u2 = philox.congestion_uniforms(5 + seed, ids, t, A)
        moves = no.congestion_noise_moves(actions[t], u1, u2, noise)
        all_moves.append(moves)
        pos, r, c, _, _ = no.congestion_step(pos, actions[t], moves, size, demand)
        return r.astype(np.float32).astype(np.float64), c          # rewards are published as f32
    want = no.rollout(step_fn, T, gamma, lam_np)
    want_G = {0: None, 1: want["G"], 2: no.discounted_terms(want["mod_reward"], gamma)}[g_mode]
    scale = np.max(np.abs(want["modR"]))
    lam = torch.as_tensor(lam_np, device="cuda")
    act_k = kernel_layout(actions.astype(np.uint8), env.ld)
    mv_k = kernel_layout(np.stack(all_moves).astype(np.uint8), env.ld) if mode == "recorded" else None
    out = env.rollout(act_k, lambdas=lam, gamma=gamma, thresholds=thr, g_mode=g_mode, moves=mv_k)
    assert np.array_equal(env.state().cpu().numpy(), pos)
    check_products(out, want, want_G, thr, E, scale)


@pytest.mark.parametrize("size,A,L,E,T,seed", COLLISION_CASES)
@pytest.mark.parametrize("g_mode", [0, 1, 2])
def test_collision_fused_rollout(size, A, L, E, T, seed, g_mode):
    s = smarl()
    starts, landmarks, actions = collision_setup(size, A, L, E, T, seed)
    gamma, lam_np, thr = 0.99, np.array([0.5]), np.array([1.0])
    env = s.BatchedCollisionAvoidance(size, A, n_envs=E, n_landmarks=L, starts=starts, landmarks=landmarks)
    pos, done = starts.copy(), np.zeros((E, A), dtype=bool)
    n_active = np.zeros(E, dtype=np.int64)

    def step_fn(t):
        nonlocal pos, done, n_active
        pos, r, c, done, active = no.collision_step(pos, done, actions[t].astype(np.float64), landmarks, size)
        n_active += active
        return r.astype(np.float32).astype(np.float64), c
    want = no.rollout(step_fn, T, gamma, lam_np)
    want_G = {0: None, 1: want["G"], 2: no.discounted_terms(want["mod_reward"], gamma)}[g_mode]
    scale = np.max(np.abs(want["modR"]))
    lam = torch.as_tensor(lam_np, device="cuda")
    act_k = kernel_layout(actions.reshape(T, E, 2 * A), env.ld)
    out = env.rollout(act_k, lambdas=lam, gamma=gamma, thresholds=thr, g_mode=g_mode)
    assert np.array_equal(env.state().cpu().numpy(), pos)                           # bit-exact f64 final positions
    assert np.array_equal(env.agent_done[:, :E].t().cpu().numpy().astype(bool), done)
    assert np.array_equal(out["n_active"].cpu().numpy(), n_active)
    check_products(out, want, want_G, thr, E, scale)


def test_step_accepts_kernel_layout_and_env_major_actions():
    """env.step takes actions as [E, A] (reference orientation) or as the [A, ld] kernel layout
    (zero-copy); both must give the same transition."""
    s = smarl()
    rng = np.random.default_rng(0)
    E, A, S = 100, 5, 9
    starts = rng.integers(0, S, size=(E, A, 2))
    act = rng.integers(0, 5, size=(E, A)).astype(np.uint8)
    e1 = s.BatchedCoverageDiscrete(S, A, n_envs=E, starts=starts)
    e2 = s.BatchedCoverageDiscrete(S, A, n_envs=E, starts=starts)
    e3 = s.BatchedCoverageDiscrete(S, A, n_envs=E, starts=starts)
    for e in (e1, e2, e3):
        e.reset()
    e1.step(act)
    k = torch.zeros(A, e2.ld, dtype=torch.uint8, device="cuda")
    k[:, :E] = torch.as_tensor(act, device="cuda").t()
    e2.step(k)                                           # auto-detected kernel layout
    e3.action_buffer.copy_(k)
    e3.step(e3.action_buffer, agent_major=True)          # policies writing the env's own buffer
    assert torch.equal(e1.state(), e2.state()) and torch.equal(e1.state(), e3.state())
    assert torch.equal(e1.reward, e2.reward) and torch.equal(e1.reward, e3.reward)


# ----------------------------------------------------------------------------- float-position Coverage variants
@pytest.mark.parametrize("size,A,E,T,coarse,seed", [(5, 3, 300, 40, None, 0), (5, 3, 300, 40, 6, 1), (10, 8, 130, 15, 20, 2),
                                                     (3, 32, 33, 6, 2, 3)])
def test_coverage_continuous_matches_oracle(size, A, E, T, coarse, seed):
    s = smarl()
    rng = np.random.default_rng(seed)
    w = (1.0 + (np.arange(A) % 3)).tolist()
    starts = rng.random((E, A, 2)) * size
    actions = rng.normal(0, 0.8, size=(T, E, A, 2)).astype(np.float32)
    env = s.BatchedCoverageContinuous(size, A, n_envs=E, weights=w, coarseness=coarse, starts=starts)
    fv = no.coverage_fieldview(size, A)
    assert env.fieldview_size == fv
    lam_np = np.linspace(0.1, 0.3, A)
    lam = torch.as_tensor(lam_np, device="cuda")
    gamma = 0.99
    buf = env.new_rollout_buffer(T)
    env.reset()
    pos = starts.copy()
    rs, cs = [], []
    for t in range(T):
        obs, r, c, d = env.step(actions[t], lambdas=lam, out=(buf, t))
        pos, r_o, c_o, _ = no.coverage_continuous_step(pos, actions[t].astype(np.float64), size, fv, w, coarse, exact_pow=False)
        assert np.array_equal(env.state().cpu().numpy(), pos)                        # bit-exact float64
        assert np.array_equal(r.cpu().numpy(), r_o.astype(np.float32))               # f64-exact, rounded once
        assert np.array_equal(c.cpu().numpy(), c_o.astype(np.float32))
        assert not d.any()
        close(env.obs[:, :E].t().cpu().numpy().reshape(E, A, 2), pos)
        rs.append(r_o.astype(np.float32).astype(np.float64)); cs.append(c_o)
    want = no.rollout(lambda t: (rs[t], cs[t]), T, gamma, lam_np)
    out = buf.finish(gamma, thresholds=[0.5 * T] * A)
    scale = np.abs(want["modR"]).max()
    close(out["R"].cpu().numpy(), want["R"], scale)
    close(out["modR"].cpu().numpy(), want["modR"], scale)
    close(out["G"].cpu().numpy(), want["G"], scale)
    close(out["C"].cpu().numpy(), want["C"], np.abs(want["C"]).max())
    close(out["stats"].cost_sum.cpu().numpy(), want["C"].sum(0), np.abs(want["C"].sum(0)).max())
    assert np.array_equal(out["stats"].violations.cpu().numpy(), (out["C"].cpu().numpy() > 0.5 * T).sum(0))


@pytest.mark.parametrize("size,A,E,T,coarse,seed", [(5, 3, 300, 40, 20, 0), (5, 3, 100, 40, 6, 1), (10, 8, 130, 15, 7, 2),
                                                     (3, 32, 33, 6, 3, 3)])
def test_coverage_discretized_matches_oracle(size, A, E, T, coarse, seed):
    s = smarl()
    rng = np.random.default_rng(seed)
    w = (1.0 + (np.arange(A) % 3)).tolist()
    zoom = coarse / size
    starts = np.floor(rng.random((E, A, 2)) * size * zoom) / zoom
    actions = rng.integers(0, 9, size=(T, E, A))
    env = s.BatchedCoverageDiscretized(size, A, n_envs=E, coarseness=coarse, weights=w, starts=starts)
    fv = no.coverage_fieldview(size, A)
    env.reset()
    pos = starts.copy()
    for t in range(T):
        obs, r, c, d = env.step(actions[t].astype(np.uint8))
        pos, r_o, c_o, _ = no.coverage_discretized_step(pos, actions[t], size, coarse, fv, w, exact_pow=False)
        assert np.array_equal(env.state().cpu().numpy(), pos)
        assert np.array_equal(r.cpu().numpy(), r_o.astype(np.float32))
        assert np.array_equal(c.cpu().numpy(), c_o.astype(np.float32))


# ----------------------------------------------------------------------------- PPO standardised returns
def test_ppo_standardised_returns_coverage_and_collision():
    """g_mode 3 of the returns kernel (PPOAgent.step, agent.py:276-281), incl. early-ending
    Collision episodes whose length comes from the env's episode_len counter."""
    s = smarl()
    # Coverage: every episode runs T steps
    size, A, E, T, fv, seed = COVERAGE_CASES[0]
    starts, actions, weights, lut = coverage_setup(size, A, E, T, fv, seed)
    env = s.BatchedCoverageDiscrete(size, A, n_envs=E, weights=weights, starts=starts)
    lam_np = np.array([0.1, 0.2, 0.3])
    pos = starts.copy()
    def step_fn(t):
        nonlocal pos
        pos, r, c, _ = no.coverage_discrete_step(pos, actions[t], size, lut, weights)
        return r, c
    want = no.rollout(step_fn, T, 0.999, lam_np)
    act = torch.as_tensor(actions.astype(np.uint8), device="cuda")
    out = env.rollout_closed_loop(lambda obs, t: act[t], T, torch.as_tensor(lam_np, device="cuda"), 0.999,
                                  g_mode=s.G_PPO_STANDARDISED)
    close(out["G"].cpu().numpy(), no.ppo_standardised_returns(want["mod_reward"], 0.999), 1.0, rtol=2e-5)
    close(out["modR"].cpu().numpy(), want["modR"], np.abs(want["modR"]).max())
    # Collision: a quarter of the episodes end early
    size, A, L, E, T, seed = COLLISION_CASES[0]
    starts, landmarks, actions = collision_setup(size, A, L, E, T, seed)
    env = s.BatchedCollisionAvoidance(size, A, n_envs=E, n_landmarks=L, starts=starts, landmarks=landmarks)
    pos, done = starts.copy(), np.zeros((E, A), dtype=bool)
    n_active = np.zeros(E, dtype=np.int64)
    def step_fn2(t):
        nonlocal pos, done, n_active
        pos, r, c, done, active = no.collision_step(pos, done, actions[t].astype(np.float64), landmarks, size)
        n_active += active
        return r.astype(np.float32).astype(np.float64), c
    want = no.rollout(step_fn2, T, 0.99, [0.5])
    act = torch.as_tensor(actions, device="cuda")
    out = env.rollout_closed_loop(lambda obs, t: act[t], T, torch.as_tensor([0.5], dtype=torch.float64, device="cuda"),
                                  0.99, g_mode=s.G_PPO_STANDARDISED)
    assert np.array_equal(env.episode_len[:E].cpu().numpy(), n_active) and (n_active < T).any()
    ok = n_active >= 2
    got = out["G"].cpu().numpy()
    ref = no.ppo_standardised_returns(want["mod_reward"], 0.99, n_active)
    close(got[:, ok], ref[:, ok], 1.0, rtol=2e-5)
    assert np.isnan(got[0, n_active == 1]).all()            # torch: the std of a single sample is nan


def test_collision_normalize_state_and_shuffle_obs():
    """normalize_state=True: the returned state is state / size (collision_avoidance.py:87-88,:146-147,
    :164-165) while the dynamics keep running on the raw positions; with shuffle the landmark rows are
    normalised too."""
    s = smarl()
    rng = np.random.default_rng(3)
    E, A, size = 77, 4, 5
    starts, lm = rng.random((E, A, 2)) * size, rng.random((E, 1, 2)) * size
    act = rng.normal(0, 0.5, (E, A, 2)).astype(np.float32)
    plain = s.BatchedCollisionAvoidance(size, A, n_envs=E, starts=starts, landmarks=lm)
    norm = s.BatchedCollisionAvoidance(size, A, n_envs=E, starts=starts, landmarks=lm, normalize_state=True)
    o0 = norm.reset(); plain.reset()
    assert np.array_equal(o0.cpu().numpy().reshape(E, A, 2), (starts / size).astype(np.float32))
    o1, r1, c1, d1 = norm.step(act)
    o2, r2, c2, d2 = plain.step(act)
    assert torch.equal(norm.state(), plain.state()) and torch.equal(r1, r2) and torch.equal(c1, c2)
    assert np.array_equal(o1.cpu().numpy().reshape(E, A, 2), (plain.state().cpu().numpy() / size).astype(np.float32))
    sh = s.BatchedCollisionAvoidance(size, A, n_envs=E, n_landmarks=2, shuffle=True, normalize_state=True, seed=1)
    o = sh.reset().cpu().numpy()
    want = np.concatenate([sh.state().cpu().numpy().reshape(E, 2 * A),
                           sh.landmarks[:, :E].t().cpu().numpy()], axis=1) / size
    assert np.array_equal(o, want.astype(np.float32))


# ----------------------------------------------------------------------------- lean rollout buffer
@pytest.mark.parametrize("size,A,E,T,fv,seed", COVERAGE_CASES[:4])
@pytest.mark.parametrize("g_mode", [0, 1, 2])
def test_coverage_lean_rollout_buffer(size, A, E, T, fv, seed, g_mode):
    """lean=True stores one env-reward row (weights applied in the accounting kernel) and no done flags;
    the episode products must match the oracle and the default buffer."""
    s = smarl()
    starts, actions, weights, lut = coverage_setup(size, A, E, T, fv, seed)
    gamma, lam_np, thr = 0.999, np.linspace(0.1, 0.5, A), np.full(A, 0.4 * T)
    lam = torch.as_tensor(lam_np, dtype=torch.float64, device="cuda")
    env = s.BatchedCoverageDiscrete(size, A, n_envs=E, weights=weights, fieldview_size=fv, starts=starts)
    pos = starts.copy()
    def step_fn(t):
        nonlocal pos
        pos, r, c, _ = no.coverage_discrete_step(pos, actions[t], size, lut, weights)
        return r, c
    want = no.rollout(step_fn, T, gamma, lam_np)
    want_G = {0: None, 1: want["G"], 2: no.discounted_terms(want["mod_reward"], gamma)}[g_mode]
    scale = np.max(np.abs(want["modR"])) + 1e-12
    act = torch.as_tensor(actions.astype(np.uint8), device="cuda")
    lean = env.new_rollout_buffer(T, g_mode=max(g_mode, 1), lean=True)
    assert lean.reward.shape[1] == 1 and lean.done is None
    out = env.rollout_closed_loop(lambda obs, t: act[t], T, lam, gamma, thresholds=thr, buffer=lean, g_mode=g_mode)
    check_products(out, want, want_G, thr, E, scale)
    close(lean.rewards().cpu().numpy(), want["reward"], np.abs(want["reward"]).max() + 1e-12)
    obs, r, c, d = env.step(act[0], out=(lean, 0))
    assert r.shape == (E, 1) and d.shape == (E, A) and not d.any()


def test_coverage_lean_rollout_buffer_ppo_standardised():
    """g_mode 3 through the shared-reward accounting kernel: the standardised reward-to-go of every agent
    (agent.py:276-281) from one Horner pair per env and five running sums."""
    s = smarl()
    size, A, E, T, fv, seed = COVERAGE_CASES[1]
    starts, actions, weights, lut = coverage_setup(size, A, E, T, fv, seed)
    gamma, lam_np = 0.999, np.linspace(0.1, 0.5, A)
    env = s.BatchedCoverageDiscrete(size, A, n_envs=E, weights=weights, fieldview_size=fv, starts=starts)
    pos = starts.copy()
    def step_fn(t):
        nonlocal pos
        pos, r, c, _ = no.coverage_discrete_step(pos, actions[t], size, lut, weights)
        return r, c
    want = no.rollout(step_fn, T, gamma, lam_np)
    act = torch.as_tensor(actions.astype(np.uint8), device="cuda")
    lean = env.new_rollout_buffer(T, g_mode=s.G_PPO_STANDARDISED, lean=True)
    out = env.rollout_closed_loop(lambda obs, t: act[t], T, torch.as_tensor(lam_np, device="cuda"), gamma,
                                  buffer=lean, g_mode=s.G_PPO_STANDARDISED)
    close(out["G"].cpu().numpy(), no.ppo_standardised_returns(want["mod_reward"], gamma), 1.0, rtol=2e-5)
    close(out["modR"].cpu().numpy(), want["modR"], np.abs(want["modR"]).max())
    close(out["R"].cpu().numpy(), want["R"], np.abs(want["R"]).max())


# ----------------------------------------------------------------------------- drop-in protocol (lists in / lists out)
def test_single_env_adapter_follows_the_reference_protocol():
    """SingleEnvAdapter(n_envs=1) behaves like a reference env object for the unmodified driver loop
    (main.py:28-57): lists in, lists out, same values as the one-env CPU port."""
    from oracle import scalar_port as sp
    s = smarl()
    rng = np.random.default_rng(0)
    size, A, T = 5, 3, 25
    starts = rng.integers(0, size, (A, 2))
    port = sp.CoveragePort(size, A, starts, weights=[1.0, 2.0, 3.0])
    env = s.SingleEnvAdapter(s.BatchedCoverageDiscrete(size, A, n_envs=1, weights=[1.0, 2.0, 3.0], starts=starts[None]))
    assert env.action_space == 5 and env.state_space == 2 * A and env.constraint_space == [1] * A   # attrs main.py reads
    st, st_p = env.reset(), port.reset()
    assert st == st_p and isinstance(st, list) and isinstance(st[0], list)
    for t in range(T):
        actions = [int(a) for a in rng.integers(0, 5, A)]
        st, r, c, d = env.step(actions)
        st_p, r_p, c_p, d_p = port.step(actions)
        assert st == st_p and c == c_p and d == d_p
        np.testing.assert_allclose(r, r_p, rtol=1e-6)
    # continuous actions arrive as nested [[dx, dy]] lists (agent.py:124-125)
    lm, cs = rng.random((1, 2)) * 5, rng.random((A, 2)) * 5
    cport = sp.CollisionPort(5, A, cs, lm)
    cenv = s.SingleEnvAdapter(s.BatchedCollisionAvoidance(5, A, n_envs=1, starts=cs[None], landmarks=lm[None]))
    assert cenv.reset() == cport.reset()
    for t in range(T):
        act = rng.normal(0, 0.5, (A, 1, 2)).astype(np.float32).astype(np.float64).tolist()
        st, r, c, d = cenv.step(act)
        st_p, r_p, c_p, d_p = cport.step(act)
        assert st == st_p and d == d_p and c == [float(c_p[0])]
        np.testing.assert_allclose(r, r_p, rtol=1e-6)


def test_collision_lean_rollout_buffer():
    s = smarl()
    size, A, L, E, T, seed = COLLISION_CASES[0]
    starts, landmarks, actions = collision_setup(size, A, L, E, T, seed)
    gamma, lam_np, thr = 0.99, np.array([0.5]), np.array([1.0])
    env = s.BatchedCollisionAvoidance(size, A, n_envs=E, n_landmarks=L, starts=starts, landmarks=landmarks)
    pos, done = starts.copy(), np.zeros((E, A), dtype=bool)
    def step_fn(t):
        nonlocal pos, done
        pos, r, c, done, _ = no.collision_step(pos, done, actions[t].astype(np.float64), landmarks, size)
        return r.astype(np.float32).astype(np.float64), c
    want = no.rollout(step_fn, T, gamma, lam_np)
    act = torch.as_tensor(actions, device="cuda")
    lean = env.new_rollout_buffer(T, lean=True)
    assert lean.reward.shape[1] == 1 and lean.done is not None          # agents do finish here: done flags are kept
    out = env.rollout_closed_loop(lambda obs, t: act[t], T, torch.as_tensor(lam_np, device="cuda"), gamma,
                                  thresholds=thr, buffer=lean)
    check_products(out, want, want["G"], thr, E, np.abs(want["modR"]).max())
    assert np.array_equal(lean.done[T - 1][:, :E].t().cpu().numpy().astype(bool), done)


# ----------------------------------------------------------------------------- edge cases
@pytest.mark.parametrize("E", [1, 2, 3, 4, 5, 15, 16, 17])
def test_tiny_batches_all_envs(E):
    """n_envs around the 4-env thread group and the 16-env padding, one-step episodes, one agent,
    gamma = 1, no lambdas / thresholds."""
    s = smarl()
    rng = np.random.default_rng(E)
    # Coverage, single agent: no pairs, reward 0, cost = moved
    env = s.BatchedCoverageDiscrete(4, 1, n_envs=E, starts=rng.integers(0, 4, (E, 1, 2)))
    act = rng.integers(0, 5, (1, E, 1)).astype(np.uint8)
    out = env.rollout_closed_loop(lambda obs, t: torch.as_tensor(act[t], device="cuda"), 1, None, 1.0)
    assert np.array_equal(out["C"].cpu().numpy(), (act[0] != 4).astype(np.int64)) and float(out["R"].abs().max()) == 0.0
    assert float(out["stats"].count) == E
    # Congestion with two agents on a 1x1 grid, T = 3, gamma = 1
    st = np.zeros((E, 2, 2), dtype=np.int64)
    dem = np.array([[2.0, 3.0], [4.0, 5.0]])
    cenv = s.BatchedCongestion(1, 2, n_envs=E, noise=0.0, starts=st, demand_rate=dem)
    acts = rng.integers(0, 5, (3, E, 2))
    pos = st.copy()
    def step_fn(t):
        nonlocal pos
        pos, r, c, _, _ = no.congestion_step(pos, acts[t], acts[t], 1, dem)
        return r.astype(np.float32).astype(np.float64), c
    want = no.rollout(step_fn, 3, 1.0, [0.0])
    a_dev = torch.as_tensor(acts.astype(np.uint8), device="cuda")
    o = cenv.rollout_closed_loop(lambda obs, t: a_dev[t], 3, None, 1.0)
    assert np.array_equal(cenv.state().cpu().numpy(), pos) and np.array_equal(o["C"].cpu().numpy(), want["C"])
    close(o["R"].cpu().numpy(), want["R"], np.abs(want["R"]).max() + 1e-9)
    fo = cenv.rollout(kernel_layout(acts.astype(np.uint8), cenv.ld), gamma=1.0, g_mode=1)
    close(fo["G"].cpu().numpy(), want["G"], np.abs(want["R"]).max() + 1e-9)
    # Collision with a single agent that starts on the landmark: done after the first step, episode length 1
    lm = rng.random((E, 1, 2)) * 3
    kenv = s.BatchedCollisionAvoidance(3, 1, n_envs=E, starts=lm.copy(), landmarks=lm)
    zero = torch.zeros(4, 2, kenv.ld, device="cuda")
    ko = kenv.rollout(zero, gamma=0.9, g_mode=1)
    assert np.array_equal(ko["n_active"].cpu().numpy(), np.ones(E, dtype=np.int32))
    assert float(ko["R"].abs().max()) == 0.0 and bool(kenv.agent_done[:, :E].all())
    assert float(ko["G"].abs().max()) == 0.0


@pytest.mark.parametrize("kind,size,A,E,T,coarse,seed", [("continuous", 5, 3, 300, 40, 6, 1), ("continuous", 10, 8, 130, 15, None, 2),
                                                          ("discretized", 5, 3, 300, 40, 20, 0), ("discretized", 3, 32, 33, 6, 3, 3),
                                                          # >= 2^18 envs with A <= 4: the register-capped kernel builds
                                                          ("continuous", 5, 3, (1 << 18) + 8, 6, 6, 4),
                                                          ("discretized", 5, 4, 1 << 18, 5, 6, 5)])
@pytest.mark.parametrize("g_mode", [0, 1, 2])
def test_coverage_float_fused_rollout(kind, size, A, E, T, coarse, seed, g_mode):
    s = smarl()
    rng = np.random.default_rng(seed)
    w = (1.0 + (np.arange(A) % 3)).tolist()
    fv = no.coverage_fieldview(size, A)
    lam_np, gamma, thr = np.linspace(0.1, 0.3, A), 0.99, np.full(A, 0.3 * T)
    if kind == "continuous":
        starts = rng.random((E, A, 2)) * size
        actions = rng.normal(0, 0.8, size=(T, E, A, 2)).astype(np.float32)
        env = s.BatchedCoverageContinuous(size, A, n_envs=E, weights=w, coarseness=coarse, starts=starts)
        act_k = kernel_layout(actions.reshape(T, E, 2 * A), env.ld)
    else:
        zoom = coarse / size
        starts = np.floor(rng.random((E, A, 2)) * size * zoom) / zoom
        actions = rng.integers(0, 9, size=(T, E, A))
        env = s.BatchedCoverageDiscretized(size, A, n_envs=E, coarseness=coarse, weights=w, starts=starts)
        act_k = kernel_layout(actions.astype(np.uint8), env.ld)
    pos = starts.copy()
    def step_fn(t):
        nonlocal pos
        if kind == "continuous":
            pos, r, c, _ = no.coverage_continuous_step(pos, actions[t].astype(np.float64), size, fv, w, coarse, exact_pow=False)
        else:
            pos, r, c, _ = no.coverage_discretized_step(pos, actions[t], size, coarse, fv, w, exact_pow=False)
        return r.astype(np.float32).astype(np.float64), c.astype(np.float32).astype(np.float64)
    want = no.rollout(step_fn, T, gamma, lam_np)
    want_G = {0: None, 1: want["G"], 2: no.discounted_terms(want["mod_reward"], gamma)}[g_mode]
    out = env.rollout(act_k, lambdas=torch.as_tensor(lam_np, device="cuda"), gamma=gamma, thresholds=thr, g_mode=g_mode)
    assert np.array_equal(env.state().cpu().numpy(), pos)                       # bit-exact float64 final positions
    scale = np.abs(want["modR"]).max()
    close(out["R"].cpu().numpy(), want["R"], scale)
    close(out["modR"].cpu().numpy(), want["modR"], scale)
    close(out["C"].cpu().numpy(), want["C"], np.abs(want["C"]).max())
    if g_mode:
        close(out["G"].cpu().numpy(), want_G, scale)
    st = out["stats"]
    close(st.cost_sum.cpu().numpy(), want["C"].sum(0), np.abs(want["C"].sum(0)).max())
    close(st.return_sum.cpu().numpy(), want["R"].sum(0), scale * E)
    assert float(st.count) == E


# ----------------------------------------------------------------------------- host mirrors: MetaAgent gating, Buffer
def test_meta_agent_gating_and_buffer_mirror_follow_main_py():
    """main.py:25-68: lambda is updated once per meta cycle from the episodes of the LAST
    (n_agents_learning_cycles - start) agent cycles only (meta_agent.py:19,:25-39); Buffer.mean_score
    reports means over the last n episodes with costs relative to the thresholds (buffer.py:45-48)."""
    import argparse
    s = smarl()
    rng = np.random.default_rng(0)
    size, A, E, T, gamma, lr, decay = 5, 3, 64, 12, 0.99, 0.05, 2.0
    thr = [4.0, 5.0, 6.0]
    env = s.BatchedCoverageDiscrete(size, A, n_envs=E, weights=[1.0, 2.0, 3.0], starts=rng.integers(0, size, (E, A, 2)))
    n_cycles, start = 5, 3
    meta = s.BatchedMetaAgent(env.constraint_space, gamma, lr, thr, start_learning_cycle=start, decay=decay, lambda_0=0.1,
                              n_agents=A)
    hist = s.BatchedBuffer(argparse.Namespace(gamma=gamma, thresholds=thr))
    lam_ref, lr_ref = np.full(A, 0.1), lr
    for mc in range(3):
        recorded = []
        for ac in range(n_cycles):
            act = torch.as_tensor(rng.integers(0, 5, (T, E, A)).astype(np.uint8), device="cuda")
            out = env.rollout_closed_loop(lambda obs, t: act[t], T, meta.lambdas, gamma, thresholds=thr)
            hist.extend(out["R"], out["modR"], out["C"])
            meta.step(out["stats"])
            if ac >= start:
                recorded.append(out["C"].cpu().numpy().astype(np.float64))
            meta.increment_learning_cycle()
        meta.update()
        hist.append_lambdas(meta.lambdas)
        lam_ref = no.lambda_update(lam_ref, np.concatenate(recorded).mean(0), thr, lr_ref)
        lr_ref /= decay
        np.testing.assert_allclose(meta.lambdas.cpu().numpy(), lam_ref, rtol=1e-13)
        assert meta.learning_cycle == 0 and abs(meta.lr - lr_ref) < 1e-15
    assert hist.scores.shape == (3 * n_cycles * E, A) and hist.constraints.shape == (3 * n_cycles * E, A)
    sc, msc, viol = hist.mean_score(n=100)
    np.testing.assert_allclose(sc, hist.scores[-100:].mean(0))
    np.testing.assert_allclose(viol, hist.constraints[-100:].mean(0) - np.array(thr))
    assert len(hist.lambdas) == 3 and np.allclose(hist.lambdas[-1], lam_ref)

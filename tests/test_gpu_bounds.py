"""Own bounds check (compute-sanitizer is closed on this pool): every buffer handed to the C ABI is
carved out of a larger sentinel-filled allocation; after the calls the sentinels on both sides must be
intact, for ragged n_envs (ld > n_envs) and several agent counts."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 4096          # bytes on each side


class Arena:
    def __init__(self):
        self.bufs = []

    def make(self, rows, ld, dtype, fill=None):
        n = rows * ld
        item = torch.empty((), dtype=dtype).element_size()
        raw = torch.full((n * item + 2 * GUARD,), 0xA5, dtype=torch.uint8, device="cuda")
        view = raw[GUARD:GUARD + n * item].view(dtype).view(rows, ld)
        if fill is not None:
            view.copy_(fill)
        else:
            view.zero_()
        self.bufs.append(raw)
        return view

    def check(self):
        for i, raw in enumerate(self.bufs):
            assert bool((raw[:GUARD] == 0xA5).all()) and bool((raw[-GUARD:] == 0xA5).all()), f"buffer {i} overrun"


@pytest.mark.parametrize("A,E", [(3, 37), (16, 1001), (32, 5), (1, 16)])
def test_coverage_step_rollout_returns_stay_in_bounds(A, E):
    from safe_multiagent_rl_b200 import _lib
    lib = _lib.load()
    ld = (E + 15) // 16 * 16
    T, S = 6, 9
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    ar = Arena()
    u8, f32, f64, i32 = torch.uint8, torch.float32, torch.float64, torch.int32
    sx = ar.make(A, ld, u8, torch.randint(0, S, (A, ld), generator=g, device="cuda", dtype=u8))
    sy = ar.make(A, ld, u8, torch.randint(0, S, (A, ld), generator=g, device="cuda", dtype=u8))
    px, py = ar.make(A, ld, u8), ar.make(A, ld, u8)
    acts = ar.make(T * A, ld, u8, torch.randint(0, 5, (T * A, ld), generator=g, device="cuda", dtype=u8))
    obs, rew = ar.make(2 * A, ld, f32), ar.make(T * A, ld, f32)
    cost, done, pen = ar.make(T * A, ld, u8), ar.make(T * A, ld, u8), ar.make(T, ld, f32)
    R, M, Cs, G = ar.make(A, ld, f32), ar.make(A, ld, f32), ar.make(A, ld, i32), ar.make(T * A, ld, f32)
    gs = ar.make(2 * T, ld, f32)
    lut = ar.make(1, 16, f32, torch.rand(1, 16, device="cuda"))
    w = ar.make(1, 32, f32, torch.rand(1, 32, device="cuda"))
    lam = ar.make(1, 32, f64, torch.rand(1, 32, device="cuda", dtype=f64))
    thr = ar.make(1, 32, f64, torch.full((1, 32), 3.0, device="cuda", dtype=f64))
    stats = ar.make(1, lib.smarl_stats_len(A, A), f64)
    scratch = ar.make(1, max(1, lib.smarl_stats_scratch_len(A, A, E)), f64)
    P = _lib.ptr
    st = torch.cuda.current_stream().cuda_stream
    p = _lib.CoverageParams(S, A, 9, 0, P(lut), P(w))
    _lib.check(lib.smarl_grid_reset(P(sx), P(sy), P(px), P(py), P(obs), A, E, ld, st))
    for t in range(T):
        _lib.check(lib.smarl_coverage_step(C.byref(p), P(px), P(py), P(acts[t * A:]), P(obs), P(rew[t * A:]),
                                           P(cost[t * A:]), P(done[t * A:]), P(lam), P(pen[t:]), E, ld, st))
    for g_mode in (1, 2, 3):
        acc = _lib.Accounting(0.99, T, g_mode, P(thr))
        _lib.check(lib.smarl_rollout_returns(C.byref(acc), P(rew), P(cost), 0, P(pen), None, P(R), P(M), P(Cs), P(G),
                                             P(stats), P(scratch), A, A, E, ld, st))
    for g_mode in (0, 1, 2):
        acc = _lib.Accounting(0.99, T, g_mode, P(thr))
        _lib.check(lib.smarl_coverage_rollout(C.byref(p), C.byref(acc), P(sx), P(sy), P(acts), P(lam), P(px),
                                              P(py), P(R), P(M), P(Cs), P(G) if g_mode else None,
                                              P(gs) if g_mode == 1 else None, P(stats), P(scratch), E, ld, st))
    # lean path: one unweighted env-reward row per step, no done flags, shared-reward accounting
    ps = _lib.CoverageParams(S, A, 9, 1, P(lut), P(w))
    renv = ar.make(T, ld, f32)
    for t in range(T):
        _lib.check(lib.smarl_coverage_step(C.byref(ps), P(px), P(py), P(acts[t * A:]), P(obs), P(renv[t:]),
                                           P(cost[t * A:]), None, P(lam), P(pen[t:]), E, ld, st))
    for g_mode in (0, 1, 2):
        acc = _lib.Accounting(0.99, T, g_mode, P(thr))
        _lib.check(lib.smarl_rollout_returns_shared(C.byref(acc), P(renv), P(w), P(cost), 0, P(pen), None, P(R), P(M), P(Cs),
                                                    P(G) if g_mode else None, P(stats), P(scratch), A, A, E, ld, st))
    _lib.check(lib.smarl_lambda_update(P(lam), P(stats), P(thr), 0.01, A, A, st))
    torch.cuda.synchronize()
    ar.check()
    assert float(stats[0, -1]) == E


@pytest.mark.parametrize("A,E", [(3, 37), (8, 1001), (32, 5)])
def test_congestion_and_collision_stay_in_bounds(A, E):
    from safe_multiagent_rl_b200 import _lib
    lib = _lib.load()
    ld = (E + 15) // 16 * 16
    T, S, L = 5, 6, 2
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    ar = Arena()
    u8, f32, f64, i32 = torch.uint8, torch.float32, torch.float64, torch.int32
    P = _lib.ptr
    st = torch.cuda.current_stream().cuda_stream
    # Congestion
    sx = ar.make(A, ld, u8, torch.randint(0, S, (A, ld), generator=g, device="cuda", dtype=u8))
    sy = ar.make(A, ld, u8, torch.randint(0, S, (A, ld), generator=g, device="cuda", dtype=u8))
    px, py, mv = ar.make(A, ld, u8), ar.make(A, ld, u8), ar.make(T * A, ld, u8)
    acts = ar.make(T * A, ld, u8, torch.randint(0, 5, (T * A, ld), generator=g, device="cuda", dtype=u8))
    obs, rew = ar.make(2 * A, ld, f32), ar.make(A, ld, f32)
    cost, done, pen = ar.make(1, ld, i32), ar.make(A, ld, u8), ar.make(1, ld, f32)
    R, M, Cs, G, gs = ar.make(A, ld, f32), ar.make(A, ld, f32), ar.make(1, ld, i32), ar.make(T * A, ld, f32), ar.make(T, ld, f32)
    dem = ar.make(1, (S + 1) * (S + 1), f64, torch.rand(1, (S + 1) * (S + 1), device="cuda", dtype=f64) * 8 + 2)
    lam = ar.make(1, 32, f64, torch.rand(1, 32, device="cuda", dtype=f64))
    thr = ar.make(1, 32, f64, torch.full((1, 32), 3.0, device="cuda", dtype=f64))
    stats = ar.make(1, lib.smarl_stats_len(A, 1), f64)
    scratch = ar.make(1, max(1, lib.smarl_stats_scratch_len(A, 1, E)), f64)
    _lib.check(lib.smarl_grid_reset(P(sx), P(sy), P(px), P(py), P(obs), A, E, ld, st))
    for mode in (0, 2, 1):
        cp = _lib.CongestionParams(S, A, P(dem), mode, 0, 3865470567, 7, 123456789012, None, None)
        _lib.check(lib.smarl_congestion_step(C.byref(cp), P(px), P(py), P(acts), P(mv), P(obs), P(rew), P(cost), P(done),
                                             P(lam), P(pen), 3, E, ld, st))
        for g_mode in (0, 1, 2):
            acc = _lib.Accounting(0.9, T, g_mode, P(thr))
            _lib.check(lib.smarl_congestion_rollout(C.byref(cp), C.byref(acc), P(sx), P(sy), P(acts), P(mv), P(lam), P(px),
                                                    P(py), P(R), P(M), P(Cs), P(G) if g_mode else None,
                                                    P(gs) if g_mode == 1 else None, P(stats), P(scratch), E, ld, st))
    # Collision
    fx = ar.make(A, ld, f64, torch.rand(A, ld, device="cuda", dtype=f64) * S)
    fy = ar.make(A, ld, f64, torch.rand(A, ld, device="cuda", dtype=f64) * S)
    qx, qy = ar.make(A, ld, f64), ar.make(A, ld, f64)
    lm = ar.make(2 * L, ld, f64, torch.rand(2 * L, ld, device="cuda", dtype=f64) * S)
    fa = ar.make(T * 2 * A, ld, f32, torch.randn(T * 2 * A, ld, device="cuda") * 0.5)
    adone, dout, elen, nact = ar.make(A, ld, u8), ar.make(A, ld, u8), ar.make(1, ld, i32), ar.make(1, ld, i32)
    obs2 = ar.make(2 * A + 2 * L, ld, f32)
    gs2 = ar.make(2 * T, ld, f32)
    kp = _lib.CollisionParams(S, A, L, 1, 0.25, 1, 0)
    _lib.check(lib.smarl_collision_reset(C.byref(kp), P(fx), P(fy), P(lm), P(qx), P(qy), P(adone), P(elen), P(obs2), E, ld, st))
    for t in range(T):
        _lib.check(lib.smarl_collision_step(C.byref(kp), P(qx), P(qy), P(adone), P(fa[t * 2 * A:]), P(lm), P(obs2), P(rew),
                                            P(cost), P(dout), P(elen), P(lam), P(pen), E, ld, st))
    for g_mode in (0, 1, 2):
        acc = _lib.Accounting(0.9, T, g_mode, P(thr))
        _lib.check(lib.smarl_collision_rollout(C.byref(kp), C.byref(acc), P(fx), P(fy), P(lm), P(fa), P(lam), P(qx), P(qy),
                                               P(adone), P(nact), P(R), P(M), P(Cs), P(G) if g_mode else None,
                                               P(gs2) if g_mode == 1 else None, P(stats), P(scratch), E, ld, st))
    # random starts + float coverage
    _lib.check(lib.smarl_random_starts_u8(1, S, 5, 2, 77, P(sx), P(sy), A, E, ld, st))
    _lib.check(lib.smarl_random_starts_f64(3, S, 1.5, 5, 2, 77, 0, P(fx), P(fy), ld, A, E, st))
    _lib.check(lib.smarl_random_starts_f64(2, S, 0.0, 5, 2, 77, A, P(lm), lm.data_ptr() + 8 * ld, 2 * ld, L, E, st))
    fp = _lib.CoverageFloatParams(S, A, 0, 1, 2.0, 0.7, 0.0, 0.0, 0.0, 0.0, None)
    costf = ar.make(A, ld, f32)
    _lib.check(lib.smarl_coverage_float_reset(P(fx), P(fy), P(qx), P(qy), P(obs), A, E, ld, st))
    _lib.check(lib.smarl_coverage_float_step(C.byref(fp), P(qx), P(qy), P(fa), P(obs), P(rew), P(costf), P(done), P(lam),
                                             P(pen), E, ld, st))
    fp = _lib.CoverageFloatParams(S, A, 1, 0, 2.0, 0.0, 1.5, S * 1.5, 0.66, 0.94, None)
    _lib.check(lib.smarl_coverage_float_step(C.byref(fp), P(qx), P(qy), P(acts), P(obs), P(rew), P(costf), P(done), P(lam),
                                             P(pen), E, ld, st))
    torch.cuda.synchronize()
    ar.check()

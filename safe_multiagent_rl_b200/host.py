"""numpy-facing wrapper of the host-buffer C-ABI entry points (smarl_host_*_rollout): host arrays in
the reference's env-major orientation in, episode products out.  This is the call a user of the
reference who keeps actions / results in host memory makes; bench.py times the same entry point with
pinned buffers as ``e2e``."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .envs.congestion import keep_threshold
from .envs.coverage import penalty_table


def _am(a, ld, dtype):
    """[..., E, rows] -> contiguous agent-major [..., rows, ld]."""
    a = np.asarray(a)
    out = np.zeros(a.shape[:-2] + (a.shape[-1], ld), dtype=dtype)
    out[..., : a.shape[-2]] = np.swapaxes(a, -1, -2)
    return out


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostRollout:
    """One session per shape.  ``kind``: "coverage" | "congestion" | "collision"."""

    def __init__(self, kind, n_agents, n_steps, n_envs, n_landmarks=1):
        self.lib = _lib.load()
        self.kind = kind
        self.A, self.T, self.E, self.L = int(n_agents), int(n_steps), int(n_envs), int(n_landmarks)
        self.K = self.A if kind == "coverage" else 1
        code = {"coverage": _lib.ENV_COVERAGE, "congestion": _lib.ENV_CONGESTION, "collision": _lib.ENV_COLLISION}[kind]
        self._sess = C.c_void_p()
        _lib.check(self.lib.smarl_host_session_create(C.byref(self._sess), code, self.A, self.T, self.E, self.L))
        self.ld = int(self.lib.smarl_host_session_ld(self._sess))

    def close(self):
        if self._sess:
            self.lib.smarl_host_session_destroy(self._sess)
            self._sess = C.c_void_p()

    __del__ = close

    def _outputs(self):
        return (np.zeros((self.A, self.ld), np.float32), np.zeros((self.A, self.ld), np.float32),
                np.zeros((self.K, self.ld), np.int32), np.zeros(self.lib.smarl_stats_len(self.A, self.K), np.float64))

    def _result(self, R, M, Cs, st, **extra):
        E = self.E
        return dict(R=R[:, :E].T, modR=M[:, :E].T, C=Cs[:, :E].T, stats=st, **extra)

    def coverage(self, size, starts, actions, weights=None, lambdas=None, gamma=0.99, thresholds=None,
                 fieldview_size=None, packed4=False):
        """starts [E, A, 2] ints, actions [T, E, A] ints 0..4 -> dict(R [E,A], modR [E,A], C [E,A], stats).
        The arrays go to the library as they are (env-major, smarl_host_coverage_rollout_envmajor: the layout
        change runs on the device).  packed4: ship the actions agent-major as two 4-bit values per byte
        (half the PCIe traffic; the packing is done here on the CPU)."""
        A, ld, E, T = self.A, self.ld, self.E, self.T
        _, table = penalty_table(size, A, fieldview_size)
        lut = np.ascontiguousarray(table if A > 1 else table[:0], dtype=np.float32)
        w = None if weights is None else np.ascontiguousarray(np.asarray(weights, np.float64)[:A], dtype=np.float32)
        lam = None if lambdas is None else np.ascontiguousarray(lambdas, dtype=np.float64)
        thr = None if thresholds is None else np.ascontiguousarray(thresholds, dtype=np.float64)
        p = _lib.CoverageParams(size, A, len(lut), 0, _p(lut) if len(lut) else None, _p(w))
        acc = _lib.Accounting(gamma, T, 0, _p(thr))
        st = np.zeros(self.lib.smarl_stats_len(A, self.K), np.float64)
        if not packed4:
            starts = np.ascontiguousarray(starts, dtype=np.uint8)
            actions = np.ascontiguousarray(actions, dtype=np.uint8)
            if starts.shape != (E, A, 2) or actions.shape != (T, E, A):
                raise ValueError(f"expected starts {(E, A, 2)} and actions {(T, E, A)}, got {starts.shape}, {actions.shape}")
            R, M, Cs = np.empty((E, A), np.float32), np.empty((E, A), np.float32), np.empty((E, A), np.int32)
            _lib.check(self.lib.smarl_host_coverage_rollout_envmajor(self._sess, C.byref(p), C.byref(acc), _p(starts),
                                                                     _p(actions), _p(lam), _p(R), _p(M), _p(Cs), _p(st)))
            return dict(R=R, modR=M, C=Cs, stats=st)
        sx, sy = _am(np.asarray(starts)[:, :, 0], ld, np.uint8), _am(np.asarray(starts)[:, :, 1], ld, np.uint8)
        act = _am(actions, ld, np.uint8)
        act = np.ascontiguousarray(act[..., 0::2] | (act[..., 1::2] << 4))            # [T, A, ld/2]
        R, M, Cs, st = self._outputs()
        _lib.check(self.lib.smarl_host_coverage_rollout_packed4(self._sess, C.byref(p), C.byref(acc), _p(sx), _p(sy),
                                                                _p(act), _p(lam), _p(R), _p(M), _p(Cs), _p(st)))
        return self._result(R, M, Cs, st)

    def congestion(self, size, starts, actions, demand_rate, noise=0.0, seed=0, env_offset=0, moves=None, lambdas=None,
                   gamma=0.99, thresholds=None, episode=0):
        """starts [E, A, 2] ints, actions (and recorded moves) [T, E, A] ints -> dict(R [E,A], modR [E,A], C [E,1], stats);
        the arrays go to smarl_host_congestion_rollout_envmajor as they are.  ``episode`` indexes the Philox noise realisation
        (pass a running episode counter to get fresh noise per call, as the reference's random() gives)."""
        A, E, T = self.A, self.E, self.T
        dem = np.ascontiguousarray(demand_rate, dtype=np.float64)
        assert dem.shape == (size + 1, size + 1)
        lam = None if lambdas is None else np.ascontiguousarray(lambdas, dtype=np.float64)
        thr = None if thresholds is None else np.ascontiguousarray(thresholds, dtype=np.float64)
        starts = np.ascontiguousarray(starts, dtype=np.uint8)
        actions = np.ascontiguousarray(actions, dtype=np.uint8)
        mv = None if moves is None else np.ascontiguousarray(moves, dtype=np.uint8)
        if starts.shape != (E, A, 2) or actions.shape != (T, E, A) or (mv is not None and mv.shape != (T, E, A)):
            raise ValueError(f"expected starts {(E, A, 2)} and actions / moves {(T, E, A)}")
        mode = 1 if moves is not None else (2 if noise > 0 else 0)
        R, M, Cs = np.empty((E, A), np.float32), np.empty((E, A), np.float32), np.empty((E, 1), np.int32)
        st = np.zeros(self.lib.smarl_stats_len(A, 1), np.float64)
        p = _lib.CongestionParams(size, A, _p(dem), mode, int(episode) & 0x1FFFFFFF, keep_threshold(noise), seed & (2 ** 64 - 1),
                                  env_offset, None, None)
        acc = _lib.Accounting(gamma, T, 0, _p(thr))
        _lib.check(self.lib.smarl_host_congestion_rollout_envmajor(self._sess, C.byref(p), C.byref(acc), _p(starts),
                                                                   _p(actions), _p(mv), _p(lam), _p(R), _p(M), _p(Cs), _p(st)))
        return dict(R=R, modR=M, C=Cs, stats=st)

    def collision(self, size, starts, landmarks, actions, lambdas=None, gamma=0.99, thresholds=None, agents_size=0.25):
        """starts [E,A,2] f64, landmarks [E,L,2] f64, actions [T,E,A,2] f32 -> dict(R, modR [E,A], C [E,1], n_active [E],
        stats); the arrays go to smarl_host_collision_rollout_envmajor as they are."""
        A, L, T, E = self.A, self.L, self.T, self.E
        lam = None if lambdas is None else np.ascontiguousarray(lambdas, dtype=np.float64)
        thr = None if thresholds is None else np.ascontiguousarray(thresholds, dtype=np.float64)
        starts = np.ascontiguousarray(starts, dtype=np.float64)
        landmarks = np.ascontiguousarray(landmarks, dtype=np.float64)
        actions = np.ascontiguousarray(actions, dtype=np.float32)
        if starts.shape != (E, A, 2) or landmarks.shape != (E, L, 2) or actions.shape != (T, E, A, 2):
            raise ValueError(f"expected starts {(E, A, 2)}, landmarks {(E, L, 2)}, actions {(T, E, A, 2)}")
        n_active = np.zeros(E, np.int32)
        R, M, Cs = np.empty((E, A), np.float32), np.empty((E, A), np.float32), np.empty((E, 1), np.int32)
        st = np.zeros(self.lib.smarl_stats_len(A, 1), np.float64)
        p = _lib.CollisionParams(size, A, L, 0, agents_size, 0, 0)
        acc = _lib.Accounting(gamma, T, 0, _p(thr))
        _lib.check(self.lib.smarl_host_collision_rollout_envmajor(self._sess, C.byref(p), C.byref(acc), _p(starts),
                                                                  _p(landmarks), _p(actions), _p(lam), _p(R), _p(M), _p(Cs),
                                                                  _p(n_active), _p(st)))
        return dict(R=R, modR=M, C=Cs, stats=st, n_active=n_active)

"""ctypes wrapper of the plain-C oracle (oracle/c/oracle.c).  TEST INFRASTRUCTURE ONLY.

Arrays are numpy, in the product's agent-major layout ``[rows, ld]`` so that device buffers
(``tensor.cpu().numpy()``) compare directly.  Build: ``make -C oracle/c`` (done by
``__graft_entry__.build()``)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "c", "oracle.c")):
            subprocess.run(["make", "-s", "-C", os.path.join(HERE, "c")], check=True)
        _lib = C.CDLL(LIB)
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def num_threads():
    return load().oracle_num_threads()


def coverage_rollout(size, start_x, start_y, actions, lut, weights, lambdas, gamma, n_envs, want_G=False):
    """start_x/y u8 [A, ld], actions u8 [T, A, ld] -> dict(final_x, final_y, R, modR, C, G, reward_last)."""
    T, A, ld = actions.shape
    start_x, start_y, actions = _c(start_x, np.uint8), _c(start_y, np.uint8), _c(actions, np.uint8)
    lut, weights, lambdas = _c(lut, np.float64), _c(weights, np.float64), _c(lambdas, np.float64)
    out = dict(final_x=np.zeros((A, ld), np.uint8), final_y=np.zeros((A, ld), np.uint8),
               R=np.zeros((A, ld)), modR=np.zeros((A, ld)), C=np.zeros((A, ld), np.int32),
               G=np.zeros((T, A, ld)) if want_G else None, reward_last=np.zeros((A, ld)))
    rc = load().oracle_coverage_rollout(
        C.c_int(size), C.c_int(A), C.c_int64(n_envs), C.c_int64(ld), C.c_int(T), _p(start_x), _p(start_y),
        _p(actions), _p(lut), C.c_int(len(lut)), _p(weights), _p(lambdas), C.c_double(gamma), _p(out["final_x"]),
        _p(out["final_y"]), _p(out["R"]), _p(out["modR"]), _p(out["C"]), _p(out["G"]), _p(out["reward_last"]))
    assert rc == 0
    return out


def congestion_rollout(size, start_x, start_y, actions, demand, lambdas, gamma, n_envs, moves=None, noise_mode=0,
                       keep_threshold=1 << 32, seed=0, env_offset=0, round_f32=True, want_G=False):
    T, A, ld = actions.shape
    start_x, start_y, actions = _c(start_x, np.uint8), _c(start_y, np.uint8), _c(actions, np.uint8)
    moves, demand, lambdas = _c(moves, np.uint8), _c(demand, np.float64), _c(lambdas, np.float64)
    out = dict(final_x=np.zeros((A, ld), np.uint8), final_y=np.zeros((A, ld), np.uint8),
               R=np.zeros((A, ld)), modR=np.zeros((A, ld)), C=np.zeros((1, ld), np.int32),
               G=np.zeros((T, A, ld)) if want_G else None)
    rc = load().oracle_congestion_rollout(
        C.c_int(size), C.c_int(A), C.c_int64(n_envs), C.c_int64(ld), C.c_int(T), _p(start_x), _p(start_y),
        _p(actions), _p(moves), C.c_int(noise_mode), C.c_uint64(keep_threshold), C.c_uint64(seed),
        C.c_int64(env_offset), _p(demand), _p(lambdas), C.c_double(gamma), C.c_int(int(round_f32)),
        _p(out["final_x"]), _p(out["final_y"]), _p(out["R"]), _p(out["modR"]), _p(out["C"]), _p(out["G"]))
    assert rc == 0
    return out


def collision_rollout(size, start_x, start_y, landmarks, actions, lambdas, gamma, n_envs, agents_size=0.25,
                      round_f32=True, want_G=False):
    T, A2, ld = actions.shape
    A = A2 // 2
    L = landmarks.shape[0] // 2
    start_x, start_y, landmarks = _c(start_x, np.float64), _c(start_y, np.float64), _c(landmarks, np.float64)
    actions, lambdas = _c(actions, np.float32), _c(lambdas, np.float64)
    out = dict(final_x=np.zeros((A, ld)), final_y=np.zeros((A, ld)), final_done=np.zeros((A, ld), np.uint8),
               n_active=np.zeros(ld, np.int32), R=np.zeros((A, ld)), modR=np.zeros((A, ld)),
               C=np.zeros((1, ld), np.int32), G=np.zeros((T, A, ld)) if want_G else None)
    rc = load().oracle_collision_rollout(
        C.c_int(size), C.c_int(A), C.c_int(L), C.c_int64(n_envs), C.c_int64(ld), C.c_int(T), _p(start_x),
        _p(start_y), _p(landmarks), _p(actions), _p(lambdas), C.c_double(gamma), C.c_double(agents_size),
        C.c_int(int(round_f32)), _p(out["final_x"]), _p(out["final_y"]), _p(out["final_done"]), _p(out["n_active"]),
        _p(out["R"]), _p(out["modR"]), _p(out["C"]), _p(out["G"]))
    assert rc == 0
    return out

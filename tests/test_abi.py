"""CPU checks of the drop-in boundary: libsmarl.so loads, exports every symbol that
include/smarl.h declares, the ctypes prototypes cover the header, argument validation
fails loudly before any CUDA work, and the product has no CPU / oracle fallback."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "smarl.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smarl_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from safe_multiagent_rl_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_symbols_exported_and_bound(lib):
    from safe_multiagent_rl_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in smarl.h but not exported"
    assert sorted(_lib.PROTOTYPES) == syms, "ctypes prototypes and header disagree"


def test_every_declared_symbol_cites_the_reference():
    text = open(HEADER).read()
    for needle in ["coverage.py:", "congestion.py:", "collision_avoidance.py:", "meta_agent.py:", "buffer.py:",
                   "agent.py:", "main.py:"]:
        assert needle in text


def test_abi_struct_sizes():
    from safe_multiagent_rl_b200 import _lib
    assert C.sizeof(_lib.CoverageParams) == 32
    assert C.sizeof(_lib.Accounting) == 24
    assert C.sizeof(_lib.CongestionParams) == 64      # ABI v2: + episode_dev
    assert C.sizeof(_lib.CollisionParams) == 32


def test_validation_errors_before_any_cuda_work(lib):
    from safe_multiagent_rl_b200 import _lib
    p = _lib.CoverageParams(5, 3, 0, 0, None, None)
    fake = C.c_void_p(0x1000)   # never dereferenced: validation rejects the call first
    rc = lib.smarl_coverage_step(C.byref(p), fake, fake, fake, None, fake, fake, None, None, None, 10, 10, None)
    assert rc == -1 and b"multiple of 16" in lib.smarl_last_error()
    p = _lib.CoverageParams(300, 3, 0, 0, None, None)
    rc = lib.smarl_coverage_step(C.byref(p), fake, fake, fake, None, fake, fake, None, None, None, 16, 16, None)
    assert rc == -1 and b"size=300" in lib.smarl_last_error()
    p = _lib.CoverageParams(5, 3, 0, 0, None, None)
    rc = lib.smarl_coverage_step(C.byref(p), C.c_void_p(0x1001), fake, fake, None, fake, fake, None, None, None,
                                 16, 16, None)
    assert rc == -1 and b"aligned" in lib.smarl_last_error()
    with pytest.raises(_lib.SmarlError):
        _lib.check(rc)
    assert lib.smarl_stats_len(16, 16) == 65
    assert lib.smarl_stats_scratch_len(3, 3, 50) == 26      # one row of partials per 32 envs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import safe_multiagent_rl_b200 as s
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.BatchedCoverageDiscrete(5, 3, n_envs=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.BatchedCongestion(3, 3, n_envs=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.BatchedCollisionAvoidance(5, 3, n_envs=4)


def test_missing_library_fails_loudly(tmp_path):
    code = ("import os; os.environ['SMARL_LIB']=%r; from safe_multiagent_rl_b200 import _lib\n"
            "try:\n    _lib.load()\nexcept _lib.SmarlError as e:\n    print('RAISED', e)\n") % str(tmp_path / "nope.so")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert "RAISED" in out.stdout and "no CPU fallback" in out.stdout


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "safe_multiagent_rl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text, f
    code = "import sys, safe_multiagent_rl_b200; print(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.stdout.strip() == "False", out.stderr


def test_host_parameter_helpers_match_oracle():
    """Host-side parameter setup (penalty table, Philox keep threshold) against the oracle."""
    import numpy as np
    from oracle import numpy_oracle as no
    from oracle import philox
    from safe_multiagent_rl_b200.envs.congestion import keep_threshold
    from safe_multiagent_rl_b200.envs.coverage import penalty_table
    for size, A, fv in [(5, 3, None), (32, 16, None), (64, 32, None), (3, 2, 10.0), (10, 7, 2.5)]:
        f, table = penalty_table(size, A, fv)
        want = no.coverage_penalty_lut(size, no.coverage_fieldview(size, A, fv))
        assert f == no.coverage_fieldview(size, A, fv)
        n = len(table)
        assert np.array_equal(table, want[:n]) and not want[n:].any() and (n == 0 or table[-1] > 0)
    for noise in [0.0, 0.1, 0.5, 1.0, 0.3333]:
        assert keep_threshold(noise) == philox.keep_threshold(noise)

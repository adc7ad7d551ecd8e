"""CPU oracle for the env-step hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import
it, and there only as the checker / the timed CPU baseline -- never as the thing
shipped.  ``safe_multiagent_rl_b200`` must not import this package (a test
enforces that) and has no CPU fallback: without the CUDA library it raises.

Contents
--------
numpy_oracle.py      vectorised (over envs) numpy restatement of the reference
                     algorithms; every function cites the reference file:line.
scalar_port.py       one-env-at-a-time Python port with the reference's own loop
                     structure (list-of-agents, per-step Python calls).  It has
                     the reference's performance character and is what the
                     ``cpu_baseline`` / ``--impl reference`` legs time on the GPU
                     box, where ``/root/reference`` does not exist.
philox.py            Philox4x32-10 in numpy (the device RNG for Congestion noise).
reference_harness.py imports the UNMODIFIED reference from ``/root/reference``
                     (this container only) to pin the restatement and to produce
                     ``tests/golden/*.npz`` (script: tests/golden/make_golden.py).
c/                   plain-C restatement (gcc, OpenMP) used for full-size parity
                     checks and as a strong multi-core CPU baseline.

Parity pin: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the reference itself run in this
container: live differential tests (tests/test_oracle_vs_reference.py, skipped
where /root/reference is absent) and the committed fixtures in tests/golden/.
"""

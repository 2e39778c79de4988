"""CPU oracle for the PPO data path of rl-algo-impls.  TEST INFRASTRUCTURE, NOT PRODUCT.

Every function here is a CPU restatement (numpy / torch-CPU, the reference's own
arithmetic libraries) of one reference function on the hot path and cites the
reference file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package, and only as the checker or the timed CPU baseline -- never as a
fallback for the CUDA path.  ``rl_algo_impls_b200`` never imports it.

Pinning.  The reference ships no test, golden vector or fixture for this path
(its only test is ``tests/shared/policy/test_actor_critic.py:8`` -- clamp_actions),
so the reference's own tests leave parity unpinned.  The oracle is instead pinned
against the *live reference code*: ``tests/golden/make_golden.py`` imports the
unmodified reference from ``/root/reference`` in the build container, runs it on
seeded inputs, asserts this oracle reproduces its outputs, and commits the
inputs + reference outputs under ``tests/golden/*.npz``.  ``tests/test_oracle_*``
re-check the oracle against those fixtures on every run (no GPU, no reference).
"""

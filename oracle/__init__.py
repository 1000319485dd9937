"""oracle/ -- TEST INFRASTRUCTURE ONLY.

A CPU (torch fp32 / fp64, numpy) restatement of the reference's 3D CycleGAN hot path
(/root/reference/models/networks3D.py, models/cycle_gan_model.py, test.py:96-185).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from here, and only as the checker or the
timed CPU baseline -- never as the product path.  ``mra_gan_b200`` itself never imports it.

Parity status: the reference ships no tests, golden vectors or fixtures of its own
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF,
imported unmodified in the build container by ``oracle/make_golden.py`` (fixtures under
``tests/golden/``), and -- where /root/reference is present -- against the live reference
in ``tests/test_oracle_vs_reference.py``.
"""

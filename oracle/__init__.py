"""CPU oracle for the gpras GPR hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in ``gpras_b200/`` (the product) may import this package.  The only
permitted users are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``, and there only as
the checker or the timed CPU baseline -- never as a compute fallback.

PARITY PINNED TO THE REFERENCE for ``cells.py``, ``preprocess.py`` and ``metrics.py``: their subjects
(``gpras.preprocess.PreProcessor`` and ``gpras/metrics.py``) are NumPy / scikit-learn code that runs offline, and
``tests/golden/make_golden_reference.py`` committed the reference's own outputs as golden vectors.

PARITY UNPINNED BY THE REFERENCE for the GP core: ``/root/reference`` ships no tests, golden
vectors or expected outputs for ``gpras/gpr.py`` and its arithmetic lives in the
un-vendored, un-pinned third-party packages ``gpflow`` / ``tensorflow`` /
``tensorflow_probability`` (``pyproject.toml:16-23``), none of which can be
installed here.  The oracle therefore restates

* the exact-GP form named by BASELINE.json's north_star (``exact_gp.py``), pinned
  against the independent scikit-learn 1.9.0 ``GaussianProcessRegressor``
  (``tests/golden/make_golden.py`` generates the committed vectors) and against
  central finite differences; and
* GPflow 2.x ``SGPR`` as called from ``gpras/gpr.py:293-308`` (``sgpr.py``, torch
  CPU float64 + autograd), pinned only by algebraic identities (ELBO <= LML,
  Z == X collapse, finite differences).
"""

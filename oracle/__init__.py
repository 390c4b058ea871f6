"""CPU oracle for the MAGI posterior-evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``magi_v2_b200``)
may import this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` -- there only as the checker / the timed CPU baseline.

Parity pinning status (see DESIGN.md "Oracle"):
  * ``build_matrices`` (magi_v2.py:774-823) is pinned against the GENUINE
    reference code, executed in the build container through
    ``oracle/ref_loader.py`` (TF-stub import); vectors committed under
    ``tests/golden/`` by ``oracle/make_golden.py``.
  * ``discretize`` / ``linear_interpolate`` likewise pinned against the genuine
    reference methods.
  * ``log_posterior`` (magi_v2.py:308-348) restates a TFP/TF graph that cannot
    be executed here (TensorFlow is not installable): op-for-op restatement,
    checked by two independent implementations (torch autograd vs. analytic
    numpy) -- the reference itself has no golden vector for it:
    "parity unpinned" for that function.
  * leapfrog / HMC / dual averaging / NUTS (``nuts_transition``: recursive,
    single chain) restate tensorflow-probability==0.24.0 (requirements.txt:8),
    which is not vendored: "parity unpinned".  The notebook's printed theta
    means (vignette.ipynb:281-283) come from an earlier run of the notebook
    (its saved predict cell ends in an exception) and do not pin this code
    (profiles/r01_notes.md).
"""

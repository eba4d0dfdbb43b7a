"""CPU oracle for the Mava Anakin PPO hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``mava_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import or execute it, and only as the checker / the CPU baseline.

Parity status (see DESIGN.md "Oracle"):

* ``oracle.threefry``   - pinned: Random123 threefry2x32 known-answer vectors and the
  ``jax.random.split`` values published in the JAX documentation.
* ``oracle.ppo``        - restated line by line from ``/root/reference/mava/systems/ppo``;
  the reference's own tests assert no values (test/integration_test.py:35-46), so
  **parity unpinned** beyond the restatement itself.
* ``oracle.rware`` / ``oracle.lbf`` - the dynamics live in the third-party ``jumanji``
  package (requirements/requirements.txt:12, an unpinned fork) which is absent from
  ``/root/reference`` and from this image: **parity unpinned**.  They restate the published
  Jumanji algorithm; the Mava wrapper stack on top is restated from the reference files.
"""

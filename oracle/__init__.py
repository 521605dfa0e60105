"""CPU oracle for the cube-transition hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's gym-cube
environment (``gym-cube/gym_cube/envs/cube_env.py`` and ``assets/py333.py``,
plus the un-vendored ``assets/py222.py``; see ``oracle/tables.py`` for the
citations).  It exists to *check* the CUDA path and to be *timed beside it* as
the CPU baseline.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``rubiks_cube_solver_b200/`` imports it: the product path has no
CPU fallback and raises when the CUDA library is missing.

Parity pin status
-----------------
* 3x3x3: pinned.  Every table and function is compared against the reference's
  own ``py333.py`` / ``cube_env.py`` (imported from ``/root/reference`` under
  the test-only shims of ``oracle/ref_harness.py`` when that tree is present)
  and against the golden vectors under ``tests/golden/`` that
  ``oracle/gen_golden.py`` produced from that same shimmed reference.
* 2x2x2: **parity unpinned at the py222 boundary** -- ``assets/py222.py`` is a
  third-party module (MeepMoop/py222, no version pin anywhere in the reference)
  that is absent from the reference tree.  Its published algorithm is restated
  in ``oracle/tables.py``; everything the reference does
  *around* it (``cube_env.py`` one-hot layout, reset/step, ADI expansion) is
  pinned by running the reference's ``cube_env.py`` on top of that restatement.
  The restatement itself is pinned STATISTICALLY by the reference's own trained
  checkpoint ``pretrained/222model.pt`` (``oracle/gen_pin222.py`` ->
  ``tests/golden/pin222.npz``, ``tests/test_pin222.py``): run through the
  reference's ``cube_env.py`` + ``model.py`` on the restatement it solves 100 %
  of 200 scrambles at depths 1-5 and 96.5 % at depth 8, and stops solving as
  soon as a move or an encoding table is perturbed.
* The opt-in EXACT 3x3x3 encoding (``cube_np.encode_exact`` / ``decode_exact``) is
  not a restatement of anything in the reference -- the reference has no such
  encoding -- but the SPEC of the library's ``CUBE_ENCODING_EXACT``: it is pinned
  by its own properties (bijection, decode o encode = id, parities), and its edge
  part equals the reference's edge table.
"""

"""CPU oracle for the MoTIF per-pixel inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or as the timed
CPU baseline.  The product path (``motif_b200``) never imports this package and
fails loudly when its CUDA library is missing.

Each function restates, on the CPU, the arithmetic of one reference function
and cites the reference ``file:line`` it follows (paths relative to the
reference checkout).  How the restatement is pinned:

* splat x3 and correlation: against the reference's *own CUDA kernel strings*
  (``models/softsplat*_cp.py``, ``OpticalFlow/correlation.py``) expanded by the
  reference's own ``cupy_kernel`` macro pre-processor and compiled for the host
  with ``g++`` through ``oracle/build_ref.py`` into ``oracle/_ref/`` (the
  kernels are scalar CUDA-C; a small host prelude supplies ``blockIdx``,
  ``atomicAdd`` ...).  Golden input/output vectors produced that way are
  committed under ``tests/golden/``.
* decoder (``Ours.py:659-858``): against the reference's ``LunaTokis.forward``
  imported unmodified under the shims in ``oracle/ref_shims.py`` and run on the
  CPU of the build container; captured hot-path inputs and outputs are committed
  under ``tests/golden/`` together with ``oracle/make_golden.py``.

The reference repository ships no tests or golden vectors of its own
(SURVEY.md section 4), so these captured runs are the pin.
"""

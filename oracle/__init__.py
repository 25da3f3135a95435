"""
ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's batched-einsum execution path, used as the
checker by ``tests/``, ``__graft_entry__.smoke()`` and by ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  Nothing under ``feinsum_b200/``
imports this package; the product path fails loudly without its CUDA library.

Parity status: **pinned against the reference's own acceptance oracle, not
against its loopy kernel.**  The reference defines correctness of every kernel
as agreement with ``numpy.einsum(subscripts, *inputs, optimize="optimal")``
(reference ``src/feinsum/measure.py:149-159,178-192``) and stores no golden
vectors.  ``np_oracle`` restates exactly that expression; the fixtures in
``tests/golden/`` were produced by importing the reference's own front-end and
input generator (``tests/golden/make_golden.py``) in the build container.
The reference's loopy -> OpenCL -> pocl kernel itself cannot be executed here
(``loopy``, ``pyopencl``, ``islpy``, ``pymbolic`` and an OpenCL ICD are not
installed and there is no network), so "matches the loopy kernel" is unpinned
and is replaced by the reference's own tolerance test against numpy.

* ``np_oracle``  -- numpy restatement (inputs, expected outputs, tolerances,
  FLOP/byte model known answers).
* ``cgen``       -- C/OpenMP restatement of the loop nest ``generate_loopy``
  emits (reference ``src/feinsum/codegen/loopy.py:242-315``), the stand-in
  for loopy->pocl when a CPU time is reported.
"""

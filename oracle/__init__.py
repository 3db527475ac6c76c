"""CPU oracle for the drfProc PSD/STI hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the thing timed *beside* the GPU path -- never as a fallback for it.

Two restatements live here (see each module's header for the reference
file:line every function follows):

* ``oracle.ref_port``  -- the reference's call sequence into scipy.signal /
  numpy restated (what ``drfProc.py`` executes on the CPU).  This is the parity
  checker and the "port" CPU baseline.
* ``oracle.np_oracle`` -- an independent float64 restatement in plain numpy
  (no scipy): Kaiser window, DFT, |X|^2 scaling, averaging, fftshift, median.
  Used to cross-check ``ref_port`` and to provide float64 truth.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4).  Both restatements are pinned against (i) outputs of the
reference's own, unmodified ``sti_proc_data`` / ``proc_data`` / ``get_ref``
function bodies executed in the build container (fixtures under
``tests/golden/``, made by ``tools/make_golden.py``) and (ii) scipy's upstream
known-answer tests for the functions the reference calls.
"""

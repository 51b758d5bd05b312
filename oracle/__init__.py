"""CPU oracle for the Visuelle 2.0 forecaster hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline, never on the CUDA product path.

What is here
------------
* ``rnn.py``  - explicit-math restatement (plain torch CPU ops, no nn.GRU /
  nn.MultiheadAttention) of CrossAttnRNN21 / CrossAttnRNN210 /
  CrossAttnRNNDemand *as written* in the reference (projections recomputed
  every decode step, repeat_interleave of the item encodings, ...).
* ``gtm.py``  - the same for GTM_Visuelle2 and Proposed_model v1-v4.
* ``refshim.py`` - three shims that let the *unmodified* reference modules in
  ``/root/reference`` import in this container (build container only).
* ``make_golden.py`` - runs the unmodified reference (through the shims) on
  seeded inputs and writes ``tests/golden/*.pt``.

Pinning: the reference ships no tests / golden vectors / checkpoints
(SURVEY.md section 4), so the oracle is pinned against *outputs of the reference
itself run in the build container* (``tests/golden/*.pt`` + the generating
script).  ``tests/test_oracle_golden.py`` checks every oracle function against
those fixtures; when ``/root/reference`` is mounted it additionally re-runs the
live reference.
"""

"""CPU oracle for the SCONE input-embedding hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker / the timed CPU baseline.  ``scone_b200`` never imports this
package; its hot path fails loudly when the CUDA library is missing.
"""

"""CPU oracle for the waveform -> LLM-embedding path of DeSTA2.5-Audio.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``desta_b200``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
there only as the checker / the CPU arm that is timed beside the GPU path.

Parity pinning: the reference ships no golden vectors for this path (its tests
assert shapes only, ref: tests/test_modeling.py:21-36).  The restatements here
are pinned against outputs of the reference's own code
(``/root/reference/desta/models/modeling_desta25.py`` driving the pinned
third-party arithmetic in ``transformers`` 5.5.0) generated in the build
container by ``oracle/make_golden.py`` and committed under ``tests/golden/``.
"""

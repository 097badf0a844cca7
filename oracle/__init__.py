"""CPU oracle for the CodeFuse / GPT-NeoX hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import it, and there only as the checker / the timed CPU baseline.
The product path (``fastertransformer4codefuse_b200``) never imports this
package and fails loudly when its CUDA library is missing.
"""

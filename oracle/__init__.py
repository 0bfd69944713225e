"""CPU oracle for the SCD clustering-and-naming hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU reference.  The shipped path (``scd_b200``) never imports this package
and fails loudly when its CUDA library is missing.

The reference is research Python/PyTorch, so the oracle is a torch-CPU / NumPy
restatement of the reference's arithmetic, function by function, each citing the
reference ``file:line`` it follows (paths relative to the reference checkout).

Parity pinning: ``oracle/gen_golden.py`` imports the *real* reference modules
(only possible in the build container, where the reference checkout exists), runs
them on seeded inputs and writes the fixtures under ``tests/golden/``.
``tests/test_oracle_golden.py`` then checks every oracle function against those
fixtures, so the oracle is pinned to outputs of the reference itself.  The one
exception is the size-constrained assignment (``sskm_constrained._labels_constrained``):
its arithmetic lives in Google OR-Tools 9.3.10497 which is not installed and not
in the reference tree -> that piece is **parity unpinned** (see DESIGN.md).
"""

from . import kmeans_oracle, naming_oracle, hungarian_oracle  # noqa: F401

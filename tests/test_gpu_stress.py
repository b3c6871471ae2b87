"""Randomised shape/scale sweep of the contrast kernels against the fp64 oracle (GPU box only).

The cases live in tools/gpu_stress.py (ragged N, 1..40 column blocks, unsorted labels, scaled / offset / unnormalised
rows, two-view SupCon); tolerances: loss 1e-4 relative, gradient 5e-3 of its max-abs, and bit-identical results when the
same call is repeated (the kernels use no floating-point atomics)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

CASES = [
    (129, 3, 1, True, 0.3, 1.0, True, 0.1, 0),
    (257, 12, 2, True, 3.0, 0.0, True, 0.5, 0),
    (1153, 15, 3, True, 3.0, 0.0, False, 0.5, 0),
    (2049, 4, 4, True, 1.0, 0.0, True, 0.1, 0),
    (3333, 7, 5, True, 1.0, 0.0, False, 0.07, 0),
    (1000, 3, 6, False, 1.0, 0.0, True, 0.07, 0),
    (2000, 40, 7, False, 1.0, 0.0, False, 0.07, 0),
    (700, 2, 8, True, 1.0, 0.0, True, 0.07, 0),
    (5000, 19, 9, True, 10.0, 0.0, False, 0.07, 0),
    (4096, 64, 10, True, 1.0, 2.0, False, 0.07, 0),
    (256, 8, 12, True, 1.0, 0.0, True, 0.07, 1),
    (33, 3, 13, True, 1.0, 0.0, True, 0.07, 1),
]


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=lambda c: f"n{c[0]}-K{c[1]}-m{c[8]}")
def test_random_case_matches_oracle(c):
    import gpu_stress
    assert gpu_stress.case(*c)

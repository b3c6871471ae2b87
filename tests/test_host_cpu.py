"""CPU-only tests: the C-ABI library loads and exports every declared symbol, and the host-side
sampling plan reproduces the reference's indices when fed counts/rank lookups computed on the CPU."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest
import torch

from oracle import dcl_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as entry
    entry.build()
    from doubly_contrastive_semseg_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "dcl_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(dcl_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 14
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(built_lib.SIGNATURES) == declared
    assert built_lib.load().dcl_version() >= 100
    assert built_lib.workspace_bytes(64, 64) > 0


def test_no_cpu_fallback(built_lib):
    import doubly_contrastive_semseg_b200 as pkg
    crit = pkg.PixelContrastLoss()
    with pytest.raises(built_lib.DclError):
        crit(torch.randn(1, 128, 4, 4), labels=torch.zeros(1, 16, 16, dtype=torch.long),
             predict=torch.randn(1, 19, 4, 4))
    with pytest.raises(built_lib.DclError):
        pkg.contrast_rows(torch.randn(8, 128), torch.zeros(8, dtype=torch.long))


@pytest.mark.parametrize("case", sorted(os.path.basename(p) for p in
                                        glob.glob(os.path.join(GOLDEN, "pixel_*.npz"))))
def test_host_plan_reproduces_reference_indices(case):
    """plan_anchors + layout_rows (host logic of the product) with the device kernels emulated in
    numpy: histogram counts in, (image,label,easy,rank) requests out -> golden pixel indices."""
    from doubly_contrastive_semseg_b200.loss import layout_rows, plan_anchors
    g = dict(np.load(os.path.join(GOLDEN, case)))
    B, C, h, w = g["feats"].shape
    lab = g["lab_ds"].reshape(B, -1).astype(np.int64)
    pred = g["pred"].reshape(B, -1).astype(np.int64)
    counts = np.zeros((B, 256, 2), dtype=np.int64)
    for b in range(B):
        for c in range(256):
            m = lab[b] == c
            counts[b, c, 1] = int((m & (pred[b] == c)).sum())
            counts[b, c, 0] = int(m.sum()) - counts[b, c, 1]
    torch.manual_seed(int(g["call_seed"]))
    plan = plan_anchors(counts, 255, int(g["max_samples"]), int(g["max_views"]))
    A, V = g["pixels"].shape
    assert plan.A == A and plan.n_view == V and np.array_equal(plan.cls, g["y"])
    lay = layout_rows(plan, np.arange(plan.A), 0)
    assert lay.n == A * V and lay.n_pad % 128 == 0 and np.all(np.diff(lay.y[: lay.n]) >= 0)
    got = np.full((A, V), -1, dtype=np.int64)
    for n in range(lay.n):
        b, c, easy, rank = (int(v) for v in lay.req[n])
        sel = np.nonzero((lab[b] == c) & ((pred[b] == c) == bool(easy)))[0]
        v, a = divmod(int(lay.ref_row[n]), A)
        got[a, v] = sel[rank]
    assert np.array_equal(got, g["pixels"])
    assert np.all(lay.req[lay.n:, 0] == -1) and np.all(lay.y[lay.n:] == -1)


def test_c_rng_replay_matches_torch_randperm(built_lib):
    """dcl_host_sample_ranks == the reference's torch.randperm calls, including the generator state
    it leaves behind (so later RNG users see the same stream)."""
    from doubly_contrastive_semseg_b200 import loss as L
    assert L._verify_host_rng() is True
    rng = np.random.default_rng(3)
    for trial in range(4):
        A, n_view = int(rng.integers(1, 40)), int(rng.integers(1, 70))
        nh = rng.integers(0, 3000, A)
        ne = rng.integers(n_view, 3000, A)
        kh = np.minimum(nh, rng.integers(0, n_view + 1, A))
        torch.manual_seed(100 + trial)
        got = L._c_sample_ranks(nh, ne, kh, n_view)
        tail_c = torch.randperm(11)
        torch.manual_seed(100 + trial)
        want = L._torch_sample_ranks(nh, ne, kh, n_view, L._torch_randperm_prefix)
        tail_t = torch.randperm(11)
        assert np.array_equal(got, want)
        assert torch.equal(tail_c, tail_t)


def test_tile_partition_invariants(built_lib):
    """The flat and the exclusive (whole CTAs per unit) partitions of the tile list: every tile belongs to exactly one
    CTA, every CTA of a non-empty grid has work, a unit's CTAs are the contiguous range [first, first + nseg), and no
    unit needs more segment slots than the workspace provides (maxseg)."""
    import ctypes
    from doubly_contrastive_semseg_b200 import _lib
    lib = _lib.load()
    shapes = [(nI, nJ) for nJ in (1, 2, 3, 5, 8, 13, 31, 32, 33, 64, 65, 100, 128, 147, 148, 149, 200, 256, 512, 1024)
              for nI in sorted({1, 2, 3, nJ // 8, nJ // 4, nJ // 2, nJ - 1, nJ} - {0}) if nI <= nJ]
    seen_excl = {0: 0, 1: 0}
    for ctas in (148, 132, 7):
        for nI, nJ in shapes:
            for backward in (0, 1):
                begin = (ctypes.c_longlong * (ctas + 1))()
                units_cap = nI if backward else (nI + 1) // 2
                first = (ctypes.c_int * units_cap)()
                nseg = (ctypes.c_int * units_cap)()
                meta = (ctypes.c_int * 4)()
                assert lib.dcl_debug_partition(nI, nJ, ctas, backward, begin, first, nseg, meta) == 0
                G, excl, maxseg, units = meta[0], meta[1], meta[2], meta[3]
                seen_excl[excl] += 1
                total = units * nJ
                assert units == units_cap and 1 <= G <= ctas and G <= total
                b = [begin[c] for c in range(G + 1)]
                assert b[0] == 0 and b[G] == total
                assert all(b[c] < b[c + 1] for c in range(G)), (nI, nJ, ctas, backward)      # every CTA has >= 1 tile
                for u in range(units):
                    lo, hi = u * nJ, (u + 1) * nJ
                    touching = [c for c in range(G) if b[c] < hi and b[c + 1] > lo]
                    assert touching == list(range(first[u], first[u] + nseg[u])), (nI, nJ, ctas, backward, u)
                    assert nseg[u] <= maxseg, (nI, nJ, ctas, backward, u, nseg[u], maxseg)
                if excl:
                    assert all(b[c] // nJ == (b[c + 1] - 1) // nJ for c in range(G))         # no CTA crosses a unit
    assert seen_excl[0] > 0 and seen_excl[1] > 0

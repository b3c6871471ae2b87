"""world_size-2 gloo test (CPU) of the sharded sampler's host logic: every rank replays the same
RNG stream over the all-gathered histograms and keeps its own anchors; the union must be exactly
the single-process reference sample of the concatenated batch."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dcl_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(B, h, w, K, seed):
    g = np.random.default_rng(seed)
    lab = g.integers(0, K, size=(B, h // 4, w // 4)).repeat(4, 1).repeat(4, 2).reshape(B, h * w)
    lab[g.random((B, h * w)) < 0.05] = 255
    pred = np.where(g.random((B, h * w)) < 0.6, np.minimum(lab, 18), g.integers(0, 19, size=(B, h * w)))
    return lab.astype(np.int64), pred.astype(np.int64)


def _counts(lab, pred):
    B = lab.shape[0]
    c = np.zeros((B, 256, 2), dtype=np.int32)
    for b in range(B):
        for cls in np.unique(lab[b]):
            m = lab[b] == cls
            e = int((m & (pred[b] == cls)).sum())
            c[b, cls] = (int(m.sum()) - e, e)
    return c


def _worker(rank, world, port, B, h, w, K, mv, ms, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from doubly_contrastive_semseg_b200.loss import shard_plan
    lab, pred = _inputs(B, h, w, K, seed=5)
    bl = B // world
    mine = slice(rank * bl, (rank + 1) * bl)
    local = torch.from_numpy(_counts(lab[mine], pred[mine]).reshape(bl, 512))
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    counts_all = torch.cat(gathered).numpy().reshape(B, 256, 2)
    torch.manual_seed(77)                                   # same RNG state on every rank
    sp = shard_plan(counts_all, rank, world, bl, 255, ms, mv)
    lay = sp.layout
    # emulate the select kernel on this rank's images
    pix = np.full(lay.n_pad, -1, dtype=np.int64)
    for n in range(lay.n):
        b, c, easy, rk = (int(v) for v in lay.req[n])
        assert 0 <= b < bl
        sel = np.nonzero((lab[mine][b] == c) & ((pred[mine][b] == c) == bool(easy)))[0]
        pix[n] = sel[rk]
    out[rank] = dict(A=sp.plan.A, V=sp.plan.n_view, n_pad=sp.n_pad, n_global=sp.n_global,
                     ref_row=lay.ref_row.copy(), pix=pix, y=lay.y.copy(), n=lay.n,
                     rows=sp.rows_per_rank.copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B,h,w,K,mv,ms", [(4, 16, 32, 5, 6, 1024), (6, 12, 20, 3, 5, 40)])
def test_shard_plan_union_equals_single_process_reference(B, h, w, K, mv, ms):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, B, h, w, K, mv, ms, out), nprocs=world, join=True)
    lab, pred = _inputs(B, h, w, K, seed=5)
    torch.manual_seed(77)
    plan = O.sample_anchors(lab, pred, 255, ms, mv, O.torch_randperm_prefix)
    r0, r1 = out[0], out[1]
    assert r0["A"] == plan.A and r0["V"] == plan.n_view
    assert r0["n_pad"] == r1["n_pad"] and r0["n_pad"] % 128 == 0
    assert r0["n_global"] == plan.A * plan.n_view == r0["n"] + r1["n"]
    assert list(r0["rows"]) == [r0["n"], r1["n"]]
    got = np.full((plan.A, plan.n_view), -1, dtype=np.int64)
    for r in (r0, r1):
        for n in range(r["n"]):
            v, a = divmod(int(r["ref_row"][n]), plan.A)
            assert got[a, v] == -1
            got[a, v] = r["pix"][n]
            assert r["y"][n] == plan.cls[a]
        assert np.all(np.diff(r["y"][: r["n"]]) >= 0)          # class-sorted inside the rank block
        assert np.all(r["y"][r["n"]:] == -1)
    assert np.array_equal(got, plan.pixels)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("B,h,w,K,mv,ms", [(4, 16, 32, 5, 6, 1024), (8, 12, 20, 3, 5, 40), (4, 64, 64, 7, 48, 2048)])
def test_c_shard_plan_equals_numpy_shard_plan(world, B, h, w, K, mv, ms):
    """dcl_host_plan_rows_sharded (the product's host path) == shard_plan (numpy) for every rank: same plan, same
    local layout, same final generator state; y_all holds every rank's labels; non-local draws are skipped."""
    from doubly_contrastive_semseg_b200.loss import shard_plan, shard_plan_c
    lab, pred = _inputs(B, h, w, K, seed=9)
    counts_all = _counts(lab, pred).reshape(B, 256, 2)
    bl = B // world
    ys = []
    for rank in range(world):
        torch.manual_seed(123)
        ref = shard_plan(counts_all, rank, world, bl, 255, ms, mv)
        end_ref = torch.get_rng_state().clone()
        torch.manual_seed(123)
        got, y_all = shard_plan_c(counts_all, rank, world, bl, 255, ms, mv)
        assert torch.equal(torch.get_rng_state(), end_ref)
        assert (got.plan.A, got.plan.n_view, got.n_pad, got.n_global) == (ref.plan.A, ref.plan.n_view, ref.n_pad, ref.n_global)
        assert np.array_equal(got.rows_per_rank, ref.rows_per_rank)
        for name in ("image", "cls", "num_hard", "num_easy", "keep_hard"):
            assert np.array_equal(getattr(got.plan, name), getattr(ref.plan, name)), name
        mine = (ref.plan.image // bl) == rank
        assert np.array_equal(got.plan.ranks[mine], ref.plan.ranks[mine])
        for name in ("req", "y", "ref_row", "anchor"):
            assert np.array_equal(getattr(got.layout, name), getattr(ref.layout, name)), name
        assert got.layout.n == ref.layout.n
        assert np.array_equal(y_all[rank * ref.n_pad:(rank + 1) * ref.n_pad], ref.layout.y)
        ys.append(y_all.copy())
        # compact form (the product's path): labels right behind the requests in one caller-owned buffer
        cap = max(128, (ms + 127) // 128 * 128) + 128
        stage = np.full((4 + world) * cap, -7, dtype=np.int32)
        torch.manual_seed(123)
        got2, y2 = shard_plan_c(counts_all, rank, world, bl, 255, ms, mv, stage, None)
        assert torch.equal(torch.get_rng_state(), end_ref)
        assert np.array_equal(got2.layout.req, ref.layout.req) and np.array_equal(y2, y_all)
        assert np.shares_memory(y2, stage) and np.array_equal(stage[4 * ref.n_pad:(4 + world) * ref.n_pad], y_all)
    for y in ys[1:]:
        assert np.array_equal(y, ys[0])


def _consecutive_plans_check(world):
    from doubly_contrastive_semseg_b200 import _lib
    from doubly_contrastive_semseg_b200.loss import shard_plan, shard_plan_c
    B, h, w, K, mv, ms = 8, 96, 128, 9, 40, 4096
    lab, pred = _inputs(B, h, w, K, seed=21)
    counts_all = _counts(lab, pred).reshape(B, 256, 2)
    bl = B // world
    rank = world - 1
    steps = 6
    torch.manual_seed(77)
    refs = [shard_plan(counts_all, rank, world, bl, 255, ms, mv) for _ in range(steps)]
    end_ref = torch.get_rng_state().clone()
    stats0 = (ctypes.c_longlong * 4)()
    _lib.load().dcl_host_lookahead_stats(stats0)
    torch.manual_seed(77)
    for ref in refs:
        got, _ = shard_plan_c(counts_all, rank, world, bl, 255, ms, mv)
        mine = (ref.plan.image // bl) == rank
        assert np.array_equal(got.plan.ranks[mine], ref.plan.ranks[mine])
        assert np.array_equal(got.layout.req, ref.layout.req)
        assert np.array_equal(got.layout.ref_row, ref.layout.ref_row)
    assert torch.equal(torch.get_rng_state(), end_ref)
    stats1 = (ctypes.c_longlong * 4)()
    _lib.load().dcl_host_lookahead_stats(stats1)
    if os.environ.get("DCL_HOST_LOOKAHEAD", "1") != "0":
        assert stats1[0] - stats0[0] >= steps - 2          # all but the first plans after the seed came from the stream


@pytest.mark.parametrize("world", [1, 2])
@pytest.mark.parametrize("threads", ["", "3"])
def test_consecutive_plans_use_the_lookahead_stream_and_stay_exact(world, threads):
    """Plans that follow each other without any other use of the generator run from the look-ahead stream, their
    permutations replayed in parallel on large problems (anchor positions in the stream are known up front; here
    forced with DCL_HOST_THREADS, which the library reads once per process, hence the child process); every plan must
    still equal the numpy / torch.randperm plan, including the generator state it leaves behind."""
    import subprocess
    import sys
    env = dict(os.environ)
    env.pop("DCL_HOST_THREADS", None)
    if threads:
        env["DCL_HOST_THREADS"] = threads
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path[:0] = [%r, %r]; from test_sharded_cpu import _consecutive_plans_check as c; c(%d)"
            % (root, os.path.join(root, "tests"), world))
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_c_plan_equals_numpy_plan_on_random_tables():
    """dcl_host_plan_rows(_sharded) == the numpy / torch.randperm plan on random count tables: classes below the
    max_views threshold, hard-only / easy-only classes, capped n_view, single image, many ranks."""
    from doubly_contrastive_semseg_b200.loss import shard_plan, shard_plan_c
    rng = np.random.default_rng(2024)
    done = 0
    for trial in range(40):
        world = int(rng.choice([1, 1, 2, 3, 4]))
        bl = int(rng.integers(1, 4))
        B = world * bl
        K = int(rng.integers(1, 12))
        mv = int(rng.integers(1, 70))
        ms = int(rng.choice([64, 200, 1024, 4096]))
        counts = np.zeros((B, 256, 2), dtype=np.int32)
        for b in range(B):
            for c in rng.choice(19, size=K, replace=False):
                kind = rng.integers(0, 4)
                nh = int(rng.integers(0, 3000)) if kind != 1 else 0
                ne = int(rng.integers(0, 3000)) if kind != 2 else 0
                if kind == 3:
                    nh, ne = int(rng.integers(0, mv + 2)), int(rng.integers(0, mv + 2))     # around the threshold
                counts[b, c] = (nh, ne)
        counts[:, 255] = rng.integers(0, 500, size=(B, 2))                                    # ignored label
        rank = int(rng.integers(0, world))
        torch.manual_seed(1000 + trial)
        try:
            ref = shard_plan(counts, rank, world, bl, 255, ms, mv)
        except Exception:                      # the reference's unreachable split branch: both sides must agree
            torch.manual_seed(1000 + trial)
            with pytest.raises(Exception):
                shard_plan_c(counts, rank, world, bl, 255, ms, mv)
            continue
        end_ref = torch.get_rng_state().clone()
        torch.manual_seed(1000 + trial)
        got = shard_plan_c(counts, rank, world, bl, 255, ms, mv)
        assert (ref is None) == (got is None)
        assert torch.equal(torch.get_rng_state(), end_ref)
        if ref is None:
            continue
        got, y_all = got
        assert (got.plan.A, got.plan.n_view, got.n_pad, got.n_global) == (ref.plan.A, ref.plan.n_view, ref.n_pad, ref.n_global)
        if ref.plan.n_view == 0:
            continue
        mine = (ref.plan.image // bl) == rank
        assert np.array_equal(got.plan.ranks[mine], ref.plan.ranks[mine])
        for name in ("req", "y", "ref_row", "anchor"):
            assert np.array_equal(getattr(got.layout, name), getattr(ref.layout, name)), (trial, name)
        done += 1
    assert done >= 15


# ------------------------------------------------------------------ device-mode plan (k_plan emulated in numpy)
def _temper(y):
    y = y.astype(np.uint32)
    y ^= y >> np.uint32(11)
    y ^= (y << np.uint32(7)) & np.uint32(0x9D2C5680)
    y ^= (y << np.uint32(15)) & np.uint32(0xEFC60000)
    y ^= y >> np.uint32(18)
    return y


def _emulate_k_plan(anchors, n_local, ring, ring_blocks, ycls, ycnt, yoff, world, rank, n_view, n_pad):
    """What csrc/dcl_plan.cu computes, restated with numpy from the same descriptors: the kept prefix of every local
    permutation out of the mirrored generator stream -> row requests; labels of every rank block."""
    req = np.full((n_pad, 4), -9, dtype=np.int32)
    y_all = np.full(world * n_pad, -9, dtype=np.int32)
    for i in range(n_local):
        a = anchors[i]
        for easy in (0, 1):
            n = int(a["num_easy"] if easy else a["num_hard"])
            k = int(a["keep_easy"] if easy else a["keep_hard"])
            g = int(a["g_easy"] if easy else a["g_hard"])
            row = int(a["row0"]) + (int(a["keep_hard"]) if easy else 0)
            moved = {}
            for s in range(k):
                j = s
                if s < n - 1:
                    gi = g + s
                    blk, word = divmod(gi, 624)
                    x = int(_temper(ring[(blk % ring_blocks) * 624 + word: (blk % ring_blocks) * 624 + word + 1])[0])
                    j = s + x % (n - s)
                vi, vj = moved.get(s, s), moved.get(j, j)
                moved[j] = vi
                req[row + s] = (int(a["image"]), int(a["cls"]), easy, vj)
    for r in range(world):
        for i in range(n_pad):
            o = i // n_view
            valid = o < ycnt[r]
            y_all[r * n_pad + i] = ycls[yoff[r] + o] if valid else -1
            if not valid and r == rank:
                req[i] = (-1, -1, -1, -1)
    return req, y_all


_PLAN_ANCHOR = np.dtype([("g_hard", np.uint64), ("g_easy", np.uint64), ("num_hard", np.int32), ("num_easy", np.int32),
                         ("keep_hard", np.int32), ("keep_easy", np.int32), ("row0", np.int32), ("image", np.int32),
                         ("cls", np.int32), ("reserved", np.int32)])


def _device_plan_check(world):
    """Consecutive plans requested the way dcl_step_fwd requests them: once the look-ahead stream runs, the host only
    places the permutations in the stream (device mode); replaying them from the ring the way k_plan does must give
    the numpy / torch.randperm plan bit for bit, and the generator must end where the reference leaves it."""
    from doubly_contrastive_semseg_b200 import _lib
    from doubly_contrastive_semseg_b200.loss import shard_plan
    assert _PLAN_ANCHOR.itemsize == 48
    B, h, w, K, mv, ms = 8, 96, 128, 9, 40, 4096
    lab, pred = _inputs(B, h, w, K, seed=33)
    counts_all = np.ascontiguousarray(_counts(lab, pred).reshape(B, 256, 2))
    bl = B // world
    lib = _lib.load()
    steps = 5
    for rank in sorted({0, world - 1}):
        torch.manual_seed(4242 + rank)
        refs = [shard_plan(counts_all, rank, world, bl, 255, ms, mv) for _ in range(steps)]
        end_ref = torch.get_rng_state().clone()
        torch.manual_seed(4242 + rank)
        taken = 0
        for ref in refs:
            cap = max(128, (ms + 127) // 128 * 128) + 128
            info = np.zeros(8, dtype=np.int32)
            an = np.empty((5, B * 256), dtype=np.int64)
            ranks = np.zeros(cap, dtype=np.int64)
            req = np.full(cap * 4, -5, dtype=np.int32)
            y_all = np.full(world * cap, -5, dtype=np.int32)
            anchors = np.zeros(bl * 256, dtype=_PLAN_ANCHOR)
            ycls = np.zeros(B * 256, dtype=np.int32)
            yanchor = np.zeros(B * 256, dtype=np.int32)
            ycnt = np.zeros(world, dtype=np.int32)
            yoff = np.zeros(world, dtype=np.int32)
            meta = (ctypes.c_longlong * 8)()
            ring_ptr = ctypes.c_void_p()
            st = torch.get_rng_state()
            sbuf = st.numpy()
            rc = lib.dcl_debug_plan_device(
                counts_all.ctypes.data, bl, world, rank, 255, ms, mv, sbuf.ctypes.data, sbuf.nbytes, info.ctypes.data,
                an[0].ctypes.data, an[1].ctypes.data, an[2].ctypes.data, an[3].ctypes.data, an[4].ctypes.data,
                ranks.ctypes.data, req.ctypes.data, y_all.ctypes.data, anchors.ctypes.data, ycls.ctypes.data,
                yanchor.ctypes.data, ycnt.ctypes.data, yoff.ctypes.data, ctypes.cast(meta, ctypes.c_void_p),
                ctypes.cast(ctypes.byref(ring_ptr), ctypes.c_void_p))
            assert rc == 0
            torch.set_rng_state(st)
            A, n_view, n, n_pad, n_global = (int(v) for v in info[:5])
            assert (A, n_view, n_pad, n_global) == (ref.plan.A, ref.plan.n_view, ref.n_pad, ref.n_global)
            if meta[0]:
                taken += 1
                ring_blocks = int(meta[5])
                ring = np.ctypeslib.as_array(ctypes.cast(ring_ptr, ctypes.POINTER(ctypes.c_uint32)),
                                             shape=(ring_blocks * 624,))
                assert int(meta[3]) <= int(meta[4]) < int(meta[6])
                got_req, got_y = _emulate_k_plan(anchors, int(meta[1]), ring, ring_blocks, ycls, ycnt, yoff, world, rank,
                                                 n_view, n_pad)
            else:
                got_req = req[: n_pad * 4].reshape(n_pad, 4)
                got_y = y_all[: world * n_pad]
            assert np.array_equal(got_req, ref.layout.req)
            assert np.array_equal(got_y[rank * n_pad:(rank + 1) * n_pad], ref.layout.y)
            # the sorted order handed to the caller reproduces the reference row of every device row
            order = yanchor[int(yoff[rank]): int(yoff[rank]) + int(ycnt[rank])]
            assert np.array_equal(np.repeat(order, n_view), ref.layout.anchor[: ref.layout.n])
        assert torch.equal(torch.get_rng_state(), end_ref)
        if os.environ.get("DCL_HOST_LOOKAHEAD", "1") != "0":
            assert taken >= steps - 2, taken


@pytest.mark.parametrize("world", [1, 2])
def test_device_mode_plan_equals_numpy_plan(world):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path[:0] = [%r, %r]; from test_sharded_cpu import _device_plan_check as c; c(%d)"
            % (root, os.path.join(root, "tests"), world))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr

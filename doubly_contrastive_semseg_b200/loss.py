"""Drop-in loss modules: same names, constructor arguments, attributes, forward signatures and
error behaviour as the reference's `utils/loss.py` (`PixelContrastLoss` :250-415, `SupConLoss`
:84-205), with all device work done by libdcl_b200.so (hand-written sm_100a CUDA, C ABI in
include/dcl_b200.h).  PyTorch is used for memory, streams, autograd plumbing and (for the
image-level term only) the two tiny projection GEMMs.  No CPU path exists here.
"""
from __future__ import annotations

import ctypes
import os
import time
from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib

MODE_PIXEL, MODE_SUPCON = 0, 1
_DIM, _TILE = 128, 128
_CHUNK, _BINS = 2048, 512


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


_STREAM_CACHE = {}


def _stream():
    """cudaStream_t of torch's current stream (torch.cuda.current_stream() costs ~5 us per call: the raw handle is
    cached per stream id)."""
    sid = torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    p = _STREAM_CACHE.get(sid)
    if p is None:
        p = _STREAM_CACHE[sid] = ctypes.c_void_p(sid)
    return p


_LAUNCHES = 0          # kernels of ours launched since reset (bench.py's `gpu_launches`)
_PROFILE_HOOK = None   # callable(name, start_event, end_event) around the similarity calls


def launch_count() -> int:
    return _LAUNCHES


def reset_launch_count():
    global _LAUNCHES
    _LAUNCHES = 0


def set_profile_hook(fn):
    global _PROFILE_HOOK
    _PROFILE_HOOK = fn


def set_sim_timing(on: bool) -> None:
    """Measurement aid: have dcl_step_fwd bracket its contrast forward / backward with events owned by the library
    (current device; cheaper than `set_profile_hook`, which creates and records four events per step from Python)."""
    rc = _lib.load().dcl_step_sim_timing(1 if on else 0)
    if rc < 0:
        raise _lib.DclError("dcl_step_sim_timing failed: %s" % _lib.load().dcl_last_error().decode("utf-8", "replace"))


def sim_times():
    """-> (forward ms, backward ms, steps) summed over the steps recorded since `set_sim_timing(True)`; synchronize first."""
    f, b, n = ctypes.c_double(0.0), ctypes.c_double(0.0), ctypes.c_longlong(0)
    _lib.call("dcl_step_sim_elapsed", ctypes.byref(f), ctypes.byref(b), ctypes.byref(n))
    return f.value, b.value, int(n.value)


def _count(n):
    global _LAUNCHES
    _LAUNCHES += n


class _Timed:
    """Records a CUDA-event pair on the current stream around a block when a hook is installed."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _PROFILE_HOOK is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _PROFILE_HOOK is not None:
            self.b.record()
            _PROFILE_HOOK(self.name, self.a, self.b)


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.DclError("%s must be a CUDA tensor: this library has no CPU path" % name)


def _on_device(t: torch.Tensor, dev: torch.device, name: str):
    """Side inputs (labels, predict, class_labels) must live where the embeddings live: the kernels take raw
    pointers, so a CPU tensor or one on another GPU would be an illegal address, not a Python exception."""
    if not t.is_cuda:
        raise _lib.DclError("%s must be a CUDA tensor on %s: this library has no CPU path" % (name, dev))
    if t.device != dev:
        raise _lib.DclError("%s is on %s but the embeddings are on %s" % (name, t.device, dev))
    return t


# ----------------------------------------------------------------------------------------------
# host-side sampling plan (the part of loss.py:264-337 that is Python control flow + CPU RNG)
# ----------------------------------------------------------------------------------------------
@dataclass
class AnchorPlan:
    A: int                 # anchors = (image, class) pairs == reference `total_classes`
    n_view: int
    image: np.ndarray      # [A] global image index, reference order (image asc, class asc)
    cls: np.ndarray        # [A]
    num_hard: np.ndarray   # [A]
    num_easy: np.ndarray   # [A]
    keep_hard: np.ndarray  # [A]
    ranks: np.ndarray      # [A, n_view] rank inside the hard (v < keep_hard) or easy list


def _torch_randperm_prefix(n: int, k: int) -> np.ndarray:
    # consumes the global CPU generator exactly like loss.py:327,329
    return torch.randperm(n)[:k].numpy()


_HOST_RNG_OK: Optional[bool] = None


def _c_sample_ranks(nh: np.ndarray, ne: np.ndarray, kh: np.ndarray, n_view: int) -> np.ndarray:
    """All randperm prefixes of one call in C, on torch's serialized global CPU generator state
    (csrc/dcl_host_rng.cpp); the generator ends exactly where the reference's calls would leave it."""
    st = torch.get_rng_state()
    buf = st.numpy()
    A = int(nh.shape[0])
    out = np.empty((A, n_view), dtype=np.int64)
    nh, ne, kh = (np.ascontiguousarray(v, dtype=np.int64) for v in (nh, ne, kh))
    _lib.call("dcl_host_sample_ranks", buf.ctypes.data, buf.nbytes, A, n_view, nh.ctypes.data,
              ne.ctypes.data, kh.ctypes.data, out.ctypes.data)
    torch.set_rng_state(st)
    return out


def _torch_sample_ranks(nh, ne, kh, n_view, randperm) -> np.ndarray:
    A = int(nh.shape[0])
    out = np.zeros((A, n_view), dtype=np.int64)
    for a in range(A):
        k_h = int(kh[a])
        out[a, :k_h] = randperm(int(nh[a]), k_h)               # hard first, loss.py:327-328
        out[a, k_h:] = randperm(int(ne[a]), n_view - k_h)      # then easy, loss.py:329-330
    return out


def _verify_host_rng() -> bool:
    """One-time check that the C replay and torch.randperm agree on this torch build (results and
    final generator state).  The global generator is restored afterwards."""
    global _HOST_RNG_OK
    if _HOST_RNG_OK is not None:
        return _HOST_RNG_OK
    saved = torch.get_rng_state()
    try:
        nh = np.array([0, 1, 700, 5, 1300, 2], dtype=np.int64)
        ne = np.array([9, 640, 3, 2000, 1, 625], dtype=np.int64)
        kh = np.array([0, 1, 3, 3, 5, 2], dtype=np.int64)
        ne_keep = 6 - kh
        assert np.all(ne_keep <= ne)
        torch.default_generator.manual_seed(20221118)
        got = _c_sample_ranks(nh, ne, kh, 6)
        end_c = torch.get_rng_state().clone()
        torch.default_generator.manual_seed(20221118)
        want = _torch_sample_ranks(nh, ne, kh, 6, _torch_randperm_prefix)
        end_t = torch.get_rng_state()
        _HOST_RNG_OK = bool(np.array_equal(got, want) and torch.equal(end_c, end_t))
    except Exception:
        _HOST_RNG_OK = False
    finally:
        torch.set_rng_state(saved)
    return _HOST_RNG_OK


def plan_anchors(counts: np.ndarray, ignore_label: int, max_samples: int, max_views: int,
                 randperm: Optional[Callable[[int, int], np.ndarray]] = None) -> Optional[AnchorPlan]:
    """counts [B,256,2] = pixels per (image, label, hard|easy).  Returns None when no class
    qualifies (reference: `return None, None`, loss.py:287-288).  With `randperm=None` the draws
    come from torch's global CPU generator (C replay when verified, torch.randperm otherwise)."""
    tot = counts.sum(axis=2)
    keep = tot > max_views                                    # loss.py:282
    if 0 <= ignore_label <= 255:
        keep[:, ignore_label] = False                         # loss.py:281
    img, cls = np.nonzero(keep)                               # image asc, class asc
    A = int(img.shape[0])
    if A == 0:
        return None
    n_view = min(max_samples // A, max_views)                 # loss.py:290-291
    nh = counts[img, cls, 0].astype(np.int64)
    ne = counts[img, cls, 1].astype(np.int64)
    # split rule, loss.py:314-325 (true division on n_view)
    half = n_view / 2
    c1 = (nh >= half) & (ne >= half)
    c2 = ~c1 & (nh >= half)
    c3 = ~c1 & ~c2 & (ne >= half)
    if not bool(np.all(c1 | c2 | c3)):
        bad = int(np.nonzero(~(c1 | c2 | c3))[0][0])
        print("this shoud be never touched! {} {} {}".format(int(nh[bad]), int(ne[bad]), n_view))
        raise Exception
    kh = np.where(c1, n_view // 2, np.where(c2, n_view - ne, nh)).astype(np.int64)
    if n_view <= 0:
        ranks = np.zeros((A, 0), dtype=np.int64)
    elif randperm is None and _verify_host_rng():
        ranks = _c_sample_ranks(nh, ne, kh, n_view)
    else:
        ranks = _torch_sample_ranks(nh, ne, kh, n_view, randperm or _torch_randperm_prefix)
    return AnchorPlan(A, n_view, img.astype(np.int64), cls.astype(np.int64), nh, ne, kh, ranks)


@dataclass
class RowLayout:
    """Device row order: anchors stably sorted by class (positives become block-diagonal), views
    contiguous per anchor, padded to a multiple of 128 rows."""
    n: int
    n_pad: int
    req: np.ndarray        # [n_pad, 4] int32 (local image, label, easy, rank); image -1 = padding
    y: np.ndarray          # [n_pad] int32, -1 = padding
    ref_row: np.ndarray    # [n_pad] row index v*A + a in the reference's ordering, -1 = padding
    anchor: np.ndarray     # [n_pad] anchor id a (reference order), -1 = padding


def layout_rows(plan: AnchorPlan, anchors: np.ndarray, image_offset: int, n_pad: Optional[int] = None
                ) -> RowLayout:
    """Rows of the anchors listed in `anchors` (indices into the plan), class-sorted."""
    V = plan.n_view
    order = anchors[np.argsort(plan.cls[anchors], kind="stable")]
    n = int(order.shape[0]) * V
    if n_pad is None:
        n_pad = max(_TILE, (n + _TILE - 1) // _TILE * _TILE)
    req = np.full((n_pad, 4), -1, dtype=np.int32)
    y = np.full(n_pad, -1, dtype=np.int32)
    ref_row = np.full(n_pad, -1, dtype=np.int64)
    anchor = np.full(n_pad, -1, dtype=np.int64)
    if n:
        a_rep = np.repeat(order, V)
        v_rep = np.tile(np.arange(V), order.shape[0])
        req[:n, 0] = plan.image[a_rep] - image_offset
        req[:n, 1] = plan.cls[a_rep]
        req[:n, 2] = (v_rep >= plan.keep_hard[a_rep]).astype(np.int32)
        req[:n, 3] = plan.ranks[a_rep, v_rep]
        y[:n] = plan.cls[a_rep]
        ref_row[:n] = v_rep * plan.A + a_rep
        anchor[:n] = a_rep
    return RowLayout(n, n_pad, req, y, ref_row, anchor)


# ----------------------------------------------------------------------------------------------
# thin wrappers over the C ABI
# ----------------------------------------------------------------------------------------------
def classify(labels: torch.Tensor, predict: torch.Tensor, h: int, w: int):
    """-> code [B,hw] u16 (as int16 storage), chunk_hist [B,512,n_chunks] i32, counts [B,512] i32"""
    B, H, W = labels.shape
    C = predict.shape[1]
    hw = h * w
    n_chunks = (hw + _CHUNK - 1) // _CHUNK
    dev = labels.device
    code = torch.empty((B, hw), dtype=torch.int16, device=dev)
    chunk = torch.empty((B, _BINS, n_chunks), dtype=torch.int32, device=dev)
    counts = torch.empty((B, _BINS), dtype=torch.int32, device=dev)
    _lib.call("dcl_sample_classify", _p(labels), _p(predict), B, H, W, h, w, C, _p(code), _p(chunk),
              _p(counts), _stream())
    _count(1)
    return code, chunk, counts


def select_pixels(code, chunk, B, hw, req_dev, n_rows, rowof=None):
    """rank -> pixel; `rowof` [B*hw] i32 (optional) receives the inverse map pixel -> row (-1 elsewhere)"""
    pix = torch.empty(n_rows, dtype=torch.int32, device=code.device)
    _lib.call("dcl_sample_select", _p(code), _p(chunk), B, hw, _p(req_dev), n_rows, _p(pix), _p(rowof), _stream())
    _count(1)
    return pix


def gather_tiles(feats, pix, n_pad, rowof=None):
    B, C, h, w = feats.shape
    tiles = torch.empty(n_pad * _DIM * 2, dtype=torch.uint8, device=feats.device)
    sqnorm = torch.empty(n_pad, dtype=torch.float32, device=feats.device)
    _lib.call("dcl_gather_tiles", _p(feats), B, h * w, _p(pix), n_pad, _p(tiles), _p(sqnorm), _p(rowof), _stream())
    _count(1 if rowof is None else 2)
    return tiles, sqnorm


def pack_rows(Z, n_pad):
    n = Z.shape[0]
    tiles = torch.empty(n_pad * _DIM * 2, dtype=torch.uint8, device=Z.device)
    sqnorm = torch.empty(n_pad, dtype=torch.float32, device=Z.device)
    _lib.call("dcl_pack_rows", _p(Z), n, n_pad, _p(tiles), _p(sqnorm), _stream())
    _count(1)
    return tiles, sqnorm


def contrast_forward(tiles, y, sqnorm, nJ, rb0, nI, n_valid, mode, T, Tb, colA=None, colB=None):
    """Forward sweeps for local row blocks [rb0, rb0+nI).  Returns (colA, colB, rowloss, loss_sum)."""
    dev = tiles.device
    if colA is None:
        colA = torch.empty((nJ * _TILE, 4), dtype=torch.float32, device=dev)
        colB = torch.empty((nJ * _TILE, 4), dtype=torch.float32, device=dev)
    rowloss = torch.empty(nJ * _TILE, dtype=torch.float32, device=dev)
    loss_sum = torch.empty(2, dtype=torch.float32, device=dev)
    nbytes = _lib.workspace_bytes(nI, nJ)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with _Timed("contrast_fwd"):
        _lib.call("dcl_contrast_fwd", _p(tiles), _p(y), _p(sqnorm), nJ, rb0, nI, n_valid, mode, float(T),
                  float(Tb), _p(ws), nbytes, _p(colA), _p(colB), _p(rowloss), _p(loss_sum), _stream())
    _count(_lib.contrast_launches(mode, 0))
    return colA, colB, rowloss, loss_sum


def contrast_backward(tiles, y, colA, colB, nJ, rb0, nI, mode):
    dev = tiles.device
    dF = torch.empty((nI * _TILE, _DIM), dtype=torch.float32, device=dev)
    nbytes = _lib.workspace_bytes(nI, nJ)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with _Timed("contrast_bwd"):
        _lib.call("dcl_contrast_bwd", _p(tiles), _p(y), _p(colA), _p(colB), nJ, rb0, nI, mode, _p(ws), nbytes,
                  _p(dF), _stream())
    _count(_lib.contrast_launches(mode, 1))
    return dF


# ----------------------------------------------------------------------------------------------
# the one-call step (dcl_step_fwd / dcl_step_bwd): cached sizes, persistent scratch, raw pointers
# ----------------------------------------------------------------------------------------------
_POOL_AFTER_CLASSIFY = os.environ.get("DCL_POOL_AFTER_CLASSIFY", "1") != "0"   # doubly step: pool issued behind the step's classify
_FUSED_HEAD = os.environ.get("DCL_FUSED_HEAD", "1") != "0"       # image-level head: fused kernels (0: torch MLP + glue)
_DEBUG_PY_TIMES = [] if os.environ.get("DCL_DEBUG_PY_TIMES") else None      # diagnostics: host milliseconds of _run_step's parts
_FUSED_STEP = os.environ.get("DCL_FUSED_STEP", "1") != "0"     # 0: the stage-by-stage Python path (same results)
_DEVICE_PLAN = os.environ.get("DCL_DEVICE_PLAN", "1") != "0"   # 0: always replay the generator on the host (same results)
_SIDE_STREAM_FILL = os.environ.get("DCL_SIDE_STREAM_FILL", "1") != "0"
_WS_BYTES = {}
_WS_CACHE = {}
_LAUNCHES_CACHE = {}
_SIDE_STREAMS = {}


def _ws_bytes(nI, nJ):
    v = _WS_BYTES.get((nI, nJ))
    if v is None:
        v = _WS_BYTES[(nI, nJ)] = _lib.workspace_bytes(nI, nJ)
    return v


def _workspace(dev, nbytes):
    """Scratch for one contrast call.  Calls on a device are stream-ordered, so one buffer per device is reused
    (grown when needed) instead of a fresh allocation per call."""
    key = (dev.index, torch._C._cuda_getCurrentRawStream(dev.index))
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = _WS_CACHE[key] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    return ws


def _launches(mode, backward):
    v = _LAUNCHES_CACHE.get((mode, backward))
    if v is None:
        v = _LAUNCHES_CACHE[(mode, backward)] = _lib.contrast_launches(mode, backward)
    return v


def _carve(dev, sizes):
    """One uint8 tensor holding consecutive 256-byte-aligned regions of the given byte sizes -> (tensor, pointers)."""
    offs, o = [], 0
    for sz in sizes:
        offs.append(o)
        o += (sz + 255) // 256 * 256
    buf = torch.empty(max(o, 256), dtype=torch.uint8, device=dev)
    base = buf.data_ptr()
    return buf, [ctypes.c_void_p(base + x) for x in offs]


def _side_stream(dev):
    """Second stream of a device: the dense gradient buffer is cleared there, next to the step's kernels."""
    s = _SIDE_STREAMS.get(dev.index)
    if s is None:
        s = _SIDE_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    return s


_WS_CAP_BYTES = {}


def _ws_cap_bytes(cap, world):
    """Workspace that serves every row count up to `cap` per rank (the row count is only known after the plan)."""
    v = _WS_CAP_BYTES.get((cap, world))
    if v is None:
        v = _WS_CAP_BYTES[(cap, world)] = max(_ws_bytes(n, n * world) for n in range(1, cap // _TILE + 1))
    return v


_COMMS = {}


def _shard_comm(group):
    """(world, rank, ncclComm_t handle) of a torch.distributed group: the library talks to NCCL itself on the step's
    stream (csrc/dcl_comm.cpp); torch.distributed only carries the 128-byte id, once per group."""
    import torch.distributed as dist
    key = id(group) if group is not None else 0
    ent = _COMMS.get(key)
    if ent is None:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        buf = (ctypes.c_ubyte * 128)()
        if rank == 0:
            _lib.call("dcl_comm_unique_id", buf)
        box = [bytes(buf) if rank == 0 else None]
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast_object_list(box, src=src, group=group)
        comm = ctypes.c_void_p()
        _lib.call("dcl_comm_init", box[0], world, rank, ctypes.byref(comm))
        ent = _COMMS[key] = (world, rank, comm, group)
    return ent


_P2P = os.environ.get("DCL_P2P", "1") != "0"        # 0: the sharded step's exchanges go through NCCL


def _peer_arena(group, world, rank, B, cap):
    """Peer-memory arena of a sharded step (csrc/dcl_p2p.cu): created by every rank, IPC handles all-gathered once
    through torch.distributed, peers mapped.  Returns the handle, or None when peer mapping is not possible (the
    step then uses NCCL); the decision is taken collectively so that all ranks agree."""
    import torch.distributed as dist
    handle = ctypes.c_void_p()
    ipc = (ctypes.c_ubyte * 64)()
    lib = _lib.load()
    ok = _P2P and world <= 16 and lib.dcl_p2p_create(world, rank, B, cap, ctypes.byref(handle), ipc) == 0
    box = [None] * world
    dist.all_gather_object(box, bytes(ipc) if ok else None, group=group)
    if ok and all(b is not None for b in box):
        ok = lib.dcl_p2p_open(handle, b"".join(box)) == 0
    else:
        ok = False
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok), group=group)
    if all(flags):
        return handle
    if handle.value:
        lib.dcl_p2p_destroy(handle)
    return None


class _StepResult:
    """What one dcl_step_fwd left behind: the loss, and for the backward the sampled pixels and the eager dF."""
    __slots__ = ("loss", "keep", "p_pix", "p_dF", "p_rowof", "n_pad", "n", "n_global", "empty", "dzero", "shape", "info",
                 "finish")


def _run_step(crit, feats, labels, predict, shard, want_grad, zero_fill, after_begin=None, defer=False):
    """Issue one pixel-term step on feats [B,128,h,w] (contiguous f32) through dcl_step_fwd.  `shard` is None or
    (world, rank, comm).  `zero_fill`: allocate and clear the dense gradient buffer next to the step.
    `after_begin`: split issue - classify and the count-table copy go first (dcl_step_begin), then the callback's
    launches, then the rest.  `defer` (with `after_begin`): return after the callback with `res.loss` allocated and
    `res.finish` set; the caller issues more work and then calls `res.finish()` exactly once (the rest of the step:
    the wait for the count table, plan, select, gather, contrast), which completes `res`."""
    B, C, h, w = feats.shape
    dev = feats.device
    hw = h * w
    world, rank, comm, group = shard if shard is not None else (1, 0, None, None)
    _t = [time.perf_counter()] if _DEBUG_PY_TIMES is not None else None
    sb = crit._step_buffers(B, h, w, dev, world)
    cap = sb["cap"]
    res = _StepResult()
    res.shape = (B, C, h, w)
    res.dzero = torch.empty_like(feats) if (want_grad and zero_fill) else None
    if _t is not None:
        _t.append(time.perf_counter())
    # what outlives the call: sampled pixels, eager gradient of the rows, the loss
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    res.keep, (p_pix, p_rowof, p_dF) = _carve(dev, (cap * 4, B * hw * 4, cap * _DIM * 4 if want_grad else 0))
    if _t is not None:
        _t.append(time.perf_counter())
    st = torch.get_rng_state()
    sbuf = st.numpy()
    step = sb["step"]
    step.labels, step.predict, step.feats = labels.data_ptr(), predict.data_ptr(), feats.data_ptr()
    step.H, step.W, step.C_cls = labels.shape[1], labels.shape[2], predict.shape[1]
    step.ignore_label, step.max_samples, step.max_views = int(crit.ignore_label), int(crit.max_samples), int(crit.max_views)
    step.temperature, step.base_temperature = float(crit.temperature), float(crit.base_temperature)
    step.torch_rng_state, step.state_bytes = sbuf.ctypes.data, sbuf.nbytes
    step.rank, step.comm = rank, comm
    if world > 1:
        if "p2p" not in sb:
            sb["p2p"] = _peer_arena(group, world, rank, B, cap)
        step.p2p = sb["p2p"]
    else:
        step.p2p = None
    step.pix, step.rowof, step.dF, step.loss = p_pix, p_rowof, (p_dF if want_grad else None), loss.data_ptr()
    ws = _workspace(dev, sb["ws_bytes"])
    step.workspace, step.workspace_bytes = ws.data_ptr(), sb["ws_bytes"]
    if res.dzero is not None:
        step.zero_fill, step.zero_fill_bytes = res.dzero.data_ptr(), res.dzero.numel() * 4
        step.side_stream = _side_stream(dev).cuda_stream if _SIDE_STREAM_FILL else None
    else:
        step.zero_fill, step.zero_fill_bytes, step.side_stream = None, 0, None
    step.device_plan = 1 if _DEVICE_PLAN else 0
    ev = None
    if _PROFILE_HOOK is not None:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for e in ev:
            e.record()                                    # materialise the handles; the library re-records them
        step.ev_fwd_begin, step.ev_fwd_end = ev[0].cuda_event, ev[1].cuda_event
        step.ev_bwd_begin, step.ev_bwd_end = (ev[2].cuda_event, ev[3].cuda_event) if want_grad else (None, None)
    else:
        step.ev_fwd_begin = step.ev_fwd_end = step.ev_bwd_begin = step.ev_bwd_end = None
    lib = _lib.load()
    if _t is not None:
        _t.append(time.perf_counter())
    if after_begin is not None:
        # split issue: classify + count-table copy first, the caller's launches behind them, then the rest of the step
        rc0 = lib.dcl_step_begin(ctypes.byref(step), _stream())
        if rc0 != 0:
            raise _lib.DclError("dcl_step_begin failed with status %d: %s" % (rc0, lib.dcl_last_error().decode("utf-8", "replace")))
        after_begin()
        step.begun = 1
    res.loss = loss.reshape(())
    res.finish = None

    def finish():
        _finish_step(crit, res, sb, step, lib, st, loss, ev, want_grad, world, rank, dev, (p_pix, p_dF, p_rowof), _t)
        return res

    if defer and after_begin is not None:
        res.finish = finish
        return res
    return finish()


def _finish_step(crit, res, sb, step, lib, st, loss, ev, want_grad, world, rank, dev, ptrs, _t):
    """Second half of _run_step: dcl_step_fwd and what it leaves behind."""
    p_pix, p_dF, p_rowof = ptrs
    try:
        rc = lib.dcl_step_fwd(ctypes.byref(step), _stream())
    finally:
        step.begun = 0
    if _t is not None:
        _t.append(time.perf_counter())
        _DEBUG_PY_TIMES.append([(b - a) * 1e3 for a, b in zip(_t[:-1], _t[1:])])      # buffers+dzero, carve, rng+struct, C call
    info = sb["info"]
    if rc == 2:
        print("this shoud be never touched! {} {} {}".format(int(info[0]), int(info[1]), int(info[2])))
        raise Exception
    if rc == 3:
        raise RuntimeError("max_samples // total_classes == 0: no views to sample "
                           "(the reference fails in torch.cat at loss.py:345)")
    if rc == _lib.DCL_ERR_LABEL:
        raise ValueError("labels must lie in 0..255 (ACDC train ids + the ignore label): %s"
                         % lib.dcl_last_error().decode("utf-8", "replace"))
    if rc < 0 or rc > 3:
        raise _lib.DclError("dcl_step_fwd failed with status %d: %s" % (rc, lib.dcl_last_error().decode("utf-8", "replace")))
    res.empty = rc == 1
    res.info = tuple(int(v) for v in info)
    if res.empty:                                          # no class qualifies: zero loss, zero gradient
        _count(2)
        crit.last_plan = None
        loss.zero_()                                       # res.loss is a view of it (it may already have been handed out)
        res.n = res.n_pad = res.n_global = 0
        return res
    torch.set_rng_state(st)
    if ev is not None:
        _PROFILE_HOOK("contrast_fwd", ev[0], ev[1])
        if want_grad:
            _PROFILE_HOOK("contrast_bwd", ev[2], ev[3])
    A, n_view, n, n_pad, n_global, on_device = res.info[:6]
    res.n, res.n_pad, res.n_global = n, n_pad, n_global
    res.p_pix, res.p_dF, res.p_rowof = p_pix, p_dF, p_rowof
    # last_plan / last_layout / last_pix are built on first access (device buffers of this step: valid until the
    # next forward of this module)
    crit.__dict__["_last_step"] = (res.info, sb, res.keep, rank, world)
    crit.__dict__["last_n_global"] = n_global
    # classify, plan, select, gather (2), forward (+ backward) [+ pack / unpack when sharded]; the zero-fill is ours too
    _count(1 + 1 + 1 + 2 + (1 if res.dzero is not None else 0) + _launches(MODE_PIXEL, 0) + (_launches(MODE_PIXEL, 1) if want_grad else 0) + (2 if world > 1 else 0))
    return res


# ----------------------------------------------------------------------------------------------
# autograd glue
# ----------------------------------------------------------------------------------------------
class _PixelContrastFn(torch.autograd.Function):
    """loss(feats) for a fixed set of sampled anchor pixels; gradient only w.r.t. feats.  (Stage-by-stage path.)"""

    @staticmethod
    def forward(ctx, feats, pix, y_dev, n_valid, T, Tb, dzero=None):
        B, C, h, w = feats.shape
        n_pad = pix.shape[0]
        nJ = n_pad // _TILE
        dev = feats.device
        st = _stream()
        # state kept for the backward: tiles | sqnorm | colA | colB | rowloss, one allocation
        keep, (p_tiles, p_sq, p_cA, p_cB, p_rl) = _carve(dev, (n_pad * _DIM * 2, n_pad * 4, n_pad * 16, n_pad * 16,
                                                               n_pad * 4))
        loss2 = torch.empty(2, dtype=torch.float32, device=dev)
        nbytes = _ws_bytes(nJ, nJ)
        ws = _workspace(dev, nbytes)
        _lib.call("dcl_gather_tiles", _p(feats), B, h * w, _p(pix), n_pad, p_tiles, p_sq, None, st)
        with _Timed("contrast_fwd"):
            _lib.call("dcl_contrast_fwd", p_tiles, _p(y_dev), p_sq, nJ, 0, nJ, n_valid, MODE_PIXEL, float(T), float(Tb),
                      _p(ws), nbytes, p_cA, p_cB, p_rl, _p(loss2), st)
        _count(1 + _launches(MODE_PIXEL, 0))
        ctx.save_for_backward(keep, y_dev, pix)
        ctx.meta = (nJ, n_pad, (B, C, h, w), (p_tiles, p_cA, p_cB))
        ctx.dzero = dzero          # dense gradient buffer zero-filled while the host planned the sample
        return loss2[1]

    @staticmethod
    def backward(ctx, grad_out):
        keep, y, pix = ctx.saved_tensors
        nJ, n_pad, (B, C, h, w), (p_tiles, p_cA, p_cB) = ctx.meta
        dev = keep.device
        st = _stream()
        dF = torch.empty((n_pad, _DIM), dtype=torch.float32, device=dev)
        nbytes = _ws_bytes(nJ, nJ)
        ws = _workspace(dev, nbytes)
        with _Timed("contrast_bwd"):
            _lib.call("dcl_contrast_bwd", p_tiles, _p(y), p_cA, p_cB, nJ, 0, nJ, MODE_PIXEL, _p(ws), nbytes, _p(dF), st)
        g = grad_out if (grad_out.dtype == torch.float32 and grad_out.is_contiguous()) else \
            grad_out.to(torch.float32).contiguous()
        dfeats, ctx.dzero = ctx.dzero, None
        zero_fill = 0
        if dfeats is None:
            dfeats = torch.empty((B, C, h, w), dtype=torch.float32, device=dev)
            zero_fill = 1
        _lib.call("dcl_scatter_grad", _p(dF), _p(pix), n_pad, _p(g), _p(dfeats), B, h * w, zero_fill, None, st)
        _count(_launches(MODE_PIXEL, 1) + 1 + zero_fill)
        return dfeats, None, None, None, None, None, None


def _grad_scalar(grad_out):
    return grad_out if (grad_out.dtype == torch.float32 and grad_out.is_contiguous()) else \
        grad_out.to(torch.float32).contiguous()


class _StepFn(torch.autograd.Function):
    """The pixel term through dcl_step_fwd / dcl_step_bwd, single GPU or one rank of a sharded job: sampling,
    gather, the N x N forward AND its backward are issued by one C call (reference utils/loss.py:391-415); the
    autograd backward only scales and scatters the eager row gradients."""

    @staticmethod
    def forward(ctx, feats, labels, predict, crit, shard):
        res = _run_step(crit, feats, labels, predict, shard, ctx.needs_input_grad[0], True)
        ctx.res = res
        # the output must not stay reachable from ctx: output -> grad_fn -> ctx -> res -> output would be a reference
        # cycle, and everything the step holds (the 537 MB gradient buffer included) would live until the cyclic
        # garbage collector runs - the allocator then has to cudaMalloc fresh blocks in the meantime (15-45 ms stalls)
        out, res.loss = res.loss, None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        res = ctx.res
        B, C, h, w = res.shape
        dev = grad_out.device
        if res.empty:
            return torch.zeros((B, C, h, w), dtype=torch.float32, device=dev), None, None, None, None
        if res.p_dF is None:
            raise RuntimeError("backward through a pixel-contrast step that was run without requires_grad")
        dfeats, res.dzero = res.dzero, None
        zero_fill = 0
        if dfeats is None:                                 # a second backward through the same graph
            dfeats = torch.empty((B, C, h, w), dtype=torch.float32, device=dev)
            zero_fill = 1
        _lib.call("dcl_step_bwd", res.p_dF, res.p_pix, res.p_rowof, res.n_pad, _p(_grad_scalar(grad_out)), _p(dfeats), B,
                  h * w, zero_fill, None, 0, _stream())
        _count(1 + zero_fill)
        return dfeats, None, None, None, None


class _DoublyFn(torch.autograd.Function):
    """Both terms' touch points on the shared embedding tensor in one autograd node (the reference
    applies SupConLoss to `fine_feat` [2B,...] and PixelContrastLoss to `fine_feat[:B]`, trainer.py:143-158,
    network/weathernet.py:76-82): forward = global average pool of all 2B images + the pixel step on the first B;
    backward = the dense gradient written once (pooled gradient broadcast), anchor gradients added at 1M elements."""

    @staticmethod
    def forward(ctx, feats2, labels, predict, crit):
        B2, C, h, w = feats2.shape
        B = labels.shape[0]
        pooled = torch.empty((B2, C), dtype=torch.float32, device=feats2.device)

        def pool():
            _lib.call("dcl_gap_fwd", _p(feats2), B2 * C, h * w, _p(pooled), _stream())
            _count(1)

        # Order on the stream: classify + count-table copy of the pixel step, THEN the pool (HBM-bound, 0.3 ms at the
        # cfg3 shapes), then the rest of the step - the host waits for the count table and plans while the pool runs,
        # instead of the GPU idling through that round trip after a pool that was issued first.
        # ... and the rest of the step is DEFERRED (res.finish, called by DoublyContrastiveLoss.forward after it has issued
        # the image-level head on `pooled`): the head's launches and their host time then also fall under the pool.
        if _POOL_AFTER_CLASSIFY:
            res = _run_step(crit, feats2[:B], labels, predict, None, ctx.needs_input_grad[0], False, after_begin=pool,
                            defer=True)
            crit.__dict__["_pending_finish"] = res.finish
            res.finish = None                    # no cycle output -> grad_fn -> ctx -> res -> closure -> output
        else:
            pool()
            res = _run_step(crit, feats2[:B], labels, predict, None, ctx.needs_input_grad[0], False)
        ctx.res = res
        ctx.shape2 = (B2, C, h, w)
        out, res.loss = res.loss, None           # no reference cycle through the output (see _StepFn.forward)
        return pooled, out

    @staticmethod
    def backward(ctx, g_pooled, g_pixel):
        res = ctx.res
        B2, C, h, w = ctx.shape2
        B = res.shape[0]
        dev = g_pooled.device
        dx = torch.empty((B2, C, h, w), dtype=torch.float32, device=dev)
        gp = g_pooled.contiguous().to(torch.float32)
        if res.empty or res.p_dF is None:
            _lib.call("dcl_gap_bwd", _p(gp), B2 * C, h * w, _p(dx), 0, _stream())
            _count(1)
        else:
            _lib.call("dcl_step_bwd", res.p_dF, res.p_pix, res.p_rowof, res.n_pad, _p(_grad_scalar(g_pixel)), _p(dx), B,
                      h * w, 0, _p(gp), B2 * C, _stream())
            _count(2)
        return dx, None, None, None


_SMALL_ROWS = 128          # dcl_contrast_small_max_rows()


class _ContrastRowsFn(torch.autograd.Function):
    """Row-normalised contrast of a dense [n,128] matrix (image-level term).  The image-level term at its real size
    (2B rows <= 128) runs in exact fp32 in one launch, forward and gradient together (csrc/dcl_contrast_small.cu);
    larger row counts and the pixel form go through the tensor-core tiles."""

    @staticmethod
    def forward(ctx, Z, y, mode, T, Tb):
        n = Z.shape[0]
        Zc = Z.contiguous().to(torch.float32)
        if mode == MODE_SUPCON and n <= _SMALL_ROWS:
            y32 = y.to(torch.int32).contiguous()
            loss = torch.empty(1, dtype=torch.float32, device=Z.device)
            dZ = torch.empty((n, _DIM), dtype=torch.float32, device=Z.device) if ctx.needs_input_grad[0] else None
            _lib.call("dcl_contrast_small", _p(Zc), _p(y32), n, mode, float(T), float(Tb), _p(loss), _p(dZ), _stream())
            _count(1)
            ctx.small = True
            ctx.save_for_backward(dZ) if dZ is not None else ctx.save_for_backward()
            return loss.reshape(())
        ctx.small = False
        n_pad = (n + _TILE - 1) // _TILE * _TILE
        tiles, sqnorm = pack_rows(Zc, n_pad)
        y_pad = torch.full((n_pad,), -1, dtype=torch.int32, device=Z.device)
        y_pad[:n] = y.to(torch.int32)
        nJ = n_pad // _TILE
        colA, colB, rowloss, loss_sum = contrast_forward(tiles, y_pad, sqnorm, nJ, 0, nJ, n, mode, T, Tb)
        ctx.save_for_backward(tiles, y_pad, colA, colB)
        ctx.meta = (n, nJ, mode)
        return (loss_sum[0] / n).reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.small:
            (dZ,) = ctx.saved_tensors
            return dZ * grad_out.to(torch.float32), None, None, None, None
        tiles, y_pad, colA, colB = ctx.saved_tensors
        n, nJ, mode = ctx.meta
        dF = contrast_backward(tiles, y_pad, colA, colB, nJ, 0, nJ, mode)
        dZ = torch.empty((n, _DIM), dtype=torch.float32, device=tiles.device)
        g = grad_out.to(torch.float32).contiguous()
        _lib.call("dcl_unpack_rows", _p(dF), n, _p(g), _p(dZ), _stream())
        _count(1)
        return dZ, None, None, None, None


class _SupConHeadFn(torch.autograd.Function):
    """Everything between the pooled rows and the image-level loss at its real size (2B <= 128 rows, 128 channels):
    projection MLP, row-normalised contrast, and their gradients - one launch for the MLP, one for the contrast
    (forward and dZ together), two for the MLP's backward (csrc/dcl_contrast_small.cu).  loss.py:120, :161-204."""

    @staticmethod
    def forward(ctx, pooled, W1, b1, W2, b2, yy, T, Tb):
        n = pooled.shape[0]
        dev = pooled.device
        X = pooled.contiguous()
        W1c, b1c, W2c, b2c = W1.contiguous(), b1.contiguous(), W2.contiguous(), b2.contiguous()
        H = torch.empty((n, _DIM), dtype=torch.float32, device=dev)
        Z = torch.empty((n, _DIM), dtype=torch.float32, device=dev)
        dZ = torch.empty((n, _DIM), dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        st = _stream()
        _lib.call("dcl_supcon_mlp_fwd", _p(X), _p(W1c), _p(b1c), _p(W2c), _p(b2c), n, _p(H), _p(Z), st)
        _lib.call("dcl_contrast_small", _p(Z), _p(yy), n, MODE_SUPCON, float(T), float(Tb), _p(loss), _p(dZ), st)
        _count(2)
        ctx.save_for_backward(X, W1c, W2c, H, dZ)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        X, W1, W2, H, dZ = ctx.saved_tensors
        n = X.shape[0]
        dev = X.device
        g = grad_out if (grad_out.dtype == torch.float32 and grad_out.is_contiguous()) else grad_out.to(torch.float32).contiguous()
        buf = torch.empty((2 * n + 2 * _DIM + 2, _DIM), dtype=torch.float32, device=dev)     # dH | dX | dW1 | dW2 | db1 | db2
        dH, dX, dW1, dW2 = buf[:n], buf[n:2 * n], buf[2 * n:2 * n + _DIM], buf[2 * n + _DIM:2 * n + 2 * _DIM]
        db1, db2 = buf[2 * n + 2 * _DIM], buf[2 * n + 2 * _DIM + 1]
        _lib.call("dcl_supcon_mlp_bwd", _p(X), _p(W1), _p(W2), _p(H), _p(dZ), _p(g), n, _p(dH), _p(dX), _p(dW1), _p(db1),
                  _p(dW2), _p(db2), _stream())
        _count(2)
        return dX, dW1, db1, dW2, db2, None, None, None


class _GapFn(torch.autograd.Function):
    """nn.AdaptiveAvgPool2d((1,1)) + flatten, loss.py:115-116."""

    @staticmethod
    def forward(ctx, x):
        B2, C, h, w = x.shape
        xc = x.contiguous()
        pooled = torch.empty((B2, C), dtype=torch.float32, device=x.device)
        _lib.call("dcl_gap_fwd", _p(xc), B2 * C, h * w, _p(pooled), _stream())
        _count(1)
        ctx.shape = (B2, C, h, w)
        return pooled

    @staticmethod
    def backward(ctx, g):
        B2, C, h, w = ctx.shape
        dx = torch.empty((B2, C, h, w), dtype=torch.float32, device=g.device)
        gc = g.contiguous().to(torch.float32)
        _lib.call("dcl_gap_bwd", _p(gc), B2 * C, h * w, _p(dx), 0, _stream())
        _count(1)
        return dx


def contrast_rows(Z, y, mode=MODE_PIXEL, temperature=0.07, base_temperature=0.07):
    """Differentiable N x N contrast of rows Z [n,128] with integer labels y [n] (the reference's
    `_contrastive`, loss.py:339-389, on an already gathered anchor matrix)."""
    _require_cuda(Z, "Z")
    if Z.dim() != 2 or Z.shape[1] != _DIM:
        raise ValueError("Z must be [n, 128]")
    return _ContrastRowsFn.apply(Z, y, mode, temperature, base_temperature)


# ----------------------------------------------------------------------------------------------
# the two modules
# ----------------------------------------------------------------------------------------------
class PixelContrastLoss(nn.Module):
    """Pixel-level supervised contrastive term.  Reference: utils/loss.py:250-415.

    Same constructor (`device=None`) and the same mutable attributes (loss.py:255-262); call as
    `crit(feats, labels=labels, predict=predict)` like trainer.py:135-136.  The sampled pixel
    indices are bit-exact with the reference for the same state of torch's global CPU generator.
    Deviations (documented in DESIGN.md): when no class qualifies the reference crashes
    (loss.py:287-288 then :341); this returns a zero loss that still back-propagates zeros.  Labels must lie in
    0..255 (ACDC train ids + ignore label 255): anything else raises ValueError instead of becoming a class.
    """

    def __init__(self, device=None):
        super().__init__()
        self.device = device
        self.temperature = 0.07
        self.base_temperature = 0.07
        self.ignore_label = 255
        self.max_samples = 1024
        self.max_views = 2
        self.loss_weight = 1
        self.contrast_mode = "all"
        self.last_plan: Optional[AnchorPlan] = None       # exposed for tests / diagnostics
        self.last_layout: Optional[RowLayout] = None
        self.last_pix: Optional[torch.Tensor] = None
        self.last_n_global = 0

    # ---- what the last forward sampled (tests / diagnostics), built lazily after a fused step -------------
    def _last(self, key):
        d = self.__dict__
        if d.get("_last_step") is not None:
            (A, n_view, n, n_pad, n_global, on_device, _, _), sb, keep, rank, world = d["_last_step"]
            an, cap = sb["anchors"], sb["cap"]
            torch.cuda.synchronize(keep.device)
            req = sb["req_dev"][: n_pad * 4].cpu().numpy().reshape(n_pad, 4).copy()
            y = sb["y_dev"][rank * n_pad:(rank + 1) * n_pad].cpu().numpy().copy()
            # class-sorted anchor order of this rank's block -> reference row / anchor of every device row
            plan_np = sb["plan_np"]
            ycnt, yoff = plan_np["ycnt"], plan_np["yoff"]
            order = plan_np["yanchor"][int(yoff[rank]): int(yoff[rank]) + int(ycnt[rank])].astype(np.int64)
            ref_row = np.full(n_pad, -1, dtype=np.int64)
            anchor = np.full(n_pad, -1, dtype=np.int64)
            if n:
                a_rep = np.repeat(order, n_view)
                v_rep = np.tile(np.arange(n_view, dtype=np.int64), order.shape[0])
                ref_row[:n] = v_rep * A + a_rep
                anchor[:n] = a_rep
            ranks = np.zeros((A, n_view), dtype=np.int64)
            if n:
                ranks[anchor[:n], v_rep] = req[:n, 3]
            d["_last_plan"] = AnchorPlan(A, n_view, an[0, :A].copy(), an[1, :A].copy(), an[2, :A].copy(),
                                         an[3, :A].copy(), an[4, :A].copy(), ranks)
            d["_last_layout"] = RowLayout(n, n_pad, req, y, ref_row, anchor)
            d["_last_pix"] = keep[: n_pad * 4].view(torch.int32)
            d["_last_step"] = None
        return d.get("_last_" + key)

    def _set_last(self, key, value):
        self.__dict__["_last_step"] = None
        self.__dict__["_last_" + key] = value

    last_plan = property(lambda self: self._last("plan"), lambda self, v: self._set_last("plan", v))
    last_layout = property(lambda self: self._last("layout"), lambda self, v: self._set_last("layout", v))
    last_pix = property(lambda self: self._last("pix"), lambda self, v: self._set_last("pix", v))

    # ---- buffers of the one-call step ----------------------------------------------------------
    def _step_buffers(self, B, h, w, dev, world):
        """Persistent buffers of dcl_step_fwd: pinned staging (count tables, plan descriptors, host-plan rows), the
        host arrays the plan fills, and the device scratch that does not outlive the call (everything the eager
        backward has consumed by the time the call returns).  Re-made when the batch, the embedding size,
        max_samples' capacity class, the device, the stream or the world size changes."""
        cap = max(_TILE, (int(self.max_samples) + _TILE - 1) // _TILE * _TILE) + _TILE
        stream_id = torch._C._cuda_getCurrentRawStream(dev.index)
        key = (B, h, w, cap, dev.index, stream_id, world)
        sb = self.__dict__.get("_sb")
        if sb is not None and sb["key"] == key:
            return sb
        hw = h * w
        n_chunks = (hw + _CHUNK - 1) // _CHUNK
        A_cap = B * world * 256
        plan_bytes = int(_lib.load().dcl_step_plan_bytes(B, world))
        msg = (2 * cap + 1) * 16
        scratch, ptrs = _carve(dev, (
            B * hw * 2, B * n_chunks * _BINS * 4, world * B * _BINS * 4,          # code, chunk_hist, counts_dev
            cap * 16, world * cap * 4, plan_bytes,                                 # req_dev, y_dev, plan_dev
            world * cap * _DIM * 2, world * cap * 4, world * cap * 16, world * cap * 16, world * cap * 4,   # tiles .. rowloss
            16, msg if world > 1 else 0, world * msg if world > 1 else 0))         # loss_sum, xchg_send, xchg_recv
        (p_code, p_chunk, p_cnt, p_req, p_y, p_plan, p_tiles, p_sq, p_cA, p_cB, p_rl, p_ls, p_xs, p_xr) = ptrs
        base = scratch.data_ptr()
        counts = torch.empty(world * B * _BINS, dtype=torch.int32).pin_memory()
        plan_host = torch.empty(plan_bytes, dtype=torch.uint8).pin_memory()
        stage = torch.empty(cap * 4 + world * cap, dtype=torch.int32).pin_memory()
        info = np.zeros(8, dtype=np.int32)
        anchors = np.empty((5, A_cap), dtype=np.int64)
        ranks = np.empty(cap, dtype=np.int64)
        # views of the plan blob: [ycnt | yoff | ycls] [anchors] [yanchor]
        pn = plan_host.numpy()
        o_anchor = (4 * (2 * world + A_cap) + 63) // 64 * 64
        o_yanchor = (o_anchor + 48 * B * 256 + 63) // 64 * 64
        plan_np = dict(ycnt=pn[: 4 * world].view(np.int32), yoff=pn[4 * world: 8 * world].view(np.int32),
                       yanchor=pn[o_yanchor: o_yanchor + 4 * A_cap].view(np.int32))
        step = _lib.Step()
        step.B, step.h, step.w = B, h, w
        step.world, step.cap = world, cap
        step.code, step.chunk_hist, step.counts_dev = p_code, p_chunk, p_cnt
        step.req_dev, step.y_dev, step.plan_dev = p_req, p_y, p_plan
        step.tiles, step.sqnorm, step.colA, step.colB, step.rowloss, step.loss_sum = p_tiles, p_sq, p_cA, p_cB, p_rl, p_ls
        step.xchg_send, step.xchg_recv = (p_xs, p_xr) if world > 1 else (None, None)
        step.counts_host, step.plan_host, step.plan_bytes = counts.data_ptr(), plan_host.data_ptr(), plan_bytes
        step.stage_host = stage.data_ptr()
        step.info = info.ctypes.data
        step.image, step.cls, step.num_hard, step.num_easy, step.keep_hard = (anchors[i].ctypes.data for i in range(5))
        step.ranks = ranks.ctypes.data
        step.begun = 0
        i32 = scratch.view(torch.int32)
        sb = dict(key=key, cap=cap, scratch=scratch, counts=counts, plan_host=plan_host, stage=stage, info=info,
                  anchors=anchors, ranks=ranks, plan_np=plan_np, step=step, ws_bytes=_ws_cap_bytes(cap, world),
                  req_dev=i32[(p_req.value - base) // 4: (p_req.value - base) // 4 + cap * 4],
                  y_dev=i32[(p_y.value - base) // 4: (p_y.value - base) // 4 + world * cap])
        self.__dict__["_sb"] = sb
        return sb

    # ---- sampling front end (stage-by-stage path) ----------------------------------------------
    def _host_buffers(self, B, dev):
        """Persistent pinned staging for the one D2H (count table) and the one H2D (row requests) of the
        stage-by-stage path, plus the host arrays dcl_host_plan_rows fills."""
        cap = max(_TILE, (int(self.max_samples) + _TILE - 1) // _TILE * _TILE) + _TILE
        key = (B, cap, str(dev))
        hb = getattr(self, "_hb", None)
        if hb is None or hb["key"] != key:
            A_cap = B * 256
            hb = dict(key=key, cap=cap,
                      counts=torch.empty(B * _BINS, dtype=torch.int32).pin_memory(),
                      stage=torch.empty(cap * 5, dtype=torch.int32).pin_memory(),       # req [cap*4] | y [cap]
                      info=np.zeros(4, dtype=np.int32),
                      anchors=np.empty((5, A_cap), dtype=np.int64),
                      ranks=np.empty(cap, dtype=np.int64),     # cap >= any max_samples that maps to this key
                      rows=np.empty((2, cap), dtype=np.int64),
                      event=torch.cuda.Event())
            hb["stage_np"] = hb["stage"].numpy()
            hb["counts_np"] = hb["counts"].numpy()
            self._hb = hb
        return hb

    def _sample_fast(self, feats, labels, predict, want_grad):
        """classify -> (GPU: zero-fill of the gradient buffer) || (host: dcl_host_plan_rows) -> select.
        Same results as _sample + layout_rows; used when the C replay of torch's generator is verified."""
        B, C, h, w = feats.shape
        dev = feats.device
        hb = self._host_buffers(B, dev)
        code, chunk, counts = classify(labels, predict, h, w)
        hb["counts"].copy_(counts.view(-1), non_blocking=True)
        hb["event"].record()
        dzero = None
        if want_grad:
            dzero = torch.zeros_like(feats)            # runs on the GPU while the host plans below
            _count(1)
        hb["event"].synchronize()                      # the one unavoidable sync: the count table
        _check_label_range(hb["counts_np"], B, h * w)
        st = torch.get_rng_state()
        sbuf = st.numpy()
        cap, stage, an, rows, info = hb["cap"], hb["stage_np"], hb["anchors"], hb["rows"], hb["info"]
        rc = _lib.load().dcl_host_plan_rows(
            hb["counts_np"].ctypes.data, B, int(self.ignore_label), int(self.max_samples), int(self.max_views),
            sbuf.ctypes.data, sbuf.nbytes, info.ctypes.data, an[0].ctypes.data, an[1].ctypes.data,
            an[2].ctypes.data, an[3].ctypes.data, an[4].ctypes.data, hb["ranks"].ctypes.data,
            stage.ctypes.data, stage[cap * 4:].ctypes.data, rows[0].ctypes.data, rows[1].ctypes.data)
        if rc == 1:
            self.last_plan = None
            return None
        if rc == 2:
            print("this shoud be never touched! {} {} {}".format(int(info[0]), int(info[1]), int(info[2])))
            raise Exception
        if rc != 0:
            raise _lib.DclError("dcl_host_plan_rows failed with status %d: %s"
                                % (rc, _lib.load().dcl_last_error().decode("utf-8", "replace")))
        torch.set_rng_state(st)
        A, n_view, n, n_pad = (int(v) for v in info)
        # views into the persistent host buffers: valid until the next forward of this module
        self.last_plan = AnchorPlan(A, n_view, an[0, :A], an[1, :A], an[2, :A], an[3, :A], an[4, :A],
                                    hb["ranks"][: A * n_view].reshape(A, n_view))
        if n_view <= 0:
            raise RuntimeError("max_samples // total_classes == 0: no views to sample "
                               "(the reference fails in torch.cat at loss.py:345)")
        self.last_layout = RowLayout(n, n_pad, stage[: n_pad * 4].reshape(n_pad, 4),
                                     stage[cap * 4: cap * 4 + n_pad], rows[0, :n_pad], rows[1, :n_pad])
        packed = hb["stage"].to(dev, non_blocking=True)
        req_dev = packed[: n_pad * 4]
        y_dev = packed[cap * 4: cap * 4 + n_pad]
        pix = select_pixels(code, chunk, B, h * w, req_dev, n_pad)
        self.last_pix = pix
        return pix, y_dev, n, dzero

    def _sample(self, feats, labels, predict):
        B, C, h, w = feats.shape
        code, chunk, counts = classify(labels, predict, h, w)
        counts_np = counts.cpu().numpy()                           # the one unavoidable D2H sync
        _check_label_range(counts_np.reshape(-1), B, h * w)
        plan = plan_anchors(counts_np.reshape(B, 256, 2), int(self.ignore_label), int(self.max_samples),
                            int(self.max_views))
        return code, chunk, plan

    def sample(self, feats, labels, predict):
        """Integer front end: returns (pix [n_pad] i32, y [n_pad] i32, n_valid) on the device, or
        None when no class qualifies.  Consumes the global CPU RNG like the reference."""
        B, C, h, w = feats.shape
        code, chunk, plan = self._sample(feats, labels, predict)
        self.last_plan = plan
        if plan is None:
            return None
        if plan.n_view <= 0:
            raise RuntimeError("max_samples // total_classes == 0: no views to sample "
                               "(the reference fails in torch.cat at loss.py:345)")
        lay = layout_rows(plan, np.arange(plan.A), 0)
        self.last_layout = lay
        host = torch.from_numpy(np.concatenate([lay.req.reshape(-1), lay.y])).pin_memory()
        packed = host.to(feats.device, non_blocking=True)
        req_dev = packed[: lay.n_pad * 4]
        y_dev = packed[lay.n_pad * 4:]
        pix = select_pixels(code, chunk, B, h * w, req_dev, lay.n_pad)
        self.last_pix = pix
        return pix, y_dev, lay.n

    def _check_inputs(self, feats, labels, predict):
        _require_cuda(feats, "feats")
        if labels is None or predict is None:
            raise TypeError("%s needs labels and predict (trainer.py:135-136)" % type(self).__name__)
        if feats.dim() != 4 or feats.shape[1] != _DIM:
            raise ValueError("feats must be [B,128,h,w] (SwiftNet decoder width); got %s"
                             % (tuple(feats.shape),))
        B, C, h, w = feats.shape
        assert predict.shape[-1] == feats.shape[-1], "{} {}".format(predict.shape, feats.shape)
        if labels.dim() != 3 or labels.shape[0] != B or predict.dim() != 4 or predict.shape[0] != B or \
                tuple(predict.shape[2:]) != (h, w):
            raise ValueError("labels must be [B,H,W] and predict [B,C,h,w] matching feats")
        dev = feats.device
        _on_device(labels, dev, "labels")
        _on_device(predict, dev, "predict")
        feats_c = feats.contiguous().to(torch.float32)
        labels_c = labels.contiguous().to(torch.int64)
        predict_c = predict.detach().contiguous().to(torch.float32)
        return feats_c, labels_c, predict_c

    def forward(self, feats, labels=None, predict=None):
        feats_c, labels_c, predict_c = self._check_inputs(feats, labels, predict)
        with torch.cuda.device(feats_c.device):
            if _verify_host_rng() and _FUSED_STEP:
                return _StepFn.apply(feats_c, labels_c, predict_c, self, None)
            if _verify_host_rng():
                sampled = self._sample_fast(feats_c, labels_c, predict_c,
                                            feats_c.requires_grad and torch.is_grad_enabled())
                if sampled is None:
                    return feats_c.sum() * 0.0
                pix, y_dev, n_valid, dzero = sampled
                return _PixelContrastFn.apply(feats_c, pix, y_dev, n_valid, self.temperature,
                                              self.base_temperature, dzero)
            sampled = self.sample(feats_c, labels_c, predict_c)
            if sampled is None:
                return feats_c.sum() * 0.0
            pix, y_dev, n_valid = sampled
            return _PixelContrastFn.apply(feats_c, pix, y_dev, n_valid, self.temperature,
                                          self.base_temperature)


def _check_label_range(counts_flat, B, hw):
    """Every pixel lands in one of the 512 (label, hard/easy) bins unless its label is outside 0..255."""
    tot = np.asarray(counts_flat).reshape(B, -1).sum(axis=1)
    if not bool(np.all(tot == hw)):
        b = int(np.nonzero(tot != hw)[0][0])
        raise ValueError("labels must lie in 0..255 (ACDC train ids + the ignore label): image %d has %d pixels "
                         "outside that range" % (b, hw - int(tot[b])))


def _mask_to_labels(mask: torch.Tensor, batch_size: int) -> torch.Tensor:
    """SupConLoss's explicit `mask` [bsz,bsz] (loss.py:158-159) for the masks the kernels can express: a 0/1 mask
    that is an equivalence relation (mask[i,j] = 1 iff i and j belong to the same group; what `eq(labels, labels.T)`
    and `eye` produce).  Returns group ids; anything else is rejected."""
    if mask.dim() != 2 or mask.shape[0] != batch_size or mask.shape[1] != batch_size:
        raise ValueError("`mask` must be [bsz, bsz]")
    m = mask.float()
    ident = (m > 0).to(torch.int32).argmax(dim=1).to(torch.int32)        # first member of each row's group
    rebuilt = (ident[:, None] == ident[None, :]).float()
    if not bool(torch.equal(rebuilt, m)):
        raise NotImplementedError("only 0/1 masks that form an equivalence relation (same-group masks, as "
                                  "eq(labels, labels.T) or eye produce) are supported by the CUDA path")
    return ident


class SupConLoss(nn.Module):
    """Image-level (weather) supervised contrastive / SimCLR term.  Reference: utils/loss.py:84-205.

    Same constructor and sub-modules (`avgpool`, `projection`) so `state_dict()` / `.to()` behave
    identically; call as `crit(features, class_labels=weather, mask=None)` like trainer.py:117-119.
    `features` is the two-crop batch [2B,C,h,w] (first B = view 1, next B = view 2).
    """

    def __init__(self, temperature=0.07, contrast_mode="all", base_temperature=0.07, weight=None,
                 device=None, opts=None):
        super().__init__()
        self.temperature = temperature
        self.base_temperature = base_temperature
        self.device = device
        self.weight = weight
        self.opts = opts
        feat_dim = 128
        dim_in = 2048 if getattr(self.opts, "deeplab", False) else 128        # loss.py:98-101
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))     # kept for state_dict/API parity; pooling runs in dcl_gap_*
        self.projection = nn.Sequential(nn.Linear(dim_in, dim_in), nn.ReLU(inplace=True),
                                        nn.Linear(dim_in, feat_dim)).to(self.device)
        self.contrast_mode = "all"                                           # loss.py:111

    def forward(self, features, class_labels=None, mask=None):
        _require_cuda(features, "features")
        if features.dim() != 4:
            raise ValueError("features must be [2*bsz, C, h, w]")
        with torch.cuda.device(features.device):
            pooled = _GapFn.apply(features.to(torch.float32))                # loss.py:115-116
            return self.forward_pooled(pooled, class_labels, mask)

    def forward_pooled(self, pooled, class_labels=None, mask=None):
        """Everything after the global average pool (loss.py:117-204) on pooled [2*bsz, C]."""
        if pooled.shape[0] % 2:
            # torch.split(features, [bsz, bsz]) with bsz = n // 2 fails on an odd batch (loss.py:118)
            raise ValueError("features must hold two crops per image: the leading dimension must be even, got %d"
                             % pooled.shape[0])
        bsz = pooled.shape[0] // 2
        fused = self._fused_head(pooled, class_labels, mask)
        if fused is not None:
            return fused
        z = torch.stack([pooled[:bsz], pooled[bsz:2 * bsz]], dim=1)          # loss.py:117-119
        z = self.projection(z)                                               # loss.py:120
        labels = class_labels
        if len(z.shape) < 3:
            raise ValueError("`features` needs to be [bsz, n_views, ...],"
                             "at least 3 dimensions are required")
        batch_size = z.shape[0]
        if labels is not None and mask is not None:
            raise ValueError("Cannot define both `labels` and `mask`")
        elif labels is None and mask is None:
            y = torch.arange(batch_size, device=z.device, dtype=torch.int32)     # mask = eye, loss.py:151
        elif labels is not None:
            labels = labels.contiguous().view(-1, 1)
            if labels.shape[0] != batch_size:
                raise ValueError("Num of labels does not match num of features")
            labels = _on_device(labels, z.device, "class_labels")
            # mask = eq(labels, labels.T) on the raw values (loss.py:157): group id = first row with an equal label,
            # so float or negative labels compare exactly as they do in the reference
            y = torch.eq(labels, labels.T).to(torch.int32).argmax(dim=1).to(torch.int32)
        else:
            y = _mask_to_labels(_on_device(mask, z.device, "mask"), batch_size)
        if self.contrast_mode != "all":
            raise ValueError("Unknown mode: {}".format(self.contrast_mode))
        n_views = z.shape[1]
        Z = torch.cat(torch.unbind(z, dim=1), dim=0)                         # loss.py:161
        yy = y.repeat(n_views)
        return _ContrastRowsFn.apply(Z, yy, MODE_SUPCON, self.temperature, self.base_temperature)


    def _fused_head(self, pooled, class_labels, mask):
        """The image-level head in four launches when it has its usual form: 128-channel pooled rows on the GPU, at
        most 128 of them, the reference's projection (Linear 128-128, ReLU, Linear 128-128) in fp32, and labels that
        are absent, int32 / int64, or given as a mask.  stack -> projection -> unbind -> cat (loss.py:117-120, :161) is
        the projection applied to the rows in their order, and labels.repeat(2) is the group id of row i mod bsz.
        Returns None when the torch path has to run (other shapes / dtypes; same results)."""
        n, bsz = pooled.shape[0], pooled.shape[0] // 2
        proj = self.projection
        if (not pooled.is_cuda or pooled.dtype != torch.float32 or pooled.dim() != 2 or pooled.shape[1] != _DIM
                or n > _SMALL_ROWS or not _FUSED_HEAD or len(proj) != 3 or self.contrast_mode != "all"):
            return None
        l1, l2 = proj[0], proj[2]
        if not (isinstance(l1, nn.Linear) and isinstance(l2, nn.Linear) and isinstance(proj[1], nn.ReLU)
                and l1.weight.shape == (_DIM, _DIM) and l2.weight.shape == (_DIM, _DIM) and l1.bias is not None
                and l2.bias is not None and l1.weight.dtype == torch.float32 and l1.weight.device == pooled.device
                and l2.weight.device == pooled.device):
            return None
        if class_labels is not None and mask is not None:
            raise ValueError("Cannot define both `labels` and `mask`")
        dev = pooled.device
        if mask is not None:
            yy = _mask_to_labels(_on_device(mask, dev, "mask"), bsz).to(torch.int32).repeat(2).contiguous()
        else:
            lab = None
            if class_labels is not None:
                lab = _on_device(class_labels.contiguous().view(-1, 1), dev, "class_labels")
                if lab.shape[0] != bsz:
                    raise ValueError("Num of labels does not match num of features")
                if lab.dtype not in (torch.int32, torch.int64):
                    return None                                      # float labels: compared by value on the torch path
            yy = torch.empty(2 * bsz, dtype=torch.int32, device=dev)
            _lib.call("dcl_group_ids", _p(lab) if lab is not None else None, lab.element_size() if lab is not None else 0,
                      bsz, 2, _p(yy), _stream())
            _count(1)
        return _SupConHeadFn.apply(pooled, l1.weight, l1.bias, l2.weight, l2.bias, yy, self.temperature, self.base_temperature)


class DoublyContrastiveLoss(nn.Module):
    """Both contrastive terms of the `supcon_pixelcontrast_*` criteria (trainer.py:143-158) with their touch points
    on the shared embedding tensor fused: one read of `fine_feat` for the global average pool, one write of its
    dense gradient (pooled gradient broadcast + anchor gradients).  Numerically the same as calling the
    two modules separately, which keeps working.

        crit = DoublyContrastiveLoss(pixel_crit, supcon_crit)
        supcon_loss, pixelcontrast_loss = crit(fine_feat, labels=labels, predict=left_seg_beforeup,
                                               class_labels=weather)          # fine_feat [2B,128,h,w]
        total = (supcon_loss + pixelcontrast_loss) / batch_size + 1.2 * seg_loss   # trainer.py:158

    The pixel term sees fine_feat[:B] (network/weathernet.py:77: `fine_feat0`), B = labels.shape[0].
    """

    def __init__(self, pixel: Optional[PixelContrastLoss] = None, supcon: Optional[SupConLoss] = None, device=None,
                 opts=None):
        super().__init__()
        self.pixel = pixel if pixel is not None else PixelContrastLoss(device=device)
        self.supcon = supcon if supcon is not None else SupConLoss(device=device, opts=opts)

    def forward(self, fine_feat, labels=None, predict=None, class_labels=None, mask=None):
        _require_cuda(fine_feat, "fine_feat")
        if labels is None or predict is None:
            raise TypeError("DoublyContrastiveLoss needs labels and predict")
        if fine_feat.dim() != 4 or fine_feat.shape[1] != _DIM:
            raise ValueError("fine_feat must be [2B,128,h,w]; got %s" % (tuple(fine_feat.shape),))
        B = labels.shape[0]
        if fine_feat.shape[0] != 2 * B:
            raise ValueError("fine_feat must hold two crops per labelled image: got %d images for %d label maps"
                             % (fine_feat.shape[0], B))
        x = fine_feat.contiguous().to(torch.float32)
        _, labels_c, predict_c = self.pixel._check_inputs(x[:B], labels, predict)
        if not (_verify_host_rng() and _FUSED_STEP):
            return (self.supcon(fine_feat, class_labels=class_labels, mask=mask),
                    self.pixel(fine_feat[:B], labels=labels, predict=predict))
        with torch.cuda.device(x.device):
            pooled, pixel_loss = _DoublyFn.apply(x, labels_c, predict_c, self.pixel)
            try:
                supcon_loss = self.supcon.forward_pooled(pooled, class_labels, mask)
            finally:
                finish = self.pixel.__dict__.pop("_pending_finish", None)
                if finish is not None:
                    finish()                     # the rest of the pixel step (its value lands in pixel_loss's storage)
        return supcon_loss, pixel_loss


# ----------------------------------------------------------------------------------------------
# multi-GPU: anchor rows sharded over ranks, contrast set all-gathered
# ----------------------------------------------------------------------------------------------
@dataclass
class ShardPlan:
    plan: AnchorPlan
    layout: RowLayout          # this rank's rows (padded to n_pad, identical on every rank)
    n_pad: int                 # rows per rank block (multiple of 128)
    n_global: int              # valid rows over all ranks == the reference's N on the full batch
    rows_per_rank: np.ndarray  # [world] valid rows per rank


def shard_plan(counts_all: np.ndarray, rank: int, world: int, images_per_rank: int, ignore_label: int,
               max_samples: int, max_views: int,
               randperm: Optional[Callable[[int, int], np.ndarray]] = None) -> Optional[ShardPlan]:
    """Host logic of the sharded sampler.  `counts_all` [world*images_per_rank,256,2] is the
    all-gathered histogram; every rank replays the SAME host RNG stream over the global batch
    (image order = rank-major) and keeps the anchors of its own images, so the union over ranks is
    exactly the single-process reference sample of the concatenated batch."""
    plan = plan_anchors(counts_all, ignore_label, max_samples, max_views, randperm)
    if plan is None:
        return None
    owner = plan.image // images_per_rank
    rows = np.array([int((owner == r).sum()) * plan.n_view for r in range(world)], dtype=np.int64)
    n_pad = max(_TILE, int((rows.max() + _TILE - 1) // _TILE * _TILE))
    mine = np.nonzero(owner == rank)[0]
    lay = layout_rows(plan, mine, rank * images_per_rank, n_pad=n_pad)
    return ShardPlan(plan, lay, n_pad, int(rows.sum()), rows)


def shard_plan_c(counts_all: np.ndarray, rank: int, world: int, images_per_rank: int, ignore_label: int,
                 max_samples: int, max_views: int, stage_req: Optional[np.ndarray] = None,
                 stage_y: Optional[np.ndarray] = None):
    """C form of shard_plan (dcl_host_plan_rows_sharded): same plan, same local layout, plus the labels of every
    rank's row block (so they need no exchange); non-local permutations only advance the generator.
    Returns None (no class qualifies) or (ShardPlan, y_all [world*n_pad] i32).  `stage_req` / `stage_y` may be
    caller-owned (pinned) output arrays of sufficient size."""
    B = world * images_per_rank
    cap = max(_TILE, (int(max_samples) + _TILE - 1) // _TILE * _TILE) + _TILE
    counts = np.ascontiguousarray(counts_all, dtype=np.int32).reshape(-1)
    info = np.zeros(6, dtype=np.int32)
    an = np.empty((5, B * 256), dtype=np.int64)
    ranks = np.empty(cap, dtype=np.int64)
    compact = stage_req is not None and stage_y is None     # labels right behind the requests (one H2D copy)
    req = stage_req if stage_req is not None else np.empty(cap * 4, dtype=np.int32)
    y_all = None if compact else (stage_y if stage_y is not None else np.empty(world * cap, dtype=np.int32))
    rows = np.empty((2, cap), dtype=np.int64)
    st = torch.get_rng_state()
    sbuf = st.numpy()
    lib = _lib.load()
    rc = lib.dcl_host_plan_rows_sharded(counts.ctypes.data, images_per_rank, world, rank, int(ignore_label),
                                        int(max_samples), int(max_views), sbuf.ctypes.data, sbuf.nbytes,
                                        info.ctypes.data, an[0].ctypes.data, an[1].ctypes.data, an[2].ctypes.data,
                                        an[3].ctypes.data, an[4].ctypes.data, ranks.ctypes.data, req.ctypes.data,
                                        None if compact else y_all.ctypes.data, rows[0].ctypes.data, rows[1].ctypes.data)
    if rc == 1:
        return None
    if rc == 2:
        print("this shoud be never touched! {} {} {}".format(int(info[0]), int(info[1]), int(info[2])))
        raise Exception
    if rc != 0:
        raise _lib.DclError("dcl_host_plan_rows_sharded failed with status %d: %s"
                            % (rc, lib.dcl_last_error().decode("utf-8", "replace")))
    torch.set_rng_state(st)
    A, n_view, n, n_pad, n_global = (int(v) for v in info[:5])
    if compact:
        y_all = req[4 * n_pad: (4 + world) * n_pad]
    plan = AnchorPlan(A, n_view, an[0, :A], an[1, :A], an[2, :A], an[3, :A], an[4, :A],
                      ranks[: A * n_view].reshape(A, n_view))
    lay = RowLayout(n, n_pad, req[: n_pad * 4].reshape(n_pad, 4), y_all[rank * n_pad:(rank + 1) * n_pad],
                    rows[0, :n_pad], rows[1, :n_pad])
    owner = plan.image // images_per_rank
    rpr = np.bincount(owner, minlength=world).astype(np.int64) * n_view
    return ShardPlan(plan, lay, n_pad, n_global, rpr), y_all[: world * n_pad]


class ShardedPixelContrastLoss(PixelContrastLoss):
    """Data-parallel form of PixelContrastLoss for one process per GPU (torch.distributed, NCCL).

    Each rank passes ITS images (same per-rank batch size everywhere).  The result equals the
    reference's loss on the concatenated global batch (rank-major image order) and is identical on
    every rank; backward yields d(global loss)/d(local feats) exactly, including the terms that
    come from other ranks' rows.  All ranks must hold the same torch CPU RNG state on entry
    (e.g. `torch.manual_seed(step)` everywhere), because each replays the full host RNG stream.
    The whole step is issued by dcl_step_fwd (csrc/dcl_step.cu), which calls NCCL itself on the step's stream:
    all-gather of the count tables, of the F-tiles (the contrast set) and of the 32-byte row constants.
    The reference has no multi-GPU path; this is new design, not a port.
    """

    def __init__(self, device=None, process_group=None):
        super().__init__(device=device)
        self.process_group = process_group

    def forward(self, feats, labels=None, predict=None):
        feats_c, labels_c, predict_c = self._check_inputs(feats, labels, predict)
        if not _verify_host_rng():
            raise _lib.DclError("the C replay of torch's CPU generator does not match this torch build")
        with torch.cuda.device(feats_c.device):
            shard = _shard_comm(self.process_group)
            return _StepFn.apply(feats_c, labels_c, predict_c, self, shard)

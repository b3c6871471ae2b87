"""Drop-in loss modules: same names, constructor arguments, attributes, forward signatures and
error behaviour as the reference's `utils/loss.py` (`PixelContrastLoss` :250-415, `SupConLoss`
:84-205), with all device work done by libdcl_b200.so (hand-written sm_100a CUDA, C ABI in
include/dcl_b200.h).  PyTorch is used for memory, streams, autograd plumbing and (for the
image-level term only) the two tiny projection GEMMs.  No CPU path exists here.
"""
from __future__ import annotations

import ctypes
import os
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
    """-> code [B,hw] u16 (as int16 storage), chunk_prefix [B,n_chunks,512] i32, counts [B,512] i32"""
    B, H, W = labels.shape
    C = predict.shape[1]
    hw = h * w
    n_chunks = (hw + _CHUNK - 1) // _CHUNK
    dev = labels.device
    code = torch.empty((B, hw), dtype=torch.int16, device=dev)
    chunk = torch.empty((B, n_chunks, _BINS), dtype=torch.int32, device=dev)
    counts = torch.empty((B, _BINS), dtype=torch.int32, device=dev)
    _lib.call("dcl_sample_classify", _p(labels), _p(predict), B, H, W, h, w, C, _p(code), _p(chunk),
              _p(counts), _stream())
    _count(2)
    return code, chunk, counts


def select_pixels(code, chunk, B, hw, req_dev, n_rows):
    pix = torch.empty(n_rows, dtype=torch.int32, device=code.device)
    _lib.call("dcl_sample_select", _p(code), _p(chunk), B, hw, _p(req_dev), n_rows, _p(pix), _stream())
    _count(1)
    return pix


def gather_tiles(feats, pix, n_pad):
    B, C, h, w = feats.shape
    tiles = torch.empty(n_pad * _DIM * 2, dtype=torch.uint8, device=feats.device)
    sqnorm = torch.empty(n_pad, dtype=torch.float32, device=feats.device)
    _lib.call("dcl_gather_tiles", _p(feats), B, h * w, _p(pix), n_pad, _p(tiles), _p(sqnorm), _stream())
    _count(1)
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
# lean single-GPU step: one device allocation per direction, cached sizes, raw pointers into it
# ----------------------------------------------------------------------------------------------
_FUSED_STEP = os.environ.get("DCL_FUSED_STEP", "1") != "0"     # 0: the stage-by-stage Python path (same results)
_WS_BYTES = {}
_WS_CACHE = {}
_LAUNCHES_CACHE = {}


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
    buf = torch.empty(o, dtype=torch.uint8, device=dev)
    base = buf.data_ptr()
    return buf, [ctypes.c_void_p(base + x) for x in offs]


# ----------------------------------------------------------------------------------------------
# autograd glue
# ----------------------------------------------------------------------------------------------
class _PixelContrastFn(torch.autograd.Function):
    """loss(feats) for a fixed set of sampled anchor pixels; gradient only w.r.t. feats."""

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
        _lib.call("dcl_gather_tiles", _p(feats), B, h * w, _p(pix), n_pad, p_tiles, p_sq, st)
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
        _lib.call("dcl_scatter_grad", _p(dF), _p(pix), n_pad, _p(g), _p(dfeats), B, h * w, zero_fill, st)
        _count(_launches(MODE_PIXEL, 1) + 1 + zero_fill)
        return dfeats, None, None, None, None, None, None


_WS_CAP_BYTES = {}


def _ws_cap_bytes(cap):
    """Workspace that serves every row count up to `cap` (the fused step learns its row count only after the plan)."""
    v = _WS_CAP_BYTES.get(cap)
    if v is None:
        v = _WS_CAP_BYTES[cap] = max(_ws_bytes(n, n) for n in range(1, cap // _TILE + 1))
    return v


class _PixelStepFn(torch.autograd.Function):
    """The pixel term through dcl_pixel_fwd / dcl_pixel_bwd: sampling, gather and the N x N forward are issued by
    one C call (same stages, same results as _sample_fast + _PixelContrastFn; the interpreter no longer walks from
    one entry point to the next, which was most of a cfg2 step's host time)."""

    @staticmethod
    def forward(ctx, feats, labels, predict, crit):
        B, C, h, w = feats.shape
        dev = feats.device
        hw = h * w
        n_chunks = (hw + _CHUNK - 1) // _CHUNK
        hb = crit._host_buffers(B, dev)
        cap, an, rows, info = hb["cap"], hb["anchors"], hb["rows"], hb["info"]
        # per-step device state, one allocation: code | chunk prefixes | counts | requests+labels | pix | tiles |
        # sqnorm | colA | colB | rowloss
        keep, (p_code, p_chunk, p_cnt, p_stage, p_pix, p_tiles, p_sq, p_cA, p_cB, p_rl) = _carve(
            dev, (B * hw * 2, B * n_chunks * _BINS * 4, B * _BINS * 4, cap * 20, cap * 4, cap * _DIM * 2, cap * 4,
                  cap * 16, cap * 16, cap * 4))
        want_grad = ctx.needs_input_grad[0]
        dzero = torch.empty_like(feats) if want_grad else None     # cleared on the stream while the host plans
        lib = _lib.load()
        strm = _stream()
        H, W, C_cls = labels.shape[1], labels.shape[2], predict.shape[1]
        zp = dzero.data_ptr() if dzero is not None else None
        zb = dzero.numel() * 4 if dzero is not None else 0
        # first half at once (classify, count table D2H, zero-fill): everything below overlaps it
        rc = lib.dcl_pixel_begin(labels.data_ptr(), predict.data_ptr(), B, H, W, h, w, C_cls, p_code, p_chunk, p_cnt,
                                 hb["counts"].data_ptr(), zp, zb, strm)
        if rc != 0:
            raise _lib.DclError("dcl_pixel_begin failed with status %d: %s"
                                % (rc, lib.dcl_last_error().decode("utf-8", "replace")))
        loss2 = torch.empty(2, dtype=torch.float32, device=dev)
        nbytes = _ws_cap_bytes(cap)
        ws = _workspace(dev, nbytes)
        st = torch.get_rng_state()
        sbuf = st.numpy()
        step = hb.get("step")
        if step is None:
            step = hb["step"] = _lib.PixelStep()
            step.counts_host = hb["counts"].data_ptr()
            step.stage_host = hb["stage"].data_ptr()
            step.cap = cap
            step.info = info.ctypes.data
            step.image, step.cls, step.num_hard, step.num_easy, step.keep_hard = (an[i].ctypes.data for i in range(5))
            step.ranks = hb["ranks"].ctypes.data
            step.ref_row, step.anchor = rows[0].ctypes.data, rows[1].ctypes.data
            step.begun = 1
        step.labels, step.predict, step.feats = labels.data_ptr(), predict.data_ptr(), feats.data_ptr()
        step.B, step.H, step.W, step.h, step.w, step.C_cls = B, H, W, h, w, C_cls
        step.ignore_label, step.max_samples, step.max_views = int(crit.ignore_label), int(crit.max_samples), int(crit.max_views)
        step.temperature, step.base_temperature = float(crit.temperature), float(crit.base_temperature)
        step.torch_rng_state, step.state_bytes = sbuf.ctypes.data, sbuf.nbytes
        step.code, step.chunk_hist, step.counts_dev = p_code, p_chunk, p_cnt
        step.stage_dev, step.pix, step.tiles, step.sqnorm = p_stage, p_pix, p_tiles, p_sq
        step.colA, step.colB, step.rowloss, step.loss_sum = p_cA, p_cB, p_rl, loss2.data_ptr()
        step.workspace, step.workspace_bytes = ws.data_ptr(), nbytes
        step.zero_fill, step.zero_fill_bytes = zp, zb
        ev = None
        if _PROFILE_HOOK is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
            ev[1].record()                                # materialise the handles; the library re-records them
            step.ev_begin, step.ev_end = ev[0].cuda_event, ev[1].cuda_event
        else:
            step.ev_begin = step.ev_end = None
        rc = lib.dcl_pixel_fwd(ctypes.byref(step), strm)
        if rc == 2:
            print("this shoud be never touched! {} {} {}".format(int(info[0]), int(info[1]), int(info[2])))
            raise Exception
        if rc == 3:
            raise RuntimeError("max_samples // total_classes == 0: no views to sample "
                               "(the reference fails in torch.cat at loss.py:345)")
        if rc < 0 or rc > 3:
            raise _lib.DclError("dcl_pixel_fwd failed with status %d: %s"
                                % (rc, _lib.load().dcl_last_error().decode("utf-8", "replace")))
        ctx.empty = rc == 1
        ctx.shape = (B, C, h, w)
        if rc == 1:                                        # no class qualifies: zero loss, zero gradient
            _count(2)
            crit.last_plan = None
            return torch.zeros((), dtype=torch.float32, device=dev)
        torch.set_rng_state(st)
        if ev is not None:
            _PROFILE_HOOK("contrast_fwd", ev[0], ev[1])
        n_pad = int(info[3])
        # last_plan / last_layout / last_pix are built on first access (views into the persistent host buffers and
        # this step's device state: valid until the next forward of this module)
        crit.__dict__["_last_step"] = (tuple(int(v) for v in info), hb, keep, p_pix.value - keep.data_ptr())
        _count(2 + 1 + 1 + _launches(MODE_PIXEL, 0))
        ctx.save_for_backward(keep)
        ctx.meta = (n_pad, cap, (p_stage, p_pix, p_tiles, p_cA, p_cB), nbytes)
        ctx.dzero = dzero
        return loss2[1]

    @staticmethod
    def backward(ctx, grad_out):
        B, C, h, w = ctx.shape
        if ctx.empty:
            return torch.zeros((B, C, h, w), dtype=torch.float32, device=grad_out.device), None, None, None
        (keep,) = ctx.saved_tensors
        n_pad, cap, (p_stage, p_pix, p_tiles, p_cA, p_cB), nbytes = ctx.meta
        dev = keep.device
        dF = torch.empty((n_pad, _DIM), dtype=torch.float32, device=dev)
        ws = _workspace(dev, nbytes)
        g = grad_out if (grad_out.dtype == torch.float32 and grad_out.is_contiguous()) else \
            grad_out.to(torch.float32).contiguous()
        dfeats, ctx.dzero = ctx.dzero, None
        zero_fill = 0
        if dfeats is None:
            dfeats = torch.empty((B, C, h, w), dtype=torch.float32, device=dev)
            zero_fill = 1
        ev = None
        eb = ee = None
        if _PROFILE_HOOK is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
            ev[1].record()
            eb, ee = ev[0].cuda_event, ev[1].cuda_event
        _lib.call("dcl_pixel_bwd", p_tiles, ctypes.c_void_p(p_stage.value + cap * 16), p_cA, p_cB, n_pad, _p(ws), nbytes,
                  _p(dF), p_pix, _p(g), _p(dfeats), B, h * w, zero_fill, eb, ee, _stream())
        if ev is not None:
            _PROFILE_HOOK("contrast_bwd", ev[0], ev[1])
        _count(_launches(MODE_PIXEL, 1) + 1)
        return dfeats, None, None, None


class _ContrastRowsFn(torch.autograd.Function):
    """Row-normalised contrast of a dense [n,128] matrix (image-level term)."""

    @staticmethod
    def forward(ctx, Z, y, mode, T, Tb):
        n = Z.shape[0]
        n_pad = (n + _TILE - 1) // _TILE * _TILE
        Zc = Z.contiguous().to(torch.float32)
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
        tiles, y_pad, colA, colB = ctx.saved_tensors
        n, nJ, mode = ctx.meta
        dF = contrast_backward(tiles, y_pad, colA, colB, nJ, 0, nJ, mode)
        dZ = torch.empty((n, _DIM), dtype=torch.float32, device=tiles.device)
        g = grad_out.to(torch.float32).contiguous()
        _lib.call("dcl_unpack_rows", _p(dF), n, _p(g), _p(dZ), _stream())
        _count(1)
        return dZ, None, None, None, None


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
    Deviation (documented in DESIGN.md): when no class qualifies the reference crashes
    (loss.py:287-288 then :341); this returns a zero loss that still back-propagates zeros.
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

    # ---- what the last forward sampled (tests / diagnostics), built lazily after a fused step -------------
    def _last(self, key):
        d = self.__dict__
        if d.get("_last_step") is not None:
            (A, n_view, n, n_pad), hb, keep, o_pix = d["_last_step"]
            an, rows, stage, cap = hb["anchors"], hb["rows"], hb["stage_np"], hb["cap"]
            d["_last_plan"] = AnchorPlan(A, n_view, an[0, :A], an[1, :A], an[2, :A], an[3, :A], an[4, :A],
                                         hb["ranks"][: A * n_view].reshape(A, n_view))
            d["_last_layout"] = RowLayout(n, n_pad, stage[: n_pad * 4].reshape(n_pad, 4),
                                          stage[cap * 4: cap * 4 + n_pad], rows[0, :n_pad], rows[1, :n_pad])
            d["_last_pix"] = keep[o_pix:o_pix + n_pad * 4].view(torch.int32)
            d["_last_step"] = None
        return d.get("_last_" + key)

    def _set_last(self, key, value):
        self.__dict__["_last_step"] = None
        self.__dict__["_last_" + key] = value

    last_plan = property(lambda self: self._last("plan"), lambda self, v: self._set_last("plan", v))
    last_layout = property(lambda self: self._last("layout"), lambda self, v: self._set_last("layout", v))
    last_pix = property(lambda self: self._last("pix"), lambda self, v: self._set_last("pix", v))

    # ---- sampling front end ------------------------------------------------------------------
    def _host_buffers(self, B, dev):
        """Persistent pinned staging for the one D2H (count table) and the one H2D (row requests) of a step, plus
        the host arrays dcl_host_plan_rows fills; re-made only when the batch or max_samples grows."""
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
                      ranks=np.empty(max(int(self.max_samples), 1), dtype=np.int64),
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
        counts_host = counts.cpu().numpy().reshape(B, 256, 2)      # the one unavoidable D2H sync
        plan = plan_anchors(counts_host, int(self.ignore_label), int(self.max_samples),
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

    def forward(self, feats, labels=None, predict=None):
        _require_cuda(feats, "feats")
        if labels is None or predict is None:
            raise TypeError("PixelContrastLoss needs labels and predict (trainer.py:135-136)")
        if feats.dim() != 4 or feats.shape[1] != _DIM:
            raise ValueError("feats must be [B,128,h,w] (SwiftNet decoder width); got %s"
                             % (tuple(feats.shape),))
        B, C, h, w = feats.shape
        assert predict.shape[-1] == feats.shape[-1], "{} {}".format(predict.shape, feats.shape)
        if labels.dim() != 3 or labels.shape[0] != B or predict.shape[0] != B or \
                tuple(predict.shape[2:]) != (h, w):
            raise ValueError("labels must be [B,H,W] and predict [B,C,h,w] matching feats")
        feats_c = feats.contiguous().to(torch.float32)
        labels_c = labels.contiguous().to(torch.int64)
        predict_c = predict.detach().contiguous().to(torch.float32)
        if _verify_host_rng() and _FUSED_STEP:
            return _PixelStepFn.apply(feats_c, labels_c, predict_c, self)
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
        pooled = _GapFn.apply(features.to(torch.float32))                    # loss.py:115-116
        bsz = pooled.shape[0] // 2
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
            y = labels.view(-1).to(device=z.device, dtype=torch.int32)
            if bool((y < 0).any()):
                raise ValueError("class_labels must be non-negative integers")
        else:
            raise NotImplementedError("an explicit `mask` is not supported by the CUDA path "
                                      "(trainer.py always passes mask=None)")
        if self.contrast_mode != "all":
            raise ValueError("Unknown mode: {}".format(self.contrast_mode))
        n_views = z.shape[1]
        Z = torch.cat(torch.unbind(z, dim=1), dim=0)                         # loss.py:161
        yy = y.repeat(n_views)
        return _ContrastRowsFn.apply(Z, yy, MODE_SUPCON, self.temperature, self.base_temperature)


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
    ranks = np.empty(max(int(max_samples), 1), dtype=np.int64)
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


class _ShardedPixelContrastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, pix, y_all, n_global, T, Tb, group, dzero=None):
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        B, C, h, w = feats.shape
        n_pad = pix.shape[0]
        dev = feats.device
        st = _stream()
        nJ, nI, rb0 = world * n_pad // _TILE, n_pad // _TILE, rank * n_pad // _TILE
        N = world * n_pad
        m4 = n_pad * 4
        # the one real exchange step: the contrast set, gathered straight into place (the norms of the other ranks'
        # rows are recomputed from their tiles by the library, so only the local slice of `sqnorm` is filled)
        tiles = torch.empty(N * _DIM * 2, dtype=torch.uint8, device=dev)
        tl = tiles[rank * n_pad * _DIM * 2:(rank + 1) * n_pad * _DIM * 2]
        # sqnorm | colA | colB | rowloss | loss(2) | send (colA_l | colB_l | loss part), one allocation
        keep, (p_sq, p_cA, p_cB, p_rl, p_loss, p_send) = _carve(dev, (N * 4, N * 16, N * 16, N * 4, 8, (2 * m4 + 4) * 4))
        base = keep.data_ptr()
        _lib.call("dcl_gather_tiles", _p(feats), B, h * w, _p(pix), n_pad, _p(tl), ctypes.c_void_p(p_sq.value + rank * n_pad * 4), st)
        dist.all_gather_into_tensor(tiles, tl, group=group)
        nbytes = _ws_bytes(nI, nJ)
        ws = _workspace(dev, nbytes)
        with _Timed("contrast_fwd"):
            _lib.call("dcl_contrast_fwd", _p(tiles), _p(y_all), p_sq, nJ, rb0, nI, n_global, MODE_PIXEL, float(T), float(Tb),
                      _p(ws), nbytes, p_cA, p_cB, p_rl, p_loss, st)
        _count(1 + _launches(MODE_PIXEL, 0))
        # backward needs every row's constants (the dS_ki terms, 32 B per row) and the loss is the sum over ranks:
        # one all-gather of [colA | colB | local loss sum] per rank, unpacked with two strided copies
        kf = keep.view(torch.float32)
        o_send = (p_send.value - base) // 4
        send = kf[o_send:o_send + 2 * m4 + 4]
        _lib.call("dcl_shard_pack", p_cA, p_cB, p_loss, rank, n_pad, p_send, st)
        recv = torch.empty((world, 2 * m4 + 4), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        _lib.call("dcl_shard_unpack", _p(recv), world, n_pad, p_cA, p_cB, int(n_global), _p(loss), st)
        _count(2)
        ctx.save_for_backward(tiles, y_all, keep, pix)
        ctx.meta = (nJ, rb0, nI, n_pad, (B, C, h, w), (p_cA, p_cB))
        ctx.dzero = dzero
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        tiles, y_all, keep, pix = ctx.saved_tensors
        nJ, rb0, nI, n_pad, (B, C, h, w), (p_cA, p_cB) = ctx.meta
        dev = tiles.device
        st = _stream()
        dF = torch.empty((n_pad, _DIM), dtype=torch.float32, device=dev)
        nbytes = _ws_bytes(nI, nJ)
        ws = _workspace(dev, nbytes)
        with _Timed("contrast_bwd"):
            _lib.call("dcl_contrast_bwd", _p(tiles), _p(y_all), p_cA, p_cB, nJ, rb0, nI, MODE_PIXEL, _p(ws), nbytes,
                      _p(dF), st)
        g = grad_out if (grad_out.dtype == torch.float32 and grad_out.is_contiguous()) else \
            grad_out.to(torch.float32).contiguous()
        dfeats, ctx.dzero = ctx.dzero, None
        zero_fill = 0
        if dfeats is None:
            dfeats = torch.empty((B, C, h, w), dtype=torch.float32, device=dev)
            zero_fill = 1
        _lib.call("dcl_scatter_grad", _p(dF), _p(pix), n_pad, _p(g), _p(dfeats), B, h * w, zero_fill, st)
        _count(_launches(MODE_PIXEL, 1) + 1 + zero_fill)
        return dfeats, None, None, None, None, None, None, None


class ShardedPixelContrastLoss(PixelContrastLoss):
    """Data-parallel form of PixelContrastLoss for one process per GPU (torch.distributed, NCCL).

    Each rank passes ITS images (same per-rank batch size everywhere).  The result equals the
    reference's loss on the concatenated global batch (rank-major image order) and is identical on
    every rank; backward yields d(global loss)/d(local feats) exactly, including the terms that
    come from other ranks' rows.  All ranks must hold the same torch CPU RNG state on entry
    (e.g. `torch.manual_seed(step)` everywhere), because each replays the full host RNG stream.
    The reference has no multi-GPU path; this is new design, not a port.
    """

    def __init__(self, device=None, process_group=None):
        super().__init__(device=device)
        self.process_group = process_group
        self.last_n_global = 0

    def forward(self, feats, labels=None, predict=None):
        import torch.distributed as dist
        _require_cuda(feats, "feats")
        if labels is None or predict is None:
            raise TypeError("ShardedPixelContrastLoss needs labels and predict")
        if feats.dim() != 4 or feats.shape[1] != _DIM:
            raise ValueError("feats must be [B,128,h,w]; got %s" % (tuple(feats.shape),))
        group = self.process_group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        B, C, h, w = feats.shape
        feats_c = feats.contiguous().to(torch.float32)
        labels_c = labels.contiguous().to(torch.int64)
        predict_c = predict.detach().contiguous().to(torch.float32)
        code, chunk, counts = classify(labels_c, predict_c, h, w)
        counts_all = torch.empty((world * B, _BINS), dtype=torch.int32, device=feats.device)
        dist.all_gather_into_tensor(counts_all, counts, group=group)
        hc = getattr(self, "_shard_counts", None)
        if hc is None or hc[0].numel() != world * B * _BINS:
            t = torch.empty(world * B * _BINS, dtype=torch.int32).pin_memory()
            hc = self._shard_counts = (t, t.numpy(), torch.cuda.Event())
        hc[0].copy_(counts_all.view(-1), non_blocking=True)
        hc[2].record()
        dzero = None
        if feats_c.requires_grad and torch.is_grad_enabled():
            dzero = torch.zeros_like(feats_c)              # runs on the GPU while the host plans below
            _count(1)
        hc[2].synchronize()
        counts_host = hc[1].reshape(world * B, 256, 2)
        if _verify_host_rng():
            cap = max(_TILE, (int(self.max_samples) + _TILE - 1) // _TILE * _TILE) + _TILE
            key = (world, cap)
            stg = getattr(self, "_shard_stage", None)
            if stg is None or stg[0] != key:
                t = torch.empty(cap * 4 + world * cap, dtype=torch.int32).pin_memory()
                stg = self._shard_stage = (key, t, t.numpy())
            _, stage_t, stage_np = stg
            out = shard_plan_c(counts_host, rank, world, B, int(self.ignore_label), int(self.max_samples),
                               int(self.max_views), stage_np, None)
            sp = None if out is None else out[0]
        else:
            sp = shard_plan(counts_host, rank, world, B, int(self.ignore_label), int(self.max_samples),
                            int(self.max_views))
            stage_t = None
        self.last_plan = None if sp is None else sp.plan
        if sp is None:
            return feats_c.sum() * 0.0
        if sp.plan.n_view <= 0:
            raise RuntimeError("max_samples // total_classes == 0: no views to sample")
        lay = sp.layout
        self.last_layout, self.last_n_global = lay, sp.n_global
        if stage_t is not None:
            # one small H2D copy: the local requests with every rank's labels right behind them
            packed = stage_t[: (4 + world) * lay.n_pad].to(feats.device, non_blocking=True)
            req_dev, y_all = packed[: lay.n_pad * 4], packed[lay.n_pad * 4:]
        else:
            host = torch.from_numpy(np.concatenate([lay.req.reshape(-1), lay.y])).pin_memory()
            packed = host.to(feats.device, non_blocking=True)
            req_dev, y_dev = packed[: lay.n_pad * 4], packed[lay.n_pad * 4:]
            y_all = torch.empty(world * lay.n_pad, dtype=torch.int32, device=feats.device)
            dist.all_gather_into_tensor(y_all, y_dev.contiguous(), group=group)
        pix = select_pixels(code, chunk, B, h * w, req_dev, lay.n_pad)
        self.last_pix = pix
        return _ShardedPixelContrastFn.apply(feats_c, pix, y_all, sp.n_global, self.temperature,
                                             self.base_temperature, group, dzero)

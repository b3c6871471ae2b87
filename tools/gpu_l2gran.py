"""Experiment: does cudaLimitMaxL2FetchGranularity change the cost of the stride-h*w gather / scatter?"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubly_contrastive_semseg_b200 import _lib, loss as L   # noqa: E402
from doubly_contrastive_semseg_b200.loss import _p, _stream   # noqa: E402

rt = ctypes.CDLL("libcudart.so.12")
B, hw, n = 8, 256 * 512, 8192
feats = torch.randn(B, 128, 256, 512, device="cuda")
g = torch.Generator(device="cuda").manual_seed(1)
pix = torch.randperm(B * hw, generator=g, device="cuda")[:n].int().contiguous()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
dF = torch.randn(n, 128, device="cuda")
one = torch.ones((), device="cuda")
dfe = torch.zeros_like(feats)
for gran in (0, 32, 64, 128):
    if gran:
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(gran))
        val = ctypes.c_size_t()
        rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
        print("set granularity", gran, "rc", rc, "now", val.value)
    for name, fn in (("gather", lambda: L.gather_tiles(feats, pix, n)),
                     ("scatter", lambda: _lib.call("dcl_scatter_grad", _p(dF), _p(pix), n, _p(one), _p(dfe), B, hw, 0, None, _stream()))):
        ts = []
        for _ in range(7):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print("  %-8s gran %3d: median %.1f us (min %.1f)" % (name, gran, float(np.median(ts)), min(ts)))

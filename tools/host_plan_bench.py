"""CPU microbenchmark of dcl_host_plan_rows (no GPU): ns per sampled position for a cfg2 / cfg4-like count table."""
import ctypes, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubly_contrastive_semseg_b200 import _lib
lib = ctypes.CDLL(os.environ["DCL_B200_LIB"]) if os.environ.get("DCL_B200_LIB") else _lib.load()
def run(B, hw, K, max_samples, max_views, iters=40):
    rng = np.random.default_rng(0)
    counts = np.zeros((B, 256, 2), dtype=np.int32)
    for b in range(B):
        w = rng.dirichlet(np.ones(K)) * hw
        for c in range(K):
            n = int(w[c]); counts[b, c, 0] = n // 2; counts[b, c, 1] = n - n // 2
    cap = (max_samples + 127) // 128 * 128 + 128
    info = np.zeros(4, dtype=np.int32); an = np.empty((5, B * 256), dtype=np.int64)
    ranks = np.empty(max_samples, dtype=np.int64); stage = np.empty(cap * 5, dtype=np.int32); rows = np.empty((2, cap), dtype=np.int64)
    torch.manual_seed(3)
    ts = []
    for it in range(iters):
        st = torch.get_rng_state(); sbuf = st.numpy()
        t0 = time.perf_counter()
        rc = lib.dcl_host_plan_rows(counts.ctypes.data, B, 255, max_samples, max_views, sbuf.ctypes.data, sbuf.nbytes,
                                    info.ctypes.data, an[0].ctypes.data, an[1].ctypes.data, an[2].ctypes.data, an[3].ctypes.data,
                                    an[4].ctypes.data, ranks.ctypes.data, stage.ctypes.data, stage[cap * 4:].ctypes.data,
                                    rows[0].ctypes.data, rows[1].ctypes.data)
        ts.append(time.perf_counter() - t0)
        assert rc == 0
        torch.set_rng_state(st)
        time.sleep(0.002)                       # let the look-ahead worker refill
    n = int(info[2])
    tm = (ctypes.c_longlong * 8)()
    if hasattr(lib, "dcl_host_plan_timing"):
        lib.dcl_host_plan_timing(tm)
    print("   last plan sections (cumulative us): anchors %.1f | state+attach %.1f | permutations %.1f | write-back+commit %.1f | rows %.1f"
          % tuple(tm[i] / 1e3 for i in range(5)))
    t = np.median(ts[5:])
    print(f"B={B} hw={hw} K={K} rows={n}: plan {t * 1e6:7.1f} us = {t / n * 1e9:5.1f} ns per sampled position (min {min(ts[5:]) * 1e6:.1f} us)")
run(8, 131072, 16, 8192, 64)
run(8, 131072, 16, 65536, 512)

"""ctypes binding of libdcl_b200.so (the C ABI declared in include/dcl_b200.h).

There is no fallback: if the shared library is missing or the device is not sm_100 every call
raises.  Build the library with `python -m doubly_contrastive_semseg_b200.build`.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DCL_B200_LIB") or os.path.join(_HERE, "libdcl_b200.so")   # env override: tools/ experiments only

_c = ctypes
_vp, _i, _f, _sz = _c.c_void_p, _c.c_int, _c.c_float, _c.c_size_t

# name -> (restype, argtypes); mirrors include/dcl_b200.h one to one
SIGNATURES = {
    "dcl_version": (_i, []),
    "dcl_last_error": (_c.c_char_p, []),
    "dcl_check_device": (_i, []),
    "dcl_debug_flags": (_i, [_i]),
    "dcl_debug_trace": (_i, [_vp]),
    "dcl_debug_cta_times": (_i, [_vp]),
    "dcl_contrast_launches": (_i, [_i, _i]),
    "dcl_debug_partition": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "dcl_sample_classify": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "dcl_sample_select": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "dcl_host_sample_ranks": (_i, [_vp, _sz, _i, _i, _vp, _vp, _vp, _vp]),
    "dcl_host_plan_rows": (_i, [_vp, _i, _i, _i, _i, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dcl_host_plan_rows_sharded": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dcl_debug_plan_device": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _sz] + [_vp] * 16),
    "dcl_host_lookahead_stats": (_i, [_vp]),
    "dcl_host_lookahead_wait": (_i, [_vp]),
    "dcl_host_plan_timing": (_i, [_vp]),
    "dcl_gather_tiles": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "dcl_pack_rows": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "dcl_contrast_workspace_bytes": (_sz, [_i, _i]),
    "dcl_contrast_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "dcl_contrast_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "dcl_scatter_grad": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "dcl_zero_fill": (_i, [_vp, _sz, _i, _vp]),
    "dcl_contrast_small_max_rows": (_i, []),
    "dcl_contrast_small": (_i, [_vp, _vp, _i, _i, _f, _f, _vp, _vp, _vp]),
    "dcl_supcon_mlp_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "dcl_supcon_mlp_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dcl_group_ids": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "dcl_dense_grad": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "dcl_unpack_rows": (_i, [_vp, _i, _vp, _vp, _vp]),
    "dcl_gap_fwd": (_i, [_vp, _i, _i, _vp, _vp]),
    "dcl_gap_bwd": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "dcl_shard_pack": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "dcl_shard_unpack": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "dcl_step_plan_bytes": (_sz, [_i, _i]),
    "dcl_step_begin": (_i, [_vp, _vp]),
    "dcl_step_fwd": (_i, [_vp, _vp]),
    "dcl_step_timing": (_i, [_vp]),
    "dcl_step_sim_timing": (_i, [_i]),
    "dcl_step_sim_elapsed": (_i, [_vp, _vp, _vp]),
    "dcl_step_bwd": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "dcl_focal_workspace_bytes": (_sz, [_i, _i, _i]),
    "dcl_focal_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _sz, _vp]),
    "dcl_focal_bwd": (_i, [_vp, _vp, _vp, _vp, _sz, _vp]),
    "dcl_clock_sampler_start": (_i, [_i]),
    "dcl_clock_sampler_stop": (_i, [_vp]),
    "dcl_comm_unique_id": (_i, [_vp]),
    "dcl_comm_init": (_i, [_vp, _i, _i, _vp]),
    "dcl_comm_destroy": (_i, [_vp]),
    "dcl_comm_all_gather": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "dcl_p2p_create": (_i, [_i, _i, _i, _i, _vp, _vp]),
    "dcl_p2p_open": (_i, [_vp, _vp]),
    "dcl_p2p_destroy": (_i, [_vp]),
}


class Step(ctypes.Structure):
    """dcl_step_t of include/dcl_b200.h, field for field."""
    _fields_ = [
        ("labels", _vp), ("predict", _vp), ("feats", _vp),
        ("B", _i), ("H", _i), ("W", _i), ("h", _i), ("w", _i), ("C_cls", _i), ("ignore_label", _i),
        ("max_samples", _i), ("max_views", _i),
        ("temperature", _f), ("base_temperature", _f),
        ("torch_rng_state", _vp), ("state_bytes", _sz),
        ("world", _i), ("rank", _i), ("comm", _vp), ("p2p", _vp),
        ("cap", _i),
        ("code", _vp), ("chunk_hist", _vp), ("counts_dev", _vp),
        ("req_dev", _vp), ("y_dev", _vp), ("pix", _vp), ("rowof", _vp), ("plan_dev", _vp),
        ("tiles", _vp), ("sqnorm", _vp), ("colA", _vp), ("colB", _vp), ("rowloss", _vp), ("loss_sum", _vp),
        ("loss", _vp),
        ("xchg_send", _vp), ("xchg_recv", _vp), ("dF", _vp),
        ("workspace", _vp), ("workspace_bytes", _sz),
        ("counts_host", _vp), ("plan_host", _vp), ("plan_bytes", _sz), ("stage_host", _vp),
        ("info", _vp), ("image", _vp), ("cls", _vp), ("num_hard", _vp), ("num_easy", _vp), ("keep_hard", _vp),
        ("ranks", _vp),
        ("zero_fill", _vp), ("zero_fill_bytes", _sz), ("side_stream", _vp),
        ("device_plan", _i),
        ("ev_fwd_begin", _vp), ("ev_fwd_end", _vp), ("ev_bwd_begin", _vp), ("ev_bwd_end", _vp),
        ("begun", _i),
    ]


def contrast_launches(mode, backward):
    """kernels one dcl_contrast_fwd / dcl_contrast_bwd call launches (bench.py reports the total per step)"""
    return int(load().dcl_contrast_launches(int(mode), int(backward)))


_lib = None


class DclError(RuntimeError):
    pass


DCL_ERR_COMM, DCL_ERR_LABEL = -4, -5      # enum dcl_status, include/dcl_b200.h


def load():
    """Load the shared library once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DclError(
            "libdcl_b200.so not found at %s - build it with "
            "`python -m doubly_contrastive_semseg_b200.build`; there is no CPU or PyTorch fallback"
            % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale: loud by design
        fn.restype = res
        fn.argtypes = args
    if os.environ.get("DCL_DEBUG_FLAGS"):      # tools/ experiments only (see dcl_debug_flags in the header)
        lib.dcl_debug_flags(int(os.environ["DCL_DEBUG_FLAGS"]))
    _lib = lib
    return lib


def call(name, *args):
    """Call an int-returning entry point and raise DclError on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.dcl_last_error().decode("utf-8", "replace")
        raise DclError("%s failed with status %d: %s" % (name, rc, msg))
    return rc


def workspace_bytes(nI, nJ):
    return int(load().dcl_contrast_workspace_bytes(nI, nJ))

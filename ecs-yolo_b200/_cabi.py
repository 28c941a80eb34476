"""ctypes binding of libecsy.so (include/ecsy.h).  No CPU fallback: a missing library or a non-zero
return code raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ECSY_LIB") or os.path.join(_HERE, "lib", "libecsy.so")   # ECSY_LIB: A/B builds of the same ABI

_p, _i, _l, _f, _z = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes): must list every symbol include/ecsy.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "ecsy_last_error": (C.c_char_p, []),
    "ecsy_abi_version": (_i, []),
    "ecsy_sm_count": (_i, []),
    "ecsy_nchw_to_nhwc_f32": (_i, [_p, _p, _l, _i, _i, _i, _p]),
    "ecsy_nhwc_to_nchw_f32": (_i, [_p, _p, _l, _i, _i, _i, _p]),
    "ecsy_spikes_pack": (_i, [_p, _p, _l, _i, _f, _p]),
    "ecsy_spikes_unpack": (_i, [_p, _p, _l, _i, _p]),
    "ecsy_pack_conv_weight": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "ecsy_lif_ecs_ws_bytes": (_z, [_i, _l, _i, _i, _i, _i]),
    "ecsy_lif_ecs_fwd": (_i, [_p, _l, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _i, _l, _i, _i, _i,
                              _f, _f, _f, _f, _f, _p, _z, _p]),
    "ecsy_spread_dw": (_i, [_p, _p, _p, _p, _p, _l, _i, _i, _i, _i, _p]),
    "ecsy_lif_ecs_fused_supported": (_i, [_i, _i]),
    "ecsy_lif_ecs_fused_fwd": (_i, [_p, _l, _p, _p, _p, _p, _p, _i, _l, _i, _i, _i, _f, _f, _f, _f, _f, _p]),
    "ecsy_lif_ecs_wave_supported": (_i, [_i, _i, _i, _i]),
    "ecsy_lif_ecs_wave_prefers": (_i, [_i, _i, _i, _i]),
    "ecsy_lif_ecs_wave_ws_bytes": (_z, [_i, _l, _i, _i, _i]),
    "ecsy_lif_ecs_wave_fwd": (_i, [_p, _l, _p, _p, _p, _p, _p, _i, _l, _i, _i, _i, _f, _f, _f, _f, _f, _p, _z, _p]),
    "ecsy_lif_ecs_bwd_ws_bytes": (_z, [_i, _l, _i, _i, _i, _i]),
    "ecsy_lif_ecs_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _i, _l, _i, _i, _i,
                              _f, _f, _f, _f, _f, _f, _p, _z, _p]),
    "ecsy_conv_dgrad_ws_bytes": (_z, [_l, _i, _i, _i, _i, _i, _i, _i]),
    "ecsy_conv_dgrad": (_i, [_p, _p, _i, _p, _l, _i, _i, _i, _i, _i, _i, _i, _p, _z, _p]),
    "ecsy_spike_conv_wgrad_ws_bytes": (_z, [_l, _i, _i, _i, _i]),
    "ecsy_spike_conv_wgrad": (_i, [_p, _p, _p, _l, _i, _i, _i, _i, _i, _i, _i, _i, _p, _z, _p]),
    "ecsy_real_conv_wgrad_ws_bytes": (_z, [_l, _i, _i, _i, _i, _i, _i, _i, _i]),
    "ecsy_real_conv_wgrad": (_i, [_p, _p, _l, _p, _l, _i, _i, _i, _i, _i, _i, _i, _i, _p, _z, _p]),
    "ecsy_colsum2": (_i, [_p, _p, _l, _l, _i, _p, _p, _p, _z, _p]),
    "ecsy_lif_silu_ws_bytes": (_z, [_i, _l, _i, _i, _i, _i]),
    "ecsy_lif_silu_fwd": (_i, [_p, _l, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _i, _i, _l, _i, _i, _i,
                               _f, _f, _f, _f, _p, _z, _p]),
    "ecsy_lif_silu_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _i, _l, _i, _i, _i,
                               _f, _f, _f, _f, _p, _z, _p]),
    "ecsy_spike_conv_fwd": (_i, [_p, _p, _i, _p, _p, _p, _p, _l, _l, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ecsy_spike_conv_ts_fwd": (_i, [_p, _p, _i, _p, _p, _p, _p, _l, _l, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ecsy_spike_conv_pair_fwd": (_i, [_p, _p, _i, _p, _p, _p, _p, _l, _l, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ecsy_spike_conv_ts_supported": (_i, [_i, _i]),
    "ecsy_spike_conv_prefers_ts": (_i, [_i, _i, _i]),
    "ecsy_pack_spike_conv_weight": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "ecsy_real_conv_ws_bytes": (_z, [_l, _i, _i, _i, _i, _i, _i, _i, _i, _i]),
    "ecsy_real_conv_fwd": (_i, [_p, _l, _p, _p, _i, _p, _f, _p, _p, _p, _l, _i, _i, _i, _i, _i, _i, _i, _i,
                                _p, _z, _p]),
    "ecsy_tdbn_stats_ws_bytes": (_z, [_l, _i]),
    "ecsy_tdbn_stats": (_i, [_p, _l, _i, _p, _p, _p, _z, _p]),
    "ecsy_tdbn_finish": (_i, [_p, _p, _p, _p, _p, _p, _p, _f, _f, _f, _i, _p, _p, _p, _i, _p]),
    "ecsy_tdbn_bwd_coef": (_i, [_p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _i, _p]),
    "ecsy_affine_add": (_i, [_p, _l, _p, _p, _p, _l, _p, _p, _p, _l, _l, _i, _p]),
    "ecsy_resample": (_i, [_p, _l, _p, _p, _p, _l, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ecsy_maxpool_bwd": (_i, [_p, _l, _p, _p, _l, _i, _i, _i, _i, _i, _i, _p]),
    "ecsy_sumpool_slice": (_i, [_p, _p, _l, _i, _i, _i, _i, _i, _i, _p]),
    "ecsy_tsum": (_i, [_p, _p, _f, _p, _i, _l, _p]),
    "ecsy_detect_decode": (_i, [_p, _p, _p, _p, _f, _i, _i, _i, _i, _i, _l, _l, _p]),
    "ecsy_xty_bf16": (_i, [_p, _p, _p, _p, _l, _i, _i, _f, _p, _p]),
    "ecsy_nms_ws_bytes": (_z, [_l, _i, _i, _i]),
    "ecsy_nms": (_i, [_p, _l, _i, _i, _f, C.c_double, _i, _i, _p, _i, _i, _p, _p, _p, _z, _p]),
    "ecsy_optim_chunk": (_i, []),
    "ecsy_sgd_ema_step": (_i, [_p, _p, _p, _p, _p, _p, _i, _p, _p, _l, C.POINTER(C.c_float), C.POINTER(C.c_float), _i,
                               _f, _i, _i, _i, _f, _f, _p]),
    "ecsy_event_frames_ws_bytes": (_z, [_l, _i, _i, _i]),
    "ecsy_event_frames": (_i, [_p, _p, _p, _p, _l, _l, _i, _i, _i, _i, _i, _p, _p, _p, _z, _p]),
    "ecsy_spike_conv_bwd_ws_bytes": (_z, [_l, _i, _i, _i, _i, _i, _i, _i]),
    "ecsy_spike_conv_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _l, _i, _i, _i, _i, _i, _i, _i, _p, _z, _p]),
    "ecsy_yolo_loss_ws_bytes": (_z, [_i, _l, _i, _l, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ecsy_yolo_loss": (_i, [C.POINTER(_p), C.POINTER(_p), _p, _l, _p, _i, _l, _i, _i, C.POINTER(C.c_int), C.POINTER(C.c_int),
                            C.POINTER(C.c_float), _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _p, _p, _p, _z, _p]),
    "ecsy_tal_loss_ws_bytes": (_z, [_i, _l, _l, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ecsy_tal_loss": (_i, [C.POINTER(_p), C.POINTER(_p), _p, _l, _i, _l, _i, C.POINTER(C.c_int), C.POINTER(C.c_int),
                           C.POINTER(C.c_float), _f, _f, _f, _f, _f, _i, _f, _f, _p, _p, _z, _p]),
    "ecsy_stem_conv_supported": (_i, [_i, _i, _i, _i]),
    "ecsy_stem_conv_ws_bytes": (_z, [_l, _i, _i]),
    "ecsy_stem_conv": (_i, [_p, _l, _i, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _z, _p]),
    "ecsy_ddetect_decode": (_i, [_p, _p, _p, _p, _f, _i, _i, _i, _i, _l, _l, _p]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python ecs-yolo_b200/build.py` "
                "(__graft_entry__.build()).  This package has no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.ecsy_abi_version() != 2:
            raise RuntimeError("libecsy.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().ecsy_last_error().decode()}")

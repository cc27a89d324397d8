"""ctypes binding of libhmgan.so (include/hmgan.h).

There is no CPU fallback: if the shared library is missing or a call fails, the
product path raises.  Build with ``python __graft_entry__.py build`` or
``make -C gan-heightmaps_b200/csrc``.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhmgan.so")

F32, F16, BF16X3 = 0, 1, 2
ACT = {"linear": 0, "leaky_rectify": 1, "rectify": 2, "sigmoid": 3, "tanh": 4}
UP_NONE, UP_NEAREST2, UP_BILINEAR2 = 0, 1, 2


class HmError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("dtype", "B", "H", "W", "C1", "C2", "up", "kh", "kw", "stride", "pad", "transposed",
                 "Ho", "Wo", "Cout", "oH", "oW", "os", "ou", "ov", "split", "act")] + \
               [("slope", C.c_float), ("accumulate", C.c_int32)]


class TcConvPlan(C.Structure):
    _fields_ = [("opaque", C.c_uint8 * 2048)]


_P, _I, _LL, _F = C.c_void_p, C.c_int, C.c_longlong, C.c_float

_PROTOS = {
    "hm_version": ([], C.c_int),
    "hm_last_error_string": ([], C.c_char_p),
    "hm_device_supported": ([], C.c_int),
    "hm_conv_gather": ([C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P], C.c_int),
    "hm_conv_wgrad": ([C.POINTER(ConvDesc), _P, _P, _P, _P, _P], C.c_int),
    "hm_tc_conv": ([C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P], C.c_int),
    "hm_tc_conv_ws": ([C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P, _LL, _P], C.c_int),
    "hm_tc_conv_pool": ([C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P], C.c_int),
    "hm_tc_wgrad": ([C.POINTER(ConvDesc), _P, _P, _P, _P, _P], C.c_int),
    "hm_split_bf16x3": ([_P, _P, _LL, _I, _I, _I, _P], C.c_int),
    "hm_up2conv_wgrad_phases": ([C.POINTER(ConvDesc), _P, _P, _P, _P], C.c_int),
    "hm_im2col_c1": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "hm_s2d_pad64": ([_P, _P, _I, _I, _I, _I, _P], C.c_int),
    "hm_im2col_thin": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "hm_c1s2_conv": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P], C.c_int),
    "hm_c1s2_bwd": ([_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P], C.c_int),
    "hm_maxpool2_bwd_scaled": ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P], C.c_int),
    "hm_scale_rows": ([_P, _P, _P, _I, _LL, _LL, _P], C.c_int),
    "hm_adv_loss_pair": ([_P, _P, _P, _P, _P, _I, _LL, _I, _I, _I, _I, _F, _P, _P, _P], C.c_int),
    "hm_c1s2_bwd_fold": ([_P, _P, _P, _I, _P], C.c_int),
    "hm_c1s2_col2im": ([_P, _P, _I, _I, _I, _P], C.c_int),
    "hm_c1s2_wgrad": ([_P, _P, _P, _I, _I, _I, _P], C.c_int),
    "hm_pack_conv_weight": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "hm_pack_conv_weight_multi": ([_P, _I, _LL, _I, _P], C.c_int),
    "hm_pack_conv_weight_count": ([_I, _I, _I, _I, _I], C.c_longlong),
    "hm_unpack_conv_wgrad": ([_P, _P, _I, _I, _I, _I, _I, _P], C.c_int),
    "hm_bn_stats": ([_P, _I, _LL, _I, _P, _P], C.c_int),
    "hm_bn_finalize": ([_P, _LL, _I, _P, _P, _P, _P, _F, _F, _I, _P, _P, _P, _P, _P], C.c_int),
    "hm_bn_apply_act": ([_P, _P, _I, _LL, _I, _P, _P, _I, _F, _P], C.c_int),
    "hm_bn_bwd_reduce": ([_P, _P, _P, _I, _LL, _I, _P, _P, _I, _F, _P, _P], C.c_int),
    "hm_bn_bwd_apply": ([_P, _P, _P, _P, _I, _LL, _I, _P, _P, _P, _I, _F, _P, _P, _P, _P], C.c_int),
    "hm_bn_bwd_reduce_a": ([_P, _P, _P, _I, _LL, _I, _P, _P, _P, _P, _I, _F, _P, _P], C.c_int),
    "hm_bn_bwd_apply_a": ([_P, _P, _P, _P, _I, _LL, _I, _P, _P, _P, _P, _I, _F, _P, _P, _P, _P], C.c_int),
    "hm_act_bwd": ([_P, _P, _P, _I, _LL, _I, _F, _I, _P], C.c_int),
    "hm_col_sum": ([_P, _I, _LL, _I, _P, _P], C.c_int),
    "hm_maxpool2_fwd": ([_P, _P, _P, _I, _I, _I, _I, _I, _P], C.c_int),
    "hm_maxpool2_bwd": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P], C.c_int),
    "hm_upsample2_bwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "hm_upsample2_fwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "hm_nchw_to_nhwc": ([_P, _P, _I, _I, _I, _I, _I, _P], C.c_int),
    "hm_nhwc_to_nchw": ([_P, _P, _I, _I, _I, _I, _I, _P], C.c_int),
    "hm_cast": ([_P, _I, _P, _I, _LL, _P], C.c_int),
    "hm_u8_normalize": ([_P, _P, _I, _LL, _I, _P], C.c_int),
    "hm_slice_channels": ([_P, _P, _I, _LL, _I, _I, _I, _I, _P], C.c_int),
    "hm_permute": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "hm_adv_loss": ([_P, _P, _I, _LL, _I, _I, _F, _I, _I, _F, _F, _I, _P, _P], C.c_int),
    "hm_recon_loss": ([_P, _P, _P, _I, _LL, _I, _F, _F, _I, _P, _P], C.c_int),
    "hm_rmsprop": ([_P, _P, _P, _LL, _P, _F, _F, _F, _P], C.c_int),
    "hm_adam": ([_P, _P, _P, _P, _LL, _P, _F, _F, _F, _I, _F, _P], C.c_int),
    "hm_adam_dev": ([_P, _P, _P, _P, _LL, _P, _F, _F, _F, _P, _F, _P], C.c_int),
    "hm_inc_i32": ([_P, _P], C.c_int),
}

_lib = None


def load():
    """Load libhmgan.so and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HmError("libhmgan.so is not built (%s missing); run `python __graft_entry__.py build`. "
                      "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (args, res) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    for name in ("hm_tc_conv_supported", "hm_tc_wgrad_supported", "hm_tc_conv_pool_supported"):
        getattr(lib, name).argtypes = [C.POINTER(ConvDesc)]
        getattr(lib, name).restype = C.c_int
    lib.hm_tc_conv_ws_bytes.argtypes = [C.POINTER(ConvDesc)]
    lib.hm_tc_conv_ws_bytes.restype = C.c_longlong
    _lib = lib
    return lib


def exported_symbols():
    return sorted(list(_PROTOS) + ["hm_tc_conv_supported", "hm_tc_wgrad_supported", "hm_tc_conv_ws_bytes",
                                   "hm_tc_conv_pool_supported"])


def pack_count(mode, cout, cin, kh, kw):
    """Elements written by hm_pack_conv_weight(mode, ...): mirrors hm_pack_conv_weight_count (tests/test_abi_cpu.py
    checks the two against each other)."""
    if mode == 2:
        return cin * cout
    return {8: 36 * cout * cin, 11: 64 * cout, 12: 16 * cout * cin, 14: 64 * cin, 15: 256 * cout, 16: 256 * cout,
            17: 4 * cout * cin, 18: 64 * cin, 19: 64 * cout, 20: 36 * cout * cin, 22: 4 * cout * cin}.get(mode, cout * cin * kh * kw)


def pack_job_table(jobs):
    """bytes of an HmPackJob[] (include/hmgan.h) for a list of hm_pack_conv_weight argument tuples
    (w, wp, mode, cout, cin, kh, kw, u, v, dst_dtype); returns (bytes, max_n)."""
    import struct
    out, max_n = b"", 0
    for (w, wp, mode, cout, cin, kh, kw, u, v, _dt) in jobs:
        n = pack_count(mode, cout, cin, kh, kw)
        max_n = max(max_n, n)
        out += struct.pack("<QQiiiiiiiiq", int(w), int(wp), mode, cout, cin, kh, kw, u, v, 0, n)
    return out, max_n


def call(name, *args):
    """Call an int-returning entry point; raise HmError with the library's message on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise HmError("%s failed (%d): %s" % (name, rc, lib.hm_last_error_string().decode()))


def query(name, *args):
    """Call an entry point that answers a question (returns a plain int, no error convention)."""
    return int(getattr(load(), name)(*args))

"""ctypes binding of ``libcrdr_sm100.so`` (the C ABI in ``include/crdr_b200.h``) and ``libcrdr_rans.so``.

PyTorch is only the allocator / stream provider here: every wrapper passes raw device pointers and the
current stream handle across the C boundary.  There is no CPU fallback: if the library is missing the
import of a product path fails loudly with build instructions.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SM100_SO = os.path.join(_HERE, "libcrdr_sm100.so")
RANS_SO = os.path.join(_HERE, "libcrdr_rans.so")

MAX_TAPS = 25
PREC_F16X3, PREC_F16X1 = 0, 1
ENGINE_TCGEN05, ENGINE_TCGEN05_NOTMA, ENGINE_SIMT = 0, 1, 2
EPI_NONE, EPI_RESIDUAL, EPI_GATE, EPI_HALF_TANH = 0, 1, 2, 3


class Planes(C.Structure):
    _fields_ = [("hi", C.c_void_p), ("lo", C.c_void_p), ("cs", C.c_int32), ("coff", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [
        ("inp", Planes),
        ("n", C.c_int32), ("hin", C.c_int32), ("win", C.c_int32),
        ("seg0_off", C.c_int32), ("seg0_len", C.c_int32), ("seg1_off", C.c_int32), ("seg1_len", C.c_int32),
        ("hb", C.c_int32), ("wb", C.c_int32), ("in_stride", C.c_int32),
        ("ntaps", C.c_int32),
        ("dh", C.c_int8 * MAX_TAPS), ("dw", C.c_int8 * MAX_TAPS),
        ("w_hi", C.c_void_p), ("w_lo", C.c_void_p),
        ("k_pad", C.c_int32), ("cout_pad", C.c_int32), ("cout", C.c_int32),
        ("tile_n", C.c_int32),
        ("k_order", C.c_int32),
        ("hout", C.c_int32), ("wout", C.c_int32), ("out_stride", C.c_int32), ("out_ph", C.c_int32), ("out_pw", C.c_int32),
        ("out", Planes),
        ("out_f32", C.c_void_p), ("out_f32_cs", C.c_int32), ("out_f32_coff", C.c_int32),
        ("bias", C.c_void_p),
        ("relu", C.c_int32),
        ("add_vec", C.c_void_p),
        ("mode", C.c_int32),
        ("res", Planes),
        ("res_f32", C.c_void_p), ("res_f32_cs", C.c_int32), ("res_f32_coff", C.c_int32),
        ("trunk", Planes),
        ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("precision", C.c_int32), ("engine", C.c_int32),
    ]


class BottleneckDesc(C.Structure):
    _fields_ = [
        ("inp", Planes),
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("mid", C.c_int32), ("cout", C.c_int32),
        ("w2", C.c_void_p), ("k2_pad", C.c_int32), ("mid_pad", C.c_int32),
        ("w3", C.c_void_p), ("k3_pad", C.c_int32), ("cout_pad", C.c_int32),
        ("bias2", C.c_void_p), ("add2", C.c_void_p), ("bias3", C.c_void_p), ("add3", C.c_void_p),
        ("res", Planes), ("out", Planes),
        ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("precision", C.c_int32),
    ]


class GaussDesc(C.Structure):
    _fields_ = [
        ("y", C.c_void_p), ("y_cs", C.c_int32), ("y_coff", C.c_int32),
        ("mu", C.c_void_p), ("sigma", C.c_void_p),
        ("ms_cs", C.c_int32), ("mu_coff", C.c_int32), ("sigma_coff", C.c_int32),
        ("n", C.c_int32), ("hw", C.c_int32), ("c", C.c_int32),
        ("scale_bound", C.c_float),
        ("scale_table", C.c_void_p), ("ntable", C.c_int32),
        ("yq_planes", Planes),
        ("yq_f32", C.c_void_p), ("yq_f32_cs", C.c_int32), ("yq_f32_coff", C.c_int32),
        ("symbols", C.c_void_p), ("indexes", C.c_void_p), ("likelihood", C.c_void_p),
        ("c_total", C.c_int32), ("nchw_coff", C.c_int32),
        ("symbols16", C.c_void_p), ("indexes8", C.c_void_p),
        ("noise", C.c_void_p), ("likelihood_noisy", C.c_void_p),
    ]


class EbDesc(C.Structure):
    _fields_ = [
        ("z", C.c_void_p), ("z_cs", C.c_int32),
        ("n", C.c_int32), ("hw", C.c_int32), ("c", C.c_int32),
        ("params", C.c_void_p), ("medians", C.c_void_p),
        ("zhat_planes", Planes),
        ("symbols", C.c_void_p), ("zhat_nchw", C.c_void_p), ("likelihood", C.c_void_p),
        ("noise", C.c_void_p), ("likelihood_noisy", C.c_void_p),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("s", Planes), ("ca", C.c_int32),
        ("b", Planes), ("cb", C.c_int32),
        ("n", C.c_int32), ("hs", C.c_int32), ("ws", C.c_int32), ("hb", C.c_int32), ("wb", C.c_int32), ("stride", C.c_int32),
        ("ntaps", C.c_int32),
        ("dh", C.c_int8 * MAX_TAPS), ("dw", C.c_int8 * MAX_TAPS),
        ("out", C.c_void_p),
        ("sa", C.c_int64), ("sb", C.c_int64), ("st", C.c_int64),
        ("scale", C.c_float), ("accumulate", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class PackJob(C.Structure):
    _fields_ = [("master", C.c_void_p), ("map", C.c_void_p), ("count", C.c_int64), ("hi", C.c_void_p), ("lo", C.c_void_p)]


class EpiBwdDesc(C.Structure):
    _fields_ = [
        ("g", Planes), ("out", Planes),
        ("m", C.c_int64), ("c", C.c_int32), ("relu", C.c_int32),
        ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("f32_out", C.c_void_p), ("f32_res", C.c_void_p), ("f32_cs", C.c_int32), ("f32_coff", C.c_int32),
        ("dv", Planes), ("dres", Planes),
        ("partial", C.c_void_p), ("blocks", C.c_int32),
        ("add_vec", C.c_void_p),
        ("leaky_slope", C.c_float),
    ]


class GateDesc(C.Structure):
    _fields_ = [
        ("x", Planes), ("t", Planes), ("a", Planes),
        ("m", C.c_int64), ("c", C.c_int32),
        ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("out", Planes),
        ("out_f32", C.c_void_p), ("out_f32_cs", C.c_int32), ("out_f32_coff", C.c_int32),
        ("g", Planes), ("dx", Planes), ("dt", Planes), ("da", Planes),
        ("partial", C.c_void_p), ("blocks", C.c_int32),
    ]


class GaussBwdDesc(C.Structure):
    _fields_ = [
        ("y", C.c_void_p), ("y_cs", C.c_int32), ("y_coff", C.c_int32),
        ("noise", C.c_void_p), ("ms", C.c_void_p),
        ("ms_cs", C.c_int32), ("mu_coff", C.c_int32), ("sigma_coff", C.c_int32),
        ("n", C.c_int32), ("hw", C.c_int32), ("c", C.c_int32), ("c_total", C.c_int32), ("nchw_coff", C.c_int32),
        ("scale_bound", C.c_float), ("lik_bound", C.c_float), ("coef", C.c_float),
        ("gpre", Planes), ("dy", Planes), ("dmu", Planes), ("dsigma", Planes),
        ("coef_scale", C.c_void_p),
    ]


# every symbol include/crdr_b200.h declares (tests check the library exports all of them)
SM100_SYMBOLS = [
    "crdr_abi_version", "crdr_last_error", "crdr_status_reset", "crdr_status_read", "crdr_status_peek_async",
    "crdr_status_clear_bits", "crdr_conv2d", "crdr_bottleneck_bc",
    "crdr_affine_to_planes", "crdr_image_to_planes", "crdr_image_to_patches", "crdr_planes_to_image", "crdr_phases_to_image", "crdr_nhwc_to_nchw",
    "crdr_gauss_quantize", "crdr_gauss_indexes", "crdr_gauss_dequantize", "crdr_eb_quantize",
    "crdr_eb_dequantize", "crdr_bits_from_likelihood", "crdr_max_abs", "crdr_max_abs_batch",
    "crdr_image_u8_to_patches", "crdr_phases_to_image_u8", "crdr_phases_to_image_ex",
    "crdr_conv_dgrad", "crdr_conv_wgrad_workspace", "crdr_conv_wgrad", "crdr_pack_weights", "crdr_pack_weights_multi", "crdr_epilogue_backward",
    "crdr_colsum_finish", "crdr_gate_forward", "crdr_gate_backward", "crdr_gauss_backward", "crdr_mse_backward",
    "crdr_adam_step", "crdr_sum_squares", "crdr_leaky_relu", "crdr_planes_grad_to_phases",
]

_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    """Load libcrdr_sm100.so (built by ``__graft_entry__.build()`` / ``python -m crdr_b200.build``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SM100_SO):
            raise NativeError(
                f"{SM100_SO} is missing: the CUDA extension is the only implementation of the hot path "
                "(no CPU fallback). Build it with `python -m crdr_b200.build`.")
        L = C.CDLL(SM100_SO)
        L.crdr_last_error.restype = C.c_char_p
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.crdr_status_reset.argtypes = [vp]
        L.crdr_status_read.argtypes = [C.POINTER(C.c_uint32), vp]
        L.crdr_status_peek_async.argtypes = [vp, vp]
        L.crdr_status_clear_bits.argtypes = [C.c_uint32, vp]
        L.crdr_conv2d.argtypes = [C.POINTER(ConvDesc), vp]
        L.crdr_bottleneck_bc.argtypes = [C.POINTER(BottleneckDesc), vp]
        L.crdr_affine_to_planes.argtypes = [vp, i32, i32, i64, i32, vp, vp, Planes, vp]
        L.crdr_image_to_planes.argtypes = [vp, i32, i32, i32, i32, i32, Planes, vp]
        L.crdr_image_to_patches.argtypes = [vp, i32, i32, i32, i32, i32, Planes, vp]
        L.crdr_planes_to_image.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp]
        L.crdr_phases_to_image.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp]
        L.crdr_nhwc_to_nchw.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
        for name in ("crdr_gauss_quantize", "crdr_gauss_indexes", "crdr_gauss_dequantize"):
            getattr(L, name).argtypes = [C.POINTER(GaussDesc), vp]
        for name in ("crdr_eb_quantize", "crdr_eb_dequantize"):
            getattr(L, name).argtypes = [C.POINTER(EbDesc), vp]
        L.crdr_bits_from_likelihood.argtypes = [vp, i32, i64, vp, vp]
        L.crdr_max_abs.argtypes = [vp, i64, vp, vp]
        L.crdr_max_abs_batch.argtypes = [vp, i32, i64, vp, vp]
        L.crdr_image_u8_to_patches.argtypes = [vp, i32, i32, i32, i32, i32, Planes, vp]
        L.crdr_phases_to_image_u8.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp]
        L.crdr_phases_to_image_ex.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, i32, vp]
        f32 = C.c_float
        L.crdr_conv_dgrad.argtypes = [C.POINTER(ConvDesc), vp]
        L.crdr_conv_wgrad_workspace.argtypes = [C.POINTER(WgradDesc)]
        L.crdr_conv_wgrad_workspace.restype = C.c_size_t
        L.crdr_conv_wgrad.argtypes = [C.POINTER(WgradDesc), vp]
        L.crdr_pack_weights.argtypes = [vp, vp, i64, vp, vp, vp]
        L.crdr_pack_weights_multi.argtypes = [vp, i32, vp]
        L.crdr_epilogue_backward.argtypes = [C.POINTER(EpiBwdDesc), vp]
        L.crdr_colsum_finish.argtypes = [vp, i32, i32, i32, i32, vp, f32, i32, vp]
        L.crdr_gate_forward.argtypes = [C.POINTER(GateDesc), vp]
        L.crdr_gate_backward.argtypes = [C.POINTER(GateDesc), vp]
        L.crdr_gauss_backward.argtypes = [C.POINTER(GaussBwdDesc), vp]
        L.crdr_mse_backward.argtypes = [vp, i32, vp, i32, i32, i32, i32, i32, f32, vp, i32, vp]
        L.crdr_adam_step.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, vp, f32, vp, vp]
        L.crdr_sum_squares.argtypes = [vp, i64, vp, vp, vp]
        L.crdr_leaky_relu.argtypes = [Planes, i64, i32, f32, vp]
        L.crdr_planes_grad_to_phases.argtypes = [vp, i32, i32, i32, i32, f32, vp, i32, vp]
        L.crdr_debug_conv_epilogue.argtypes = [i32, i32]
        L.crdr_debug_conv_epilogue.restype = None
        assert L.crdr_abi_version() == 1
        _lib = L
    return _lib


LAUNCH_COUNT = [0]  # kernel-launching C-ABI calls issued by this process (bench.py reports it)


def check(rc, counts=True):
    if rc != 0:
        raise NativeError(f"crdr status {rc}: {lib().crdr_last_error().decode()}")
    if counts:
        LAUNCH_COUNT[0] += 1


def stream_handle():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


FLAG_OVERFLOW, FLAG_TIMEOUT, FLAG_SYM_RANGE = 1, 2, 4


def set_conv_epilogue(enabled=-1, swizzle=-1):
    """Test / bring-up aid: choose between the LEAN (TMA-in / TMA-out) and the staged convolution epilogue."""
    lib().crdr_debug_conv_epilogue(int(enabled), int(swizzle))


def status_reset():
    check(lib().crdr_status_reset(stream_handle()), counts=False)


def status_check():
    """Synchronise the current stream and raise if any kernel flagged fp16 overflow / a pipeline timeout.
    The flag is cleared once it has been reported, so one bad input does not poison later calls."""
    flags = C.c_uint32(0)
    rc = lib().crdr_status_read(C.byref(flags), stream_handle())
    if rc != 0:
        msg = lib().crdr_last_error().decode()
        if flags.value:
            lib().crdr_status_reset(stream_handle())
        raise NativeError(f"crdr status {rc}: {msg}")

"""ORACLE: ctypes front-end over oracle/csrc/rans_oracle.c mirroring
`compressai.ans.{RansEncoder,RansDecoder}` list-based semantics
(used at minnen20_charm_context_model.py:201-202,222-224)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE = os.path.abspath(os.path.join(_HERE, "..", ".."))
_SO = os.path.join(_ORACLE, "librans_oracle.so")


def build_lib(force=False):
    src = os.path.join(_ORACLE, "csrc", "rans_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build_lib())
        p = ctypes.c_void_p
        L.oracle_rans_encode.restype = ctypes.c_long
        L.oracle_rans_encode.argtypes = [p, p, ctypes.c_long, p, ctypes.c_long, p, p, p, ctypes.c_long]
        L.oracle_rans_dec_new.restype = p
        L.oracle_rans_dec_new.argtypes = [ctypes.c_char_p, ctypes.c_long]
        L.oracle_rans_dec_free.argtypes = [p]
        L.oracle_rans_dec_stream.argtypes = [p, p, ctypes.c_long, p, ctypes.c_long, p, p, p]
        L.oracle_pmf_to_quantized_cdf.restype = ctypes.c_int
        L.oracle_pmf_to_quantized_cdf.argtypes = [p, ctypes.c_long, ctypes.c_int, p]
        _lib = L
    return _lib


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _cdf_matrix(cdfs):
    if isinstance(cdfs, np.ndarray) and cdfs.ndim == 2:
        return _i32(cdfs)
    width = max(len(c) for c in cdfs)
    m = np.zeros((len(cdfs), width), dtype=np.int32)
    for i, c in enumerate(cdfs):
        m[i, : len(c)] = c
    return m


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
        s, ix = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1)
        m, sz, off = _cdf_matrix(cdfs), _i32(cdfs_sizes), _i32(offsets)
        cap = 4 * (8 * s.size + 16)
        out = np.empty(cap, dtype=np.uint8)
        n = lib().oracle_rans_encode(_ptr(s), _ptr(ix), s.size, _ptr(m), m.shape[1], _ptr(sz), _ptr(off), _ptr(out), cap)
        assert n >= 0, "oracle rans: output buffer too small"
        return out[:n].tobytes()


class RansDecoder:
    def __init__(self):
        self._h = None

    def set_stream(self, stream):
        self._free()
        self._h = lib().oracle_rans_dec_new(bytes(stream), len(stream))

    def decode_stream(self, indexes, cdfs, cdfs_sizes, offsets):
        ix = _i32(indexes).reshape(-1)
        m, sz, off = _cdf_matrix(cdfs), _i32(cdfs_sizes), _i32(offsets)
        out = np.empty(ix.size, dtype=np.int32)
        lib().oracle_rans_dec_stream(self._h, _ptr(ix), ix.size, _ptr(m), m.shape[1], _ptr(sz), _ptr(off), _ptr(out))
        return out.tolist()

    def decode_with_indexes(self, stream, indexes, cdfs, cdfs_sizes, offsets):
        self.set_stream(stream)
        return self.decode_stream(indexes, cdfs, cdfs_sizes, offsets)

    def _free(self):
        if self._h is not None:
            lib().oracle_rans_dec_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass


def pmf_to_quantized_cdf(pmf, precision=16):
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    cdf = np.empty(p.size + 1, dtype=np.uint32)
    rc = lib().oracle_pmf_to_quantized_cdf(_ptr(p), p.size, int(precision), _ptr(cdf))
    if rc != 0:
        raise ValueError(f"pmf_to_quantized_cdf failed (code {rc})")
    return cdf.astype(np.int64).tolist()

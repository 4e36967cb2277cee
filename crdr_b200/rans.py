"""numpy front-end of libcrdr_rans.so (include/crdr_rans.h): the host range coder of the bitstream."""
import ctypes as C

import numpy as np

from . import native as nv


class CdfTables(C.Structure):
    _fields_ = [("cdfs", C.c_void_p), ("cdf_stride", C.c_int32), ("cdf_sizes", C.c_void_p),
                ("offsets", C.c_void_p), ("n_cdf", C.c_int32), ("prepared", C.c_void_p)]


RANS_SYMBOLS = [
    "crdr_rans_tables_prepare", "crdr_rans_tables_free",
    "crdr_pmf_to_quantized_cdf", "crdr_rans_encode_with_indexes", "crdr_rans_encode_batch",
    "crdr_rans_decoder_new", "crdr_rans_decoder_free", "crdr_rans_decoder_set_stream", "crdr_rans_decoder_set_stream_view",
    "crdr_rans_decoder_decode_stream", "crdr_rans_decode_batch",
    "crdr_rans_encode_batch_i16u8", "crdr_rans_decode_batch_u8", "crdr_rans_pool_info",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        import os
        if not os.path.exists(nv.RANS_SO):
            raise nv.NativeError(f"{nv.RANS_SO} is missing; build it with `python -m crdr_b200.build`")
        L = C.CDLL(nv.RANS_SO)
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.crdr_rans_tables_prepare.restype = vp
        L.crdr_rans_tables_prepare.argtypes = [C.POINTER(CdfTables)]
        L.crdr_rans_tables_free.argtypes = [vp]
        L.crdr_pmf_to_quantized_cdf.argtypes = [vp, i64, i32, vp]
        L.crdr_rans_encode_with_indexes.restype = i64
        L.crdr_rans_encode_with_indexes.argtypes = [vp, vp, i64, C.POINTER(CdfTables), vp, i64]
        L.crdr_rans_encode_batch.argtypes = [i32, vp, vp, vp, C.POINTER(CdfTables), vp, vp, vp, i32]
        L.crdr_rans_decoder_new.restype = vp
        L.crdr_rans_decoder_free.argtypes = [vp]
        L.crdr_rans_decoder_set_stream.argtypes = [vp, C.c_char_p, i64]
        L.crdr_rans_decoder_set_stream_view.argtypes = [vp, C.c_char_p, i64]
        L.crdr_rans_decoder_decode_stream.argtypes = [vp, vp, i64, C.POINTER(CdfTables), vp]
        L.crdr_rans_decode_batch.argtypes = [i32, vp, vp, vp, C.POINTER(CdfTables), vp, i32]
        L.crdr_rans_encode_batch_i16u8.argtypes = [i32, vp, vp, vp, C.POINTER(CdfTables), vp, vp, vp, i32]
        L.crdr_rans_decode_batch_u8.argtypes = [i32, vp, vp, vp, C.POINTER(CdfTables), vp, i32]
        L.crdr_rans_pool_info.argtypes = [C.POINTER(i32), C.POINTER(i32)]
        _lib = L
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def pmf_to_quantized_cdf(pmf, precision=16):
    p = np.ascontiguousarray(pmf, dtype=np.float32)
    cdf = np.empty(p.size + 1, dtype=np.uint32)
    rc = lib().crdr_pmf_to_quantized_cdf(p.ctypes.data, p.size, precision, cdf.ctypes.data)
    if rc:
        raise ValueError(f"pmf_to_quantized_cdf: invalid pmf (code {rc})")
    return cdf.astype(np.int32)


class Tables:
    """CDF tables for the coder: private int32 copies of the arrays plus a prepared native handle
    (crdr_rans_tables_prepare) that owns the encoder / decoder acceleration structures.  The copies make the
    object immune to later in-place changes of the source buffers (load_state_dict, update)."""

    def __init__(self, cdfs, cdf_sizes, offsets):
        self.cdfs = np.array(cdfs, dtype=np.int32, order="C", copy=True)
        self.sizes = np.array(cdf_sizes, dtype=np.int32, copy=True).reshape(-1)
        self.offsets = np.array(offsets, dtype=np.int32, copy=True).reshape(-1)
        assert self.cdfs.ndim == 2 and self.cdfs.shape[0] == self.sizes.size == self.offsets.size
        self.c = CdfTables(self.cdfs.ctypes.data, self.cdfs.shape[1], self.sizes.ctypes.data,
                           self.offsets.ctypes.data, self.cdfs.shape[0], None)
        self._h = lib().crdr_rans_tables_prepare(C.byref(self.c))
        if not self._h:
            raise ValueError("rans tables: invalid CDF table geometry")
        self.c.prepared = self._h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().crdr_rans_tables_free(self._h)
                self._h = None
        except Exception:
            pass


def encode(symbols, indexes, tables):
    s, ix = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1)
    assert s.size == ix.size
    cap = 4 * (s.size + 64)
    while True:
        out = np.empty(cap, dtype=np.uint8)
        n = lib().crdr_rans_encode_with_indexes(s.ctypes.data, ix.ctypes.data, s.size, C.byref(tables.c),
                                                out.ctypes.data, cap)
        if n >= 0:
            return out[:n].tobytes()
        if n == -(2 ** 63):
            raise ValueError("rans encode: CDF index out of range")
        cap = -n


def pool_info():
    """(threads, first cpu) of the coder pool of this process."""
    t, c = C.c_int32(0), C.c_int32(0)
    lib().crdr_rans_pool_info(C.byref(t), C.byref(c))
    return t.value, c.value


def _compact(symbols_list, indexes_list):
    return (all(isinstance(s, np.ndarray) and s.dtype == np.int16 for s in symbols_list) and
            all(isinstance(i, np.ndarray) and i.dtype == np.uint8 for i in indexes_list))


def encode_batch(symbols_list, indexes_list, tables, threads=0):
    """Independent streams coded concurrently on the coder pool.  int32 / int32 arrays, or the compact int16 symbols +
    uint8 indexes the CUDA kernels write for the coder (no conversion either way)."""
    cnt = len(symbols_list)
    compact = _compact(symbols_list, indexes_list)
    if compact:
        ss = [np.ascontiguousarray(s).reshape(-1) for s in symbols_list]
        ii = [np.ascontiguousarray(i).reshape(-1) for i in indexes_list]
        fn = lib().crdr_rans_encode_batch_i16u8
    else:
        ss = [_i32(s).reshape(-1) for s in symbols_list]
        ii = [_i32(i).reshape(-1) for i in indexes_list]
        fn = lib().crdr_rans_encode_batch
    caps = [4 * (s.size + 64) for s in ss]
    while True:
        outs = [np.empty(c, dtype=np.uint8) for c in caps]
        PP = C.c_void_p * cnt
        n = (C.c_int64 * cnt)(*[s.size for s in ss])
        cp = (C.c_int64 * cnt)(*caps)
        lens = (C.c_int64 * cnt)()
        rc = fn(cnt, PP(*[s.ctypes.data for s in ss]), PP(*[i.ctypes.data for i in ii]), n,
                C.byref(tables.c), PP(*[o.ctypes.data for o in outs]), cp, lens, threads)
        if rc == 0:
            return [outs[k][: lens[k]].tobytes() for k in range(cnt)]
        if any(l == -(2 ** 63) for l in lens):
            raise ValueError("rans encode: CDF index out of range")
        caps = [max(c, -l) if l < 0 else c for c, l in zip(caps, lens)]


class Decoder:
    def __init__(self, stream=None):
        self._h = lib().crdr_rans_decoder_new()
        if stream is not None:
            self.set_stream(stream)

    def set_stream(self, stream):
        if not isinstance(stream, bytes):
            stream = bytes(stream)
        self._stream = stream   # the decoder reads these bytes in place (immutable, kept alive here)
        if lib().crdr_rans_decoder_set_stream_view(self._h, stream, len(stream)):
            raise ValueError("rans decoder: stream shorter than 8 bytes")

    def decode_stream(self, indexes, tables, out=None):
        ix = _i32(indexes).reshape(-1)
        if out is None:
            out = np.empty(ix.size, dtype=np.int32)
        if lib().crdr_rans_decoder_decode_stream(self._h, ix.ctypes.data, ix.size, C.byref(tables.c), out.ctypes.data):
            raise ValueError("rans decode: CDF index out of range")
        return out

    def __del__(self):
        try:
            if self._h:
                lib().crdr_rans_decoder_free(self._h)
                self._h = None
        except Exception:
            pass


def decode_batch(decoders, indexes_list, tables, threads=0, outs=None):
    """``outs``: optional list of writable contiguous int32 arrays (e.g. views of a pinned buffer) to decode into."""
    cnt = len(decoders)
    u8 = all(isinstance(i, np.ndarray) and i.dtype == np.uint8 for i in indexes_list)
    ii = [np.ascontiguousarray(i).reshape(-1) for i in indexes_list] if u8 else [_i32(i).reshape(-1) for i in indexes_list]
    if outs is None:
        outs = [np.empty(i.size, dtype=np.int32) for i in ii]
    else:
        assert len(outs) == cnt and all(o.dtype == np.int32 and o.flags.c_contiguous and o.size == i.size
                                        for o, i in zip(outs, ii))
    PP = C.c_void_p * cnt
    n = (C.c_int64 * cnt)(*[i.size for i in ii])
    fn = lib().crdr_rans_decode_batch_u8 if u8 else lib().crdr_rans_decode_batch
    rc = fn(cnt, PP(*[d._h for d in decoders]), PP(*[i.ctypes.data for i in ii]), n,
            C.byref(tables.c), PP(*[o.ctypes.data for o in outs]), threads)
    if rc:
        raise ValueError("rans decode: CDF index out of range")
    return outs


class DecodePlan:
    """A decode_batch call with fixed buffers (pinned index / symbol arrays that a caller reuses call after call):
    the pointer arrays are marshalled once, ``run`` is one foreign call."""

    def __init__(self, decoders, indexes_list, outs):
        cnt = len(decoders)
        assert cnt == len(indexes_list) == len(outs) and cnt > 0
        u8 = all(i.dtype == np.uint8 for i in indexes_list)
        assert u8 or all(i.dtype == np.int32 for i in indexes_list)
        assert all(i.flags.c_contiguous and o.flags.c_contiguous and o.dtype == np.int32 and o.size == i.size
                   for i, o in zip(indexes_list, outs))
        self._keep = (decoders, indexes_list, outs)
        PP = C.c_void_p * cnt
        self._args = (cnt, PP(*[d._h for d in decoders]), PP(*[i.ctypes.data for i in indexes_list]),
                      (C.c_int64 * cnt)(*[i.size for i in indexes_list]))
        self._out = PP(*[o.ctypes.data for o in outs])
        self._fn = lib().crdr_rans_decode_batch_u8 if u8 else lib().crdr_rans_decode_batch

    def run(self, tables, threads=0):
        if self._fn(*self._args, C.byref(tables.c), self._out, threads):
            raise ValueError("rans decode: CDF index out of range")

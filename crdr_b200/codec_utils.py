"""Bitstream container of the reference, byte for byte (src/utils/codec_utils.py:81-143).

header  = <H:u16><W:u16><max|y_hat|:u8><int(q*16):u8>            (MultiRateHeaderHandler.encode :82-103)
.bin    = for s in [header, z_string, y_string]: <len(s):u32 little endian><s>     (save_byte_strings :128-133)
"""
import struct


class MultiRateHeaderHandler:
    def __init__(self, use_non_zero_ind=False):
        if use_non_zero_ind:
            raise NotImplementedError("use_non_zero_ind headers are not used by any shipped config")

    def encode(self, img_size, y_hat=None, rate_ind=0.0, max_abs=None):
        h, w = img_size
        if not (isinstance(h, int) and isinstance(w, int)):
            raise AssertionError("img_size must be two ints")
        if max_abs is None:
            max_abs = float(y_hat.abs().max())
        if hasattr(rate_ind, "item"):
            rate_ind = float(rate_ind.item())
        # informational byte (no decoder reads it): the reference's np.array(v, dtype=np.uint8) wraps under numpy 1.x,
        # so wrap instead of failing an encode whose GPU and rANS work is already done
        m, q = int(max_abs) & 0xFF, int(rate_ind * 16)
        if not (0 <= q <= 255 and 0 <= h <= 65535 and 0 <= w <= 65535):
            raise OverflowError(f"header field out of range (H={h}, W={w}, max={m}, q16={q})")
        return struct.pack("<HHBB", h, w, m, q)

    def decode(self, header):
        h, w, m, q = struct.unpack("<HHBB", header[:6])
        return {"img_size": (h, w), "max_sample": m, "rate_ind": q / 16}


class HeaderHandler:
    """Single-rate header of the stage-1 models (codec_utils.py:12-58): <H:u16><W:u16><max|y_hat|:u8>."""

    def __init__(self, use_non_zero_ind=False):
        if use_non_zero_ind:
            raise NotImplementedError("use_non_zero_ind headers are not used by any shipped config")

    def encode(self, img_size, y_hat=None, rate_ind=None, max_abs=None):
        h, w = img_size
        if not (isinstance(h, int) and isinstance(w, int)):
            raise AssertionError("img_size must be two ints")
        if max_abs is None:
            max_abs = float(y_hat.abs().max())
        if not (0 <= h <= 65535 and 0 <= w <= 65535):
            raise OverflowError(f"header field out of range (H={h}, W={w})")
        return struct.pack("<HHB", h, w, int(max_abs) & 0xFF)

    def decode(self, header):
        h, w, m = struct.unpack("<HHB", header[:5])
        return {"img_size": (h, w), "max_sample": m, "rate_ind": None}


def save_byte_strings(save_path, string_list):
    with open(save_path, "wb") as f:
        for s in string_list:
            f.write(struct.pack("<I", len(s)))
            f.write(s)


def load_byte_strings(load_path):
    out = []
    with open(load_path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        (n,) = struct.unpack_from("<I", data, pos)
        out.append(data[pos + 4: pos + 4 + n])
        pos += 4 + n
    return out

"""Test helper: write RIFF/WAVE files in every sample encoding the decode path accepts (plain and EXTENSIBLE headers)."""
import struct

import numpy as np

U8, S16, S24, S32, F32, F64 = range(6)
_SPEC = {U8: (1, 8), S16: (1, 16), S24: (1, 24), S32: (1, 32), F32: (3, 32), F64: (3, 64)}


def encode(x, kind):
    """float array in [-1, 1) [n, ch] -> interleaved little-endian bytes of the given encoding."""
    x = np.asarray(x, np.float64)
    if kind == U8:
        return np.clip(np.round(x * 128 + 128), 0, 255).astype(np.uint8).tobytes()
    if kind == S16:
        return np.clip(np.round(x * 32768), -32768, 32767).astype("<i2").tobytes()
    if kind == S24:
        v = np.clip(np.round(x * 8388608), -8388608, 8388607).astype(np.int32).reshape(-1)
        b = np.stack([v & 255, (v >> 8) & 255, (v >> 16) & 255], axis=1).astype(np.uint8)
        return b.tobytes()
    if kind == S32:
        return np.clip(np.round(x * 2147483648), -2147483648, 2147483647).astype("<i4").tobytes()
    if kind == F32:
        return x.astype("<f4").tobytes()
    return x.astype("<f8").tobytes()


def wav_bytes(x, sr, kind, extensible=False, extra_chunks=False):
    x = np.asarray(x)
    if x.ndim == 1:
        x = x[:, None]
    ch = x.shape[1]
    code, bits = _SPEC[kind]
    data = encode(x, kind)
    align = ch * bits // 8
    if extensible:
        guid = struct.pack("<H", code) + bytes.fromhex("000000001000800000aa00389b71")
        fmt = struct.pack("<HHIIHH", 0xFFFE, ch, sr, sr * align, align, bits) + struct.pack("<HHI", 22, bits, 0) + guid
    else:
        fmt = struct.pack("<HHIIHH", code, ch, sr, sr * align, align, bits)
        if code == 3:
            fmt += struct.pack("<H", 0)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt
    if extra_chunks:
        chunks += b"fact" + struct.pack("<II", 4, x.shape[0])
        chunks += b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\x00"      # odd-sized chunk + pad byte
    chunks += b"data" + struct.pack("<I", len(data)) + data + (b"\x00" if len(data) & 1 else b"")
    return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks

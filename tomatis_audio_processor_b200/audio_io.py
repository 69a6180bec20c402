"""File edge of the Tomatis path: the part of `soundfile` the reference's process() functions use
(src/process_tomatis.py:225,243,357,434; src/process_tomatis_adaptive.py:179,351).

`soundfile` (libsndfile) is used when it is importable, so FLAC in / FLAC PCM_24 out behave exactly
as in the reference.  Where it is not installed (this build image has no libsndfile) a small
built-in RIFF/WAVE codec covers WAV PCM_16/24/32 and IEEE float, with libsndfile's scaling
conventions (read: int / 2^(bits-1); write: the clipping conversions of libsndfile's src/pcm.c that
python-soundfile switches on, see `quantise_pcm24`), and any FLAC
request raises `AudioFormatUnavailable` -- which the standard/xfade front ends turn into the
reference's own "FLAC failed -> write .wav" fallback (src/process_tomatis.py:242-251).
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

try:                                        # pragma: no cover - depends on the host image
    import soundfile as _sf
except Exception:                           # ImportError, or OSError when libsndfile is missing
    _sf = None


class AudioFormatUnavailable(RuntimeError):
    """The requested container/codec needs libsndfile, which is not installed."""


@dataclass
class AudioInfo:
    samplerate: int
    channels: int
    frames: int
    subtype: str
    format: str


def have_soundfile() -> bool:
    return _sf is not None


def _ext(path: str) -> str:
    return os.path.splitext(str(path))[1].lower()


# ------------------------------------------------------------------------------------------------
# built-in WAV codec
_WAVE_FORMAT_PCM, _WAVE_FORMAT_FLOAT, _WAVE_FORMAT_EXTENSIBLE = 0x0001, 0x0003, 0xFFFE


def _wav_header(path):
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] not in (b"RIFF", b"RF64") or head[8:12] != b"WAVE":
            raise ValueError(f"{path}: not a RIFF/WAVE file")
        fmt = None
        while True:
            ck = f.read(8)
            if len(ck) < 8:
                raise ValueError(f"{path}: no data chunk")
            cid, size = ck[:4], struct.unpack("<I", ck[4:])[0]
            if cid == b"fmt ":
                raw = f.read(size + (size & 1))
                tag, ch, sr, _, block, bits = struct.unpack("<HHIIHH", raw[:16])
                if tag == _WAVE_FORMAT_EXTENSIBLE and size >= 26:
                    tag = struct.unpack("<H", raw[24:26])[0]
                fmt = (tag, ch, sr, block, bits)
            elif cid == b"data":
                if fmt is None:
                    raise ValueError(f"{path}: data chunk before fmt chunk")
                offset = f.tell()
                remaining = os.path.getsize(path) - offset
                if size == 0xFFFFFFFF or size > remaining:
                    size = remaining
                return fmt, offset, size
            else:
                f.seek(size + (size & 1), 1)


def _wav_info(path) -> AudioInfo:
    (tag, ch, sr, block, bits), _, size = _wav_header(path)
    sub = {(_WAVE_FORMAT_PCM, 16): "PCM_16", (_WAVE_FORMAT_PCM, 24): "PCM_24", (_WAVE_FORMAT_PCM, 32): "PCM_32",
           (_WAVE_FORMAT_PCM, 8): "PCM_U8", (_WAVE_FORMAT_FLOAT, 32): "FLOAT", (_WAVE_FORMAT_FLOAT, 64): "DOUBLE"}.get((tag, bits))
    if sub is None:
        raise ValueError(f"{path}: unsupported WAV encoding (format tag {tag}, {bits} bits)")
    return AudioInfo(sr, ch, size // block, sub, "WAV")


def _wav_decode(raw, subtype, n, dtype):
    """n interleaved samples from the data chunk's bytes, scaled like libsndfile's float reads."""
    if subtype == "PCM_16":
        return raw[:2 * n].view("<i2").astype(dtype) / dtype(32768.0)
    if subtype == "PCM_24":
        b = raw[:3 * n].reshape(-1, 3).astype(np.int32)
        v = (b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16))
        v = np.where(v & 0x800000, v - 0x1000000, v)
        return v.astype(dtype) / dtype(8388608.0)
    if subtype == "PCM_32":
        return (raw[:4 * n].view("<i4").astype(np.float64) / 2147483648.0).astype(dtype)
    if subtype == "PCM_U8":
        return (raw[:n].astype(dtype) - dtype(128.0)) / dtype(128.0)
    if subtype == "FLOAT":
        return raw[:4 * n].view("<f4").astype(dtype)
    return raw[:8 * n].view("<f8").astype(dtype)


def _wav_read(path, dtype):
    info = _wav_info(path)
    _, offset, size = _wav_header(path)
    n = info.frames * info.channels
    with open(path, "rb") as f:
        f.seek(offset)
        raw = np.frombuffer(f.read(size), dtype=np.uint8)
    return _wav_decode(raw, info.subtype, n, dtype).reshape(info.frames, info.channels), info.samplerate


def _scaled_int32(y) -> np.ndarray:
    """lrint(x * 2^31) clipped to int32 -- the first half of libsndfile's f2le*_clip_array / d2le*_clip_array (src/pcm.c):
    `scaled_value = src[i] * normfact` with normfact = 8.0 * 0x10000000, values >= 0x7FFFFFFF and <= -0x80000000 pinned."""
    v = np.rint(np.asarray(y, dtype=np.float64) * 2147483648.0)            # exact for float32 data (a power-of-two scale)
    return np.clip(v, -2147483648.0, 2147483647.0).astype(np.int64)


def quantise_pcm24(y: np.ndarray, container: str = "FLAC") -> np.ndarray:
    """float -> int32 holding 24-bit samples, as libsndfile 1.2.x writes them with clipping switched on -- python-soundfile
    switches it on for every file it opens (`sf_command(SFC_SET_CLIPPING, SF_TRUE)` in `SoundFile.__init__`), and the
    reference writes through python-soundfile (src/process_tomatis.py:243,357).

    FLAC (src/flac.c, f2flac24_clip_array / d2flac24_clip_array): lrint(x * 2^23), round half to even; scaled values
          >= 0x7FFFFF give 0x7FFFFF, <= -0x800000 give -0x800000.
    WAV  (src/pcm.c, f2let_clip_array / d2let_clip_array): lrint(x * 2^31) clipped to int32, of which the top three bytes are
          stored -- an arithmetic shift, i.e. floor(x * 2^23) for float32 data.
    The two containers therefore differ by up to one step of 2^-23.  Restated from the library's published source; it could
    not be checked against a binary (no libsndfile in this image): treat the PCM edge as "within 1 LSB" (DESIGN.md section 0)."""
    if str(container).upper() == "WAV":
        return (_scaled_int32(y) >> 8).astype(np.int32)
    v = np.rint(np.asarray(y, dtype=np.float64) * 8388608.0)
    return np.clip(v, -8388608, 8388607).astype(np.int32)


def quantise_pcm16(y: np.ndarray, container: str = "WAV") -> np.ndarray:
    """float -> int16 the same way: WAV stores the top two bytes of lrint(x * 2^31) (f2les_clip_array), FLAC lrint(x * 2^15)
    clipped (f2flac16_clip_array)."""
    if str(container).upper() == "WAV":
        return (_scaled_int32(y) >> 16).astype(np.int16)
    return np.clip(np.rint(np.asarray(y, dtype=np.float64) * 32768.0), -32768, 32767).astype(np.int16)


def _wav_write(path, y, sr, subtype):
    y = np.asarray(y)
    if y.ndim == 1:
        y = y[:, None]
    frames, ch = y.shape
    if subtype == "PCM_24":
        v = quantise_pcm24(y, "WAV").reshape(-1)
        data = np.empty((v.size, 3), dtype=np.uint8)
        data[:, 0] = v & 0xFF
        data[:, 1] = (v >> 8) & 0xFF
        data[:, 2] = (v >> 16) & 0xFF
        tag, bits = _WAVE_FORMAT_PCM, 24
    elif subtype == "PCM_16":
        data = quantise_pcm16(y, "WAV").astype("<i2").reshape(-1)
        tag, bits = _WAVE_FORMAT_PCM, 16
    elif subtype == "FLOAT":
        data = y.astype("<f4").reshape(-1)
        tag, bits = _WAVE_FORMAT_FLOAT, 32
    else:
        raise ValueError(f"unsupported WAV subtype {subtype!r}")
    payload = data.tobytes()
    block = ch * bits // 8
    if len(payload) > 0xFFFFFFFF - 64:
        raise ValueError("built-in WAV writer: file would exceed 4 GiB (install soundfile for RF64/FLAC)")
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(payload) + (len(payload) & 1)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, tag, ch, int(sr), int(sr) * block, block, bits))
        f.write(b"data" + struct.pack("<I", len(payload)))
        f.write(payload)
        if len(payload) & 1:
            f.write(b"\0")


# ------------------------------------------------------------------------------------------------
def info(path) -> AudioInfo:
    if _sf is not None:
        i = _sf.info(path)
        return AudioInfo(i.samplerate, i.channels, i.frames, i.subtype, i.format)
    if _ext(path) != ".wav":
        raise AudioFormatUnavailable(f"reading {_ext(path) or 'this'} files needs the soundfile package (libsndfile)")
    return _wav_info(path)


def read(path, dtype="float32"):
    """(x [N, ch] always 2-D, samplerate) -- sf.read(path, dtype=..., always_2d=True)."""
    if _sf is not None:
        return _sf.read(path, dtype=dtype, always_2d=True)
    if _ext(path) != ".wav":
        raise AudioFormatUnavailable(f"reading {_ext(path) or 'this'} files needs the soundfile package (libsndfile)")
    return _wav_read(path, np.dtype(dtype).type)


def read_range(path, start: int, stop: int, dtype="float32"):
    """Sample-frames [start, stop) of a file, [stop - start, ch] -- what one rank of a time-sharded run loads
    (fin.seek(start); fin.read(stop - start, always_2d=True))."""
    start, stop = int(start), int(stop)
    if _sf is not None:
        x, _ = _sf.read(path, start=start, stop=stop, dtype=dtype, always_2d=True)
        return x
    if _ext(path) != ".wav":
        raise AudioFormatUnavailable(f"reading {_ext(path) or 'this'} files needs the soundfile package (libsndfile)")
    i = _wav_info(path)
    (_, ch, _, block, _), offset, size = _wav_header(path)
    start, stop = max(0, min(start, i.frames)), max(0, min(stop, i.frames))
    stop = max(stop, start)
    with open(path, "rb") as f:
        f.seek(offset + start * block)
        raw = np.frombuffer(f.read((stop - start) * block), dtype=np.uint8)
    return _wav_decode(raw, i.subtype, (stop - start) * ch, np.dtype(dtype).type).reshape(stop - start, ch)


def write(path, y, samplerate, subtype="PCM_24", format=None):
    """sf.write(path, y, sr, subtype=..., format=...); the container follows the extension unless given."""
    if _sf is not None:
        return _sf.write(path, y, samplerate, subtype=subtype, format=format)
    fmt = (format or {".wav": "WAV", ".flac": "FLAC", ".ogg": "OGG"}.get(_ext(path), "")).upper()
    if fmt != "WAV":
        raise AudioFormatUnavailable(f"writing {fmt or _ext(path)} needs the soundfile package (libsndfile)")
    _wav_write(path, y, samplerate, subtype)

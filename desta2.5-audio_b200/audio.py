"""File -> 16 kHz mono waveform on the GPU (scope row n3): host mirror of the reference's `AudioSegment.from_file`
(ref: desta/utils/audio.py:245-361 `from_file`, :126-181 `__init__`, :60-119 `select_channels`, :221-243 sample conversion,
:161-166 `librosa.core.resample`; called as ``AudioSegment.from_file(path, target_sr=16000, channel_selector="average")
.samples`` at desta/models/modeling_desta25.py:1553-1557 and desta/trainer/data/simple_dataset.py:145-149).

The host parses the RIFF header and copies the raw PCM bytes; sample conversion, channel average, resampling and the
right-pad / truncate to 30 s run in one kernel (`desta_decode_resample`, csrc/resample.cu).  `load_waveforms` is the
batched entry point that fills the `[N, 480000]` buffer the log-mel kernel reads, so a non-16 kHz corpus never touches a
CPU resampler.  Containers other than RIFF/WAVE PCM / IEEE-float need libsndfile or ffmpeg, which this path does not
replace: they raise the reference's "could not be decoded" exception.
"""
import os
import struct

import numpy as np
import torch

from . import ops
from .processor import N_SAMPLES

PCM_U8, PCM_S16, PCM_S24, PCM_S32, PCM_F32, PCM_F64 = range(6)
_BYTES = (1, 2, 3, 4, 4, 8)
_KINDS = {(1, 8): PCM_U8, (1, 16): PCM_S16, (1, 24): PCM_S24, (1, 32): PCM_S32, (3, 32): PCM_F32, (3, 64): PCM_F64}


class WavInfo:
    __slots__ = ("fmt", "channels", "sample_rate", "data_offset", "n_frames")

    def __init__(self, fmt, channels, sample_rate, data_offset, n_frames):
        self.fmt, self.channels, self.sample_rate, self.data_offset, self.n_frames = fmt, channels, sample_rate, data_offset, n_frames

    @property
    def frame_bytes(self):
        return self.channels * _BYTES[self.fmt]


def read_wav_info(f, file_size):
    """Walk the RIFF chunks of an open binary file up to the `data` chunk (no sample is read)."""
    head = f.read(12)
    if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    fmt = None
    pos = 12
    while pos + 8 <= file_size:
        f.seek(pos)
        hdr = f.read(8)
        if len(hdr) < 8:
            break
        tag, size = hdr[:4], struct.unpack("<I", hdr[4:])[0]
        body = pos + 8
        if tag == b"fmt ":
            raw = f.read(min(size, 40))
            if len(raw) < 16:
                raise ValueError("truncated fmt chunk")
            code, ch, sr, _rate, _align, bits = struct.unpack("<HHIIHH", raw[:16])
            if code == 0xFFFE:                             # WAVE_FORMAT_EXTENSIBLE: the sub-format GUID starts with the real code
                if len(raw) < 26:
                    raise ValueError("truncated extensible fmt chunk")
                code = struct.unpack("<H", raw[24:26])[0]
            fmt = (code, ch, sr, bits)
        elif tag == b"data":
            if fmt is None:
                raise ValueError("data chunk before fmt chunk")
            code, ch, sr, bits = fmt
            if (code, bits) not in _KINDS or ch < 1:
                raise ValueError(f"unsupported WAVE encoding (format tag {code}, {bits} bits, {ch} channels)")
            kind = _KINDS[(code, bits)]
            size = min(size, file_size - body)
            return WavInfo(kind, ch, sr, body, size // (ch * _BYTES[kind]))
        pos = body + size + (size & 1)
    raise ValueError("no data chunk")


def _read_pcm(path, offset=0, duration=0):
    """-> (WavInfo, bytes of the selected frames).  offset / duration in seconds as in from_file (audio.py:310-315)."""
    try:
        size = os.path.getsize(path)
        with open(path, "rb") as f:
            info = read_wav_info(f, size)
            first = int(offset * info.sample_rate) if offset > 0 else 0
            first = min(first, info.n_frames)
            count = info.n_frames - first
            if duration > 0:
                count = min(count, int(duration * info.sample_rate))
            f.seek(info.data_offset + first * info.frame_bytes)
            raw = f.read(count * info.frame_bytes)
    except (OSError, ValueError, struct.error) as e:
        # same exception type and wording as the reference when neither backend can read the file (audio.py:344-346)
        raise Exception(f"Your audio file {path} could not be decoded. We tried using the RIFF/WAVE reader of desta_b200 ({e}).") from e
    info.n_frames = len(raw) // info.frame_bytes
    return info, raw


def load_waveforms(paths, device="cuda", target_sr=16000, n_samples=N_SAMPLES, out=None, pin=True):
    """Decode, down-mix and resample a batch of files straight into the padded waveform buffer of the log-mel kernel.
    Returns (wave [N, n_samples] f32 on `device`, lengths [N] int32 host tensor of valid samples per clip)."""
    device = torch.device(device)
    n = len(paths)
    if out is None:
        out = torch.empty(n, n_samples, dtype=torch.float32, device=device)
    assert out.shape[0] >= n and out.shape[1] == n_samples and out.is_cuda
    clips, total = [], 0
    for p in paths:
        info, raw = _read_pcm(p)
        clips.append((info, raw, total))
        total += (len(raw) + 15) & ~15            # 16-byte aligned clip starts: every sample format stays naturally aligned
    host = torch.empty(max(total, 16), dtype=torch.uint8, pin_memory=pin and torch.cuda.is_available())
    hv = host.numpy()
    for info, raw, off in clips:
        hv[off:off + len(raw)] = np.frombuffer(raw, np.uint8)
    pcm = host.to(device, non_blocking=True)
    groups = {}
    for row, (info, raw, off) in enumerate(clips):
        groups.setdefault((info.fmt, info.channels, info.sample_rate), []).append((off, info.n_frames, row))
    lengths = torch.zeros(n, dtype=torch.int32)
    for (fmt, ch, sr), members in groups.items():
        got = ops.decode_resample(pcm, members, fmt, ch, sr, target_sr, out)
        for (_, _, row), m in zip(members, got):
            lengths[row] = m
    pcm.record_stream(torch.cuda.current_stream(device))
    return out[:n], lengths


class DeviceWave(np.ndarray):
    """Host copy of a waveform that remembers the device tensor it came from, so that the processor shim
    (desta_b200.processor.WaveformProcessor) can assemble the batch without a second host -> device copy.

    Only the object `AudioSegment.samples` returns carries the tensor: anything derived from it (`w * 0.5`, `w[::-1]`,
    `np.clip(w, ...)`, slices, copies) is a plain host array again as far as the processor is concerned, and the carrier
    itself is read-only, so the two copies cannot drift apart silently - a caller that edits the samples gets its edited
    host array uploaded like any other."""

    def __new__(cls, host, device_tensor):
        obj = np.asarray(host).view(cls)
        obj.device_tensor = device_tensor
        obj.flags.writeable = False
        return obj

    def __array_finalize__(self, obj):
        self.device_tensor = None


class AudioSegment:
    """Drop-in for the subset of desta.utils.audio.AudioSegment the model path uses: `from_file(...)` and the `samples`,
    `sample_rate`, `num_samples`, `duration`, `orig_sr` accessors.  Samples live on the GPU; `.samples` returns a host
    float32 copy (the reference's VAD and `tolist()` callers need one)."""

    def __init__(self, device_samples, sample_rate, orig_sr=None):
        self._device_samples = device_samples
        self._sample_rate = sample_rate
        self._orig_sr = orig_sr if orig_sr is not None else sample_rate
        self._host = None

    @classmethod
    def from_file(cls, audio_file, target_sr=None, int_values=False, offset=0, duration=0, trim=False, trim_ref=np.max,
                  trim_top_db=60, trim_frame_length=2048, trim_hop_length=512, orig_sr=None, channel_selector=None,
                  normalize_db=None, ref_channel=None, device="cuda"):
        if isinstance(audio_file, list):
            raise NotImplementedError("lists of single-channel files (from_file_list) are outside the accelerated path")
        if int_values or trim or normalize_db is not None:
            raise NotImplementedError("int_values / trim / normalize_db are not used on the model path (modeling_desta25.py:1553-1557)")
        info, raw = _read_pcm(audio_file, offset, duration)
        if info.channels > 1 and channel_selector != "average":
            raise NotImplementedError(f"channel_selector={channel_selector!r} on a {info.channels}-channel file: the model path "
                                      f"always down-mixes with 'average'")
        if info.channels == 1 and channel_selector not in (None, 0, "average"):
            # select_channels on one-dimensional input (audio.py:78-83)
            raise ValueError('Input signal is one-dimensional, channel selector (%s) cannot not be used.', str(channel_selector))
        device = torch.device(device)
        sr_out = info.sample_rate if target_sr is None else int(target_sr)
        g = np.gcd(info.sample_rate, sr_out)
        n_out = -(-info.n_frames * (sr_out // g) // (info.sample_rate // g))
        out = torch.empty(1, max(n_out, 1), dtype=torch.float32, device=device)
        buf = np.frombuffer(raw, np.uint8) if len(raw) else np.zeros(16, np.uint8)
        pcm = torch.from_numpy(buf.copy()).to(device)
        ops.decode_resample(pcm, [(0, info.n_frames, 0)], info.fmt, info.channels, info.sample_rate, sr_out, out)
        return cls(out[0, :n_out], sr_out, orig_sr=orig_sr if orig_sr is not None else sr_out)

    @property
    def device_samples(self):
        return self._device_samples

    @property
    def samples(self):
        if self._host is None:
            self._host = DeviceWave(self._device_samples.cpu().numpy(), self._device_samples)
        return self._host

    @property
    def sample_rate(self):
        return self._sample_rate

    @property
    def num_samples(self):
        return int(self._device_samples.shape[0])

    @property
    def num_channels(self):
        return 1

    @property
    def duration(self):
        return self.num_samples / float(self._sample_rate)

    @property
    def orig_sr(self):
        return self._orig_sr

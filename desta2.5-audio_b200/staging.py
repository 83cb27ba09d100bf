"""Collate-side waveform staging (scope row n1: the step BEFORE the path).

The reference's collator decodes each clip, converts it to a Python list (`feature.tolist()`), rebuilds arrays inside
the HF feature extractor and computes the log-mel on the CPU in DataLoader workers; the main process then moves a
[N, 128, 3000] fp32 tensor to the GPU (ref: desta/trainer/data/simple_dataset.py:139-160, 239-243;
desta/models/modeling_desta25.py:1570-1571).  With the log-mel on the device the input of a step is the padded
waveform, and the only host work left is getting 1.92 MB per clip into HBM without stalling the GPU:

  * `WaveformStager.fill(slot, clips)` pads / truncates the clips of the NEXT batch straight into a pinned host slot
    (one copy per clip, no intermediate list / array / unpinned tensor),
  * `submit(slot)` starts the H2D copy on a dedicated copy stream,
  * `acquire(slot)` makes the compute stream wait for that copy and returns the device buffer; `release(slot)` hands
    the buffers back once the kernels that read them have been enqueued.

Two slots give a double buffer: the copy of batch k+1 overlaps the kernels of batch k.  torch supplies pinned memory,
streams and events only.

`PcmStager` is the same double buffer one step further upstream (scope rows n1 + n3): its slots hold the RAW interleaved
PCM bytes of the files (read straight into pinned memory, half the bytes of an fp32 waveform for 16-bit audio), and
`acquire` runs the fused decode + channel-average + resample + pad kernel (`desta_decode_resample`) on the compute stream,
so neither a float conversion nor a resampler runs on the host (ref: desta/utils/audio.py:245-361 as called at
desta/trainer/data/simple_dataset.py:145-149).
"""
from typing import List, Sequence, Tuple

import numpy as np
import torch

from .processor import N_SAMPLES, pad_or_truncate_batch


class WaveformStager:
    def __init__(self, n_clips: int, device, slots: int = 2, pin: bool = True):
        self.n_clips, self.device, self.slots = n_clips, torch.device(device), slots
        cuda = self.device.type == "cuda"
        self.host = [torch.zeros(n_clips, N_SAMPLES, dtype=torch.float32) for _ in range(slots)]
        self.host_len = [torch.zeros(n_clips, dtype=torch.int32) for _ in range(slots)]
        if pin and cuda:
            self.host = [t.pin_memory() for t in self.host]
            self.host_len = [t.pin_memory() for t in self.host_len]
        self._host_np = [t.numpy() for t in self.host]
        self._len_np = [t.numpy() for t in self.host_len]
        self.count = [0] * slots
        if cuda:
            self.dev = [torch.empty(n_clips, N_SAMPLES, dtype=torch.float32, device=self.device) for _ in range(slots)]
            self.dev_len = [torch.empty(n_clips, dtype=torch.int32, device=self.device) for _ in range(slots)]
            self.copy_stream = torch.cuda.Stream(self.device)
            self.ready = [torch.cuda.Event() for _ in range(slots)]
            self.free = [torch.cuda.Event() for _ in range(slots)]
            for e in self.free:
                e.record(torch.cuda.current_stream(self.device))
        self.h2d_bytes = 0

    def fill(self, slot: int, clips: Sequence) -> int:
        """Pad / truncate up to n_clips clips (arrays, lists or a 2-D array) into pinned slot `slot`; returns the count.
        The slot's previous H2D copy must have completed (acquire() of that slot was called)."""
        if self.device.type == "cuda":
            self.ready[slot].synchronize()       # no-op unless the host buffer is still being read by the copy engine
        n = len(clips)
        if n > self.n_clips:
            raise ValueError(f"{n} clips do not fit a stager slot of {self.n_clips}")
        pad_or_truncate_batch(clips, out=self._host_np[slot], lengths=self._len_np[slot])
        self.count[slot] = n
        return n

    def submit(self, slot: int):
        """Start the asynchronous H2D copy of slot `slot` on the copy stream."""
        n = self.count[slot]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[slot])       # kernels of the batch that used this device buffer are done
            self.dev[slot][:n].copy_(self.host[slot][:n], non_blocking=True)
            self.dev_len[slot][:n].copy_(self.host_len[slot][:n], non_blocking=True)
            self.ready[slot].record(self.copy_stream)
        self.h2d_bytes += n * (N_SAMPLES * 4 + 4)

    def acquire(self, slot: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device waveform [n, 480000] fp32 and valid lengths [n] int32 of slot `slot`, ordered after its H2D copy on
        the CURRENT stream (no host synchronisation)."""
        torch.cuda.current_stream(self.device).wait_event(self.ready[slot])
        n = self.count[slot]
        return self.dev[slot][:n], self.dev_len[slot][:n]

    def release(self, slot: int):
        """Call after the kernels reading the slot's device buffer have been enqueued on the current stream."""
        self.free[slot].record(torch.cuda.current_stream(self.device))


def stage_batches(stager: WaveformStager, batches: List[Sequence]):
    """Generator: yields (device_waveforms, lengths) per batch with the next batch's copy already in flight."""
    if not batches:
        return
    stager.fill(0, batches[0])
    stager.submit(0)
    for k in range(len(batches)):
        slot = k % stager.slots
        if k + 1 < len(batches):
            nxt = (k + 1) % stager.slots
            stager.fill(nxt, batches[k + 1])
            stager.submit(nxt)
        wave, lens = stager.acquire(slot)
        yield wave, lens
        stager.release(slot)


class PcmStager:
    """Double-buffered file -> device waveform staging: `fill(slot, paths)` reads the PCM payload of up to n_clips RIFF/WAVE
    files into pinned slot `slot`, `submit` starts its H2D copy on the copy stream, `acquire` decodes / down-mixes /
    resamples it into the slot's [n, 480000] fp32 device buffer on the current stream and returns (waveforms, lengths)."""

    def __init__(self, n_clips: int, device, slot_bytes: int = None, slots: int = 2, target_sr: int = 16000, pin: bool = True):
        self.n_clips, self.device, self.slots, self.target_sr = n_clips, torch.device(device), slots, target_sr
        cuda = self.device.type == "cuda"
        # default capacity: 30 s of 48 kHz 16-bit stereo per clip
        self.slot_bytes = int(slot_bytes) if slot_bytes else n_clips * 30 * 48000 * 4
        self.host = [torch.zeros(self.slot_bytes, dtype=torch.uint8) for _ in range(slots)]
        if pin and cuda:
            self.host = [t.pin_memory() for t in self.host]
        self._host_np = [t.numpy() for t in self.host]
        self.clips = [[] for _ in range(slots)]       # per slot: (fmt, channels, sample_rate, byte offset, n_frames)
        self.used = [0] * slots
        if cuda:
            self.pcm = [torch.empty(self.slot_bytes, dtype=torch.uint8, device=self.device) for _ in range(slots)]
            self.dev = [torch.empty(n_clips, N_SAMPLES, dtype=torch.float32, device=self.device) for _ in range(slots)]
            self.copy_stream = torch.cuda.Stream(self.device)
            self.ready = [torch.cuda.Event() for _ in range(slots)]
            self.free = [torch.cuda.Event() for _ in range(slots)]
            for e in self.free:
                e.record(torch.cuda.current_stream(self.device))
        self.h2d_bytes = 0

    def fill(self, slot: int, paths: Sequence[str]) -> int:
        import os

        from .audio import read_wav_info
        if self.device.type == "cuda":
            self.ready[slot].synchronize()
        if len(paths) > self.n_clips:
            raise ValueError(f"{len(paths)} clips do not fit a stager slot of {self.n_clips}")
        import struct
        buf, off, table = self._host_np[slot], 0, []
        for path in paths:
            try:
                size = os.path.getsize(path)
                with open(path, "rb") as f:
                    info = read_wav_info(f, size)
                    nbytes = info.n_frames * info.frame_bytes
                    if off + nbytes <= self.slot_bytes:
                        f.seek(info.data_offset)
                        got = f.readinto(memoryview(buf[off:off + nbytes]))
            except (OSError, ValueError, struct.error) as e:
                # same exception type and wording as the reference when no backend can read the file (audio.py:344-346)
                raise Exception(f"Your audio file {path} could not be decoded. We tried using the RIFF/WAVE reader of desta_b200 ({e}).") from e
            if off + nbytes > self.slot_bytes:
                raise ValueError(f"slot of {self.slot_bytes} bytes is too small for this batch (needs >= {off + nbytes})")
            table.append((info.fmt, info.channels, info.sample_rate, off, got // info.frame_bytes))
            off = (off + nbytes + 15) & ~15
        self.clips[slot], self.used[slot] = table, off
        return len(table)

    def submit(self, slot: int):
        n = self.used[slot]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[slot])
            if n:
                self.pcm[slot][:n].copy_(self.host[slot][:n], non_blocking=True)
            self.ready[slot].record(self.copy_stream)
        self.h2d_bytes += n

    def acquire(self, slot: int) -> Tuple[torch.Tensor, torch.Tensor]:
        from . import ops
        torch.cuda.current_stream(self.device).wait_event(self.ready[slot])
        table = self.clips[slot]
        lengths = torch.zeros(len(table), dtype=torch.int32)
        groups = {}
        for row, (fmt, ch, sr, off, frames) in enumerate(table):
            groups.setdefault((fmt, ch, sr), []).append((off, frames, row))
        for (fmt, ch, sr), members in groups.items():
            got = ops.decode_resample(self.pcm[slot], members, fmt, ch, sr, self.target_sr, self.dev[slot])
            for (_, _, row), m in zip(members, got):
                lengths[row] = m
        return self.dev[slot][:len(table)], lengths

    def release(self, slot: int):
        self.free[slot].record(torch.cuda.current_stream(self.device))


def stage_files(stager: PcmStager, batches: List[Sequence[str]]):
    """Generator over batches of file paths: yields (device_waveforms [n, 480000], lengths) with the next batch's bytes
    already on their way to the device."""
    if not batches:
        return
    stager.fill(0, batches[0])
    stager.submit(0)
    for k in range(len(batches)):
        slot = k % stager.slots
        if k + 1 < len(batches):
            nxt = (k + 1) % stager.slots
            stager.fill(nxt, batches[k + 1])
            stager.submit(nxt)
        wave, lens = stager.acquire(slot)
        yield wave, lens
        stager.release(slot)

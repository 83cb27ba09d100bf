"""Processor shim (boundary b1): keeps the `processor(list_of_waveforms, sampling_rate=16000, return_tensors="pt")
.input_features` call of the reference (ref: desta/models/modeling_desta25.py:1570,1583;
desta/trainer/data/simple_dataset.py:239-243) but returns the right-padded / truncated WAVEFORM [N, 480000] fp32
instead of a CPU log-mel: the log-mel runs on the GPU inside WhisperPerceptionB200 (DataLoader workers cannot touch
CUDA).  Everything else is delegated to the wrapped WhisperProcessor (batch_decode, tokenizer, feature_extractor...).
"""
import numpy as np
import torch

N_SAMPLES = 480000
SAMPLING_RATE = 16000


def pad_or_truncate_batch(raw_speech, out=None, lengths=None):
    """$TF/models/whisper/feature_extraction_whisper.py:296-320: first 480000 samples, zero right-padding.
    Returns (waveforms [n, 480000] fp32, valid lengths [n] int32); with `out` / `lengths` (e.g. rows of a pinned staging
    slot, see desta_b200.staging) the clips are written in place and views of the first n rows are returned."""
    if isinstance(raw_speech, np.ndarray) and raw_speech.ndim == 2:
        clips = list(raw_speech)
    elif isinstance(raw_speech, (list, tuple)) and len(raw_speech) and isinstance(raw_speech[0], (np.ndarray, list, tuple)):
        clips = list(raw_speech)
    elif isinstance(raw_speech, (list, tuple)) and len(raw_speech) == 0:
        clips = []
    else:
        clips = [raw_speech]
    n = len(clips)
    if out is None:
        out = np.zeros((n, N_SAMPLES), dtype=np.float32)
    else:
        assert out.dtype == np.float32 and out.shape[0] >= n and out.shape[1] == N_SAMPLES
    if lengths is None:
        lengths = np.zeros(n, dtype=np.int32)
    for i, c in enumerate(clips):
        c = np.asarray(c, dtype=np.float32)
        if c.ndim != 1:
            raise ValueError("Only mono-channel audio is supported")
        m = min(c.shape[0], N_SAMPLES)
        out[i, :m] = c[:m]
        out[i, m:] = 0.0
        lengths[i] = m
    return out[:n], lengths[:n]


def _device_batch(raw_speech):
    """Clips decoded by desta_b200.audio.AudioSegment carry their device tensor: pad / truncate there (no second H2D copy)."""
    if not isinstance(raw_speech, (list, tuple)) or not raw_speech:
        return None
    devs = [getattr(c, "device_tensor", None) for c in raw_speech]
    if any(d is None for d in devs):
        return None
    wave = torch.zeros(len(devs), N_SAMPLES, dtype=torch.float32, device=devs[0].device)
    lengths = torch.zeros(len(devs), dtype=torch.int32)
    for i, d in enumerate(devs):
        m = min(d.shape[0], N_SAMPLES)
        wave[i, :m] = d[:m]
        lengths[i] = m
    return wave, lengths


class _Features(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def to(self, *a, **k):
        return _Features({n: (v.to(*a, **k) if torch.is_tensor(v) else v) for n, v in self.items()})


class StagedWaveform(torch.Tensor):
    """The pinned host waveform batch a prefetching WaveformProcessor hands out.  It behaves like the host tensor it is; only
    `.to(<the staging device>)` differs: the H2D copy of this batch was started on the processor's copy stream when the batch
    was packed, so the call makes the current stream wait for that copy and returns the device buffer instead of copying
    again.  If the host tensor was modified in place since (version counter), or the target is anything else, the regular
    copy runs."""

    @staticmethod
    def wrap(host: torch.Tensor, slot: dict, owner: "WaveformProcessor") -> "StagedWaveform":
        t = host.as_subclass(StagedWaveform)
        t._slot, t._owner, t._staged_version = slot, owner, host._version
        return t

    def to(self, *args, **kwargs):
        slot = getattr(self, "_slot", None)
        if slot is not None and slot.get("dev") is not None and self._version == self._staged_version:
            target = kwargs.get("device", args[0] if args and not isinstance(args[0], torch.dtype) else None)
            dtype = kwargs.get("dtype", next((a for a in args if isinstance(a, torch.dtype)), None))
            if target is not None and not torch.is_tensor(target) and dtype in (None, torch.float32):
                dev = torch.device(target)
                if dev.type == "cuda" and (dev.index is None or dev.index == slot["dev"].device.index):
                    return self._owner._hand_over(slot, self.shape[0])
        return self.as_subclass(torch.Tensor).to(*args, **kwargs)


class WaveformProcessor:
    def __init__(self, wrapped=None, sampling_rate: int = SAMPLING_RATE):
        self._wrapped = wrapped
        self.sampling_rate = sampling_rate
        self._slots = None          # optional pinned output slots, see enable_pinned_staging

    def enable_pinned_staging(self, n_clips: int, slots: int = 3, device=None):
        """Pack batches of up to `n_clips` clips straight into rotating PINNED host slots (main process only: DataLoader
        workers have no CUDA context and keep the pageable default).  `input_features` is then a pinned tensor, so the
        caller's `.to(device, non_blocking=True)` (modeling_desta25.py:1571) is a genuinely asynchronous copy and no
        intermediate pageable [N, 480000] array is allocated per call.  A slot is reused `slots` calls later; a caller that
        enqueues work without ever synchronising should tell the processor when the slot's copy is ordered with
        `release_after(input_features)` - the slot is then not overwritten before that point of the stream has passed.

        With `device` (scope row n1) the processor also PREFETCHES: each slot owns a device buffer, the H2D copy of a batch
        starts on a copy stream of the processor as soon as the batch is packed (typically in the loader thread, while the
        previous batch computes), and `input_features.to(device)` hands that buffer over (see StagedWaveform) - the caller's
        code stays `processor(...).input_features.to(device)`.  The device buffer of a slot is rewritten `slots` batches
        later, after the work that was enqueued on the consumer's stream up to the NEXT hand-over has finished (a loop
        that keeps more than `slots - 1` batches in flight needs more slots)."""
        assert torch.cuda.is_available(), "pinned staging needs a CUDA context"
        self._slots = [dict(wave=torch.zeros(n_clips, N_SAMPLES, dtype=torch.float32).pin_memory(),
                            lens=torch.zeros(n_clips, dtype=torch.int32).pin_memory(), event=None, dev=None, ready=None, consumed=None)
                       for _ in range(slots)]
        self._next_slot = 0
        self._last_handed = None
        if device is not None:
            dev = torch.device(device)
            self._copy_stream = torch.cuda.Stream(dev)
            for sl in self._slots:
                sl["dev"] = torch.empty(n_clips, N_SAMPLES, dtype=torch.float32, device=dev)

    def _prefetch(self, sl, n):
        """Start the H2D copy of the first n rows of slot `sl` on the copy stream (called right after packing)."""
        dev = sl["dev"].device
        if sl["consumed"] is not None:
            self._copy_stream.wait_event(sl["consumed"])       # the kernels that read this device buffer have finished
            sl["consumed"] = None
        else:
            self._copy_stream.wait_stream(torch.cuda.current_stream(dev))   # first uses / unknown consumer: order behind everything
        with torch.cuda.stream(self._copy_stream):
            sl["dev"][:n].copy_(sl["wave"][:n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        sl["ready"] = ev
        sl["event"] = ev                                        # the pinned rows may be repacked once the copy has read them

    def _hand_over(self, sl, n):
        cur = torch.cuda.current_stream(sl["dev"].device)
        # everything enqueued so far includes all work on the batch handed over before this one: its buffer is free after it
        if self._last_handed is not None and self._last_handed is not sl:
            ev = torch.cuda.Event()
            ev.record(cur)
            self._last_handed["consumed"] = ev
        self._last_handed = sl
        cur.wait_event(sl["ready"])
        return sl["dev"][:n]

    def release_after(self, input_features, stream=None):
        """Record an event on `stream` (default: current) after which the pinned slot behind `input_features` may be reused."""
        if self._slots is None:
            return
        for sl in self._slots:
            if sl["wave"].data_ptr() == input_features.data_ptr():
                ev = torch.cuda.Event()
                ev.record(stream if stream is not None else torch.cuda.current_stream())
                sl["event"] = ev

    def __call__(self, raw_speech=None, sampling_rate=None, return_tensors="pt", **kwargs):
        if raw_speech is None:
            raw_speech = kwargs.pop("audio", None)
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            # same error as WhisperFeatureExtractor.__call__ (feature_extraction_whisper.py:261-267)
            raise ValueError(
                f"The model corresponding to this feature extractor: WhisperFeatureExtractor was trained using a"
                f" sampling rate of {self.sampling_rate}. Please make sure that the provided `raw_speech` input"
                f" was sampled with {self.sampling_rate} and not {sampling_rate}.")
        on_device = _device_batch(raw_speech) if return_tensors == "pt" else None
        if on_device is not None:
            return _Features(input_features=on_device[0], num_samples=on_device[1])
        if self._slots is not None and return_tensors == "pt":
            n = len(raw_speech) if isinstance(raw_speech, (list, tuple)) or getattr(raw_speech, "ndim", 1) == 2 else 1
            sl = self._slots[self._next_slot]
            if n <= sl["wave"].shape[0]:
                self._next_slot = (self._next_slot + 1) % len(self._slots)
                if sl["event"] is not None:
                    sl["event"].synchronize()          # the copy that read this slot has been passed by the GPU
                    sl["event"] = None
                wave, lengths = pad_or_truncate_batch(raw_speech, out=sl["wave"].numpy(), lengths=sl["lens"].numpy())
                feats = sl["wave"][:len(wave)]
                if sl["dev"] is not None:
                    self._prefetch(sl, len(wave))
                    feats = StagedWaveform.wrap(feats, sl, self)
                return _Features(input_features=feats, num_samples=sl["lens"][:len(wave)])
        wave, lengths = pad_or_truncate_batch(raw_speech)
        feats = torch.from_numpy(wave) if return_tensors == "pt" else wave
        return _Features(input_features=feats, num_samples=torch.from_numpy(lengths) if return_tensors == "pt" else lengths)

    def mel_features(self, *a, **k):
        """Genuine log-mel through the wrapped processor (the ASR pass at modeling_desta25.py:1583-1590 needs it)."""
        if self._wrapped is None:
            raise RuntimeError("no wrapped WhisperProcessor")
        return self._wrapped(*a, **k)

    def __getattr__(self, name):
        if name.startswith("_") or self._wrapped is None:
            raise AttributeError(name)
        return getattr(self._wrapped, name)

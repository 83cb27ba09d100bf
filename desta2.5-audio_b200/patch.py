"""Install the B200 path into a reference DeSTA25AudioModel instance without touching its public API
(`from_pretrained`, `generate`, `forward`, `_generate_step`, `state_dict` keep working; checkpoint keys unchanged).

    from desta.models.modeling_desta25 import DeSTA25AudioModel
    import desta_b200
    model = DeSTA25AudioModel.from_pretrained(...).cuda()
    desta_b200.install(model)          # perception + splice now run on libdesta_b200.so
"""
import sys
import types

from .perception import WhisperPerceptionB200
from .processor import WaveformProcessor
from .splice import prepare_inputs_for_llm


def install(model, wrap_processor: bool = True, patch_audio: bool = False, staging_clips: int = 0):
    """staging_clips > 0 (model on a CUDA device): the wrapped processor - also the one `generate()` creates later - packs
    batches of up to that many clips into pinned slots and starts their host-to-device copy at packing time
    (`WaveformProcessor.enable_pinned_staging(staging_clips, device=...)`, scope row n1); larger batches take the regular path.
    patch_audio=True also rebinds the `AudioSegment` name that `generate()` uses (modeling_desta25.py:13, :1553-1557) to
    desta_b200.audio.AudioSegment, so RIFF/WAVE files are decoded, down-mixed and resampled to 16 kHz on the GPU.  It is
    opt-in because it is module-wide and because other containers (flac, mp3 ...) need the reference's libsndfile path."""
    if getattr(model.config, "connector_mode", "qformer_1") != "qformer_1":
        raise NotImplementedError("desta_b200 covers connector_mode='qformer_1' (ORCA modifies LLM decoding: out of scope)")
    if not isinstance(model.perception, WhisperPerceptionB200):        # idempotent: a second install() changes nothing
        model.perception = WhisperPerceptionB200.from_reference(model.perception)

    def _prepare_inputs_for_llm(self, input_ids, attention_mask, batch_features, batch_transcription_ids, batch_start_positions):
        return prepare_inputs_for_llm(self.llm_model.model.embed_tokens.weight, self.perception, input_ids, attention_mask,
                                      batch_features, batch_transcription_ids, batch_start_positions)

    model._prepare_inputs_for_llm = types.MethodType(_prepare_inputs_for_llm, model)
    if hasattr(model.perception.whisper, "generate"):
        # ASR pass of generate() (:1582-1601): encoder through the same kernels, shared with the perception call
        model.perception.wrap_whisper_generate()
    if patch_audio:
        from .audio import AudioSegment
        mod = sys.modules.get(type(model).__module__)
        if mod is None or not hasattr(mod, "AudioSegment"):
            raise RuntimeError(f"{type(model).__module__} has no AudioSegment to replace")
        mod.AudioSegment = AudioSegment
    if wrap_processor:
        _wrap_processor(model, staging_clips)
    return model


def _staged(model, proc, staging_clips):
    if staging_clips > 0 and proc._slots is None:
        dev = next(model.perception.parameters()).device
        if dev.type == "cuda":
            proc.enable_pinned_staging(staging_clips, device=dev)
    return proc


def _wrap_processor(model, staging_clips: int = 0):
    """Boundary b1.  The reference creates `self.processor` lazily: `generate()` calls `_setup_generation()` on its first
    use (modeling_desta25.py:1514-1515 -> :1467-1482, `AutoProcessor.from_pretrained` at :1475), i.e. usually AFTER
    `install(model)`.  So besides wrapping a processor that already exists, `_setup_generation` itself is wrapped: whatever
    processor it creates is handed back as a WaveformProcessor, and `generate()`'s
    `self.processor(batch_features, sampling_rate=16000, return_tensors="pt").input_features` (:1570, :1583) yields padded
    waveforms for the device log-mel instead of running the CPU WhisperFeatureExtractor."""
    if getattr(model, "processor", None) is not None and not isinstance(model.processor, WaveformProcessor):
        model.processor = _staged(model, WaveformProcessor(model.processor), staging_clips)
    setup = getattr(model, "_setup_generation", None)
    if setup is None or getattr(setup, "_desta_b200_wrapped", False):
        return

    def _setup_generation(self, *args, **kwargs):
        out = setup(*args, **kwargs)
        proc = getattr(self, "processor", None)
        if proc is not None and not isinstance(proc, WaveformProcessor):
            self.processor = _staged(self, WaveformProcessor(proc), staging_clips)
        return out

    bound = types.MethodType(_setup_generation, model)
    bound.__func__._desta_b200_wrapped = True
    model._setup_generation = bound

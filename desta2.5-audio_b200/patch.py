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


def install(model, wrap_processor: bool = True, patch_audio: bool = False):
    """patch_audio=True also rebinds the `AudioSegment` name that `generate()` uses (modeling_desta25.py:13, :1553-1557) to
    desta_b200.audio.AudioSegment, so RIFF/WAVE files are decoded, down-mixed and resampled to 16 kHz on the GPU.  It is
    opt-in because it is module-wide and because other containers (flac, mp3 ...) need the reference's libsndfile path."""
    if getattr(model.config, "connector_mode", "qformer_1") != "qformer_1":
        raise NotImplementedError("desta_b200 covers connector_mode='qformer_1' (ORCA modifies LLM decoding: out of scope)")
    ref_perception = model.perception
    model.perception = WhisperPerceptionB200.from_reference(ref_perception)

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
    if wrap_processor and getattr(model, "processor", None) is not None and not isinstance(model.processor, WaveformProcessor):
        model.processor = WaveformProcessor(model.processor)
    return model

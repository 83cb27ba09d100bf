"""Import the reference's own model code without network or HF cache.  TEST INFRASTRUCTURE ONLY.

Where the reference comes from: `/root/reference` in the build container, else `baseline/_ref` (the UNMODIFIED reference
installed with `pip install --no-deps --target baseline/_ref`, git-ignored, travels to the GPU box with the snapshot).
Used by oracle/make_golden.py (fixtures), tests/ (drop-in tests on the real `DeSTA25AudioModel`) and the CPU legs of
bench.py; no product module imports it.

Shims: (1) stand-ins for `librosa` / `soundfile`, inserted AFTER importing transformers
(ref: desta/utils/audio.py:41-43,56 import them at module import) — `soundfile` reads RIFF/WAVE PCM16 with stdlib `wave`
and libsndfile's float32 scaling, which is all the tests feed it; (2) from_pretrained of AutoConfig /
WhisperForConditionalGeneration / AutoModelForCausalLM build random-init models from explicit configs
(ref: modeling_desta25.py:505-506, 673-674, 713-718 would hit the hub); (3) WhisperEncoderLayer.forward returns a
1-tuple as under transformers 4.x when called from the reference's layer loop, which indexes it with [0] (:585).
`offline_generation(model, ...)` additionally replaces what `generate()` downloads (tokenizer, processor, Silero VAD:
modeling_desta25.py:1467-1488) by local stand-ins so that the reference's own `generate()` runs end to end offline."""
import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = ("/root/reference", os.path.join(_ROOT, "baseline", "_ref"))


def ref_root():
    """Directory that holds the reference's `desta` package, or None."""
    for c in _CANDIDATES:
        if os.path.exists(os.path.join(c, "desta", "models", "modeling_desta25.py")):
            return c
    return None


REF_ROOT = ref_root() or _CANDIDATES[0]


class _WavSoundFile:
    """Stand-in for soundfile.SoundFile on RIFF/WAVE PCM16 files (the subset the reference's AudioSegment.from_file
    touches: context manager, .samplerate, .seek(frames), .read(frames=-1, dtype)); float32 = int16 / 32768 as libsndfile."""

    def __init__(self, path, mode="r"):
        import wave
        self._w = wave.open(path if isinstance(path, str) else path, "rb")
        if self._w.getsampwidth() != 2:
            raise RuntimeError("stand-in soundfile reads PCM16 only")
        self.samplerate = self._w.getframerate()
        self.channels = self._w.getnchannels()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._w.close()
        return False

    def seek(self, frames):
        self._w.setpos(int(frames))

    def read(self, frames=-1, dtype="float32"):
        import numpy as np
        n = self._w.getnframes() - self._w.tell() if frames is None or frames < 0 else int(frames)
        raw = np.frombuffer(self._w.readframes(n), "<i2")
        x = raw.astype(np.float32) / np.float32(32768) if dtype == "float32" else raw.astype(np.int32) << 16
        return x.reshape(-1, self.channels) if self.channels > 1 else x


def load_reference(cfg, max_target_positions=8):
    import torch
    import transformers  # noqa: F401  (must be imported before stubbing)
    import transformers.audio_utils  # noqa: F401
    from transformers import (AutoConfig, AutoModelForCausalLM, AutoProcessor, AutoTokenizer,  # noqa: F401
                              LlamaConfig, LlamaForCausalLM, WhisperConfig, WhisperForConditionalGeneration)
    from transformers.models.whisper import modeling_whisper

    root = ref_root()
    if root is None:
        raise RuntimeError("the reference is neither at /root/reference nor installed under baseline/_ref")
    for name in ("librosa", "soundfile"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.available_formats = lambda: {"WAV": "WAV (Microsoft)"}
            m.SoundFile = _WavSoundFile
            sys.modules[name] = m
    if root not in sys.path:
        sys.path.insert(0, root)

    wcfg = WhisperConfig(
        num_mel_bins=cfg.n_mels, d_model=cfg.d_model, encoder_layers=cfg.n_layers,
        encoder_attention_heads=cfg.n_heads, encoder_ffn_dim=cfg.ffn, decoder_layers=1,
        decoder_attention_heads=cfg.n_heads, decoder_ffn_dim=64, max_source_positions=cfg.max_pos,
        vocab_size=64, max_target_positions=max_target_positions, pad_token_id=0, bos_token_id=1, eos_token_id=2,
        decoder_start_token_id=1)
    lcfg = LlamaConfig(hidden_size=cfg.d_llm, intermediate_size=64, num_hidden_layers=1,
                       num_attention_heads=2, num_key_value_heads=2, vocab_size=cfg.vocab,
                       max_position_embeddings=1024)

    def auto_config(model_id, *a, **k):
        return wcfg if "whisper" in model_id else lcfg

    AutoConfig.from_pretrained = staticmethod(auto_config)
    WhisperForConditionalGeneration.from_pretrained = classmethod(
        lambda cls, *a, **k: WhisperForConditionalGeneration(wcfg))
    AutoModelForCausalLM.from_pretrained = staticmethod(
        lambda *a, **k: LlamaForCausalLM(lcfg).to(torch.bfloat16))

    if not getattr(modeling_whisper.WhisperEncoderLayer, "_tuple_shim", False):
        orig = modeling_whisper.WhisperEncoderLayer.forward

        def fwd(self, *a, **k):
            out = orig(self, *a, **k)
            if isinstance(out, tuple):
                return out
            # 4.x tuple only for the reference's manual layer loop; HF's own WhisperEncoder.forward (the ASR pass of
            # generate() goes through it) expects the 5.x bare tensor
            f = sys._getframe(1)
            for _ in range(6):
                if f is None:
                    break
                if f.f_code.co_filename.endswith("modeling_desta25.py"):
                    return (out,)
                f = f.f_back
            return out

        modeling_whisper.WhisperEncoderLayer.forward = fwd
        modeling_whisper.WhisperEncoderLayer._tuple_shim = True

    from desta.models import modeling_desta25 as ref
    return ref


# ------------------------------------------------------------------------------------------------------------------
# offline stand-ins for what generate() downloads
# ------------------------------------------------------------------------------------------------------------------
WORDS = ["hello", "this", "is", "my", "audio", "help", "me", "transcribe", "what", "sound", "focus", "on", "the", "clips",
         "and", "instructions", "a", "dog", "cat", "bark", "meow", "speech", "noise"] + [f"w{i}" for i in range(40)]
CHAT_TEMPLATE = ("{% for m in messages %}<|{{ m['role'] }}|> {{ m['content'] }} <|eot|> {% endfor %}"
                 "{% if add_generation_prompt %}<|assistant|> {% endif %}")


def offline_tokenizer(placeholder_token):
    """A whitespace word-level tokenizer (tokenizers library, built in memory) with the special tokens generate() relies on:
    tokenize -> convert_tokens_to_string -> tokenize round-trips exactly (modeling_desta25.py:1621-1643)."""
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    specials = ["<unk>", "<|eos|>", "<|system|>", "<|user|>", "<|assistant|>", "<|eot|>", "<start_audio>", "<end_audio>",
                placeholder_token]
    vocab = {t: i for i, t in enumerate(specials + WORDS)}
    tk = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    tk.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    tk.decoder = decoders.Fuse()          # placeholder; convert_tokens_to_string is overridden below
    fast = PreTrainedTokenizerFast(tokenizer_object=tk, unk_token="<unk>", eos_token="<|eos|>")
    fast.add_special_tokens({"additional_special_tokens": specials[2:]})
    fast.chat_template = CHAT_TEMPLATE
    fast.convert_tokens_to_string = lambda tokens: " ".join(tokens)
    return fast


class OfflineWhisperProcessor:
    """What `AutoProcessor.from_pretrained(encoder_model_id)` yields, minus the hub: the genuine HF WhisperFeatureExtractor
    (the reference's CPU log-mel, $TF/models/whisper/feature_extraction_whisper.py) plus a `batch_decode` over WORDS."""

    def __init__(self, n_mels):
        from transformers import WhisperFeatureExtractor
        self.feature_extractor = WhisperFeatureExtractor(feature_size=n_mels)
        self.calls = 0

    def __call__(self, audio=None, sampling_rate=None, return_tensors=None, **kw):
        self.calls += 1
        return self.feature_extractor(audio, sampling_rate=sampling_rate, return_tensors=return_tensors, **kw)

    def batch_decode(self, sequences, skip_special_tokens=True):
        return [" ".join(WORDS[int(t) % len(WORDS)] for t in seq if int(t) > 2) for seq in sequences]


def offline_generation(model, n_mels, is_speech=None):
    """Patch the `from_pretrained` factories `_setup_generation` calls and the Silero VAD loader on `model`'s class module.
    `is_speech(feature) -> bool` decides the VAD answer (default: mean |x| > 0.02)."""
    import numpy as np
    from transformers import AutoProcessor, AutoTokenizer
    AutoTokenizer.from_pretrained = staticmethod(lambda *a, **k: offline_tokenizer(model.placeholder_token))
    AutoProcessor.from_pretrained = staticmethod(lambda *a, **k: OfflineWhisperProcessor(n_mels))
    speech = is_speech or (lambda x: float(np.abs(np.asarray(x)).mean()) > 0.02)

    def _setup_vad():
        model.vad_model = object()
        model.get_speech_timestamps = lambda feature, vad: ([{"start": 0, "end": len(feature)}] if speech(feature) else [])

    model._setup_vad = _setup_vad
    return model

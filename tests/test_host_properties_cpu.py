"""Property tests (hypothesis) of the host-side integer logic: the product's functions against the oracle restatement on
random inputs, plus invariants the reference's algorithm implies (desta/models/modeling_desta25.py:99-123, :1645-1651;
$TF/models/whisper/feature_extraction_whisper.py:296-320).  Everything here is bit-exact integer / copy work."""
import numpy as np
from hypothesis import given, settings, strategies as st

from desta_b200 import splice as p_sp
from desta_b200.processor import N_SAMPLES, pad_or_truncate_batch
from oracle import splice as o_sp

LOC, PAD = 7, 0


@st.composite
def prompts(draw):
    n_loc = draw(st.integers(0, 5))
    toks = draw(st.lists(st.integers(10, 50), min_size=0, max_size=20))
    pos = sorted(draw(st.lists(st.integers(0, len(toks)), min_size=n_loc, max_size=n_loc)))
    for k, p in enumerate(pos):
        toks.insert(p + k, LOC)
    sizes = [draw(st.integers(1, 9)) for _ in range(n_loc + draw(st.integers(0, 2)))]      # spare entries stay in the lists
    trans = [draw(st.integers(0, 6)) for _ in range(len(sizes))]
    return toks, n_loc, sizes, trans


@settings(max_examples=200, deadline=None)
@given(prompts())
def test_start_positions_product_equals_oracle_and_invariants(case):
    toks, n_loc, sizes, trans = case
    a1, t1, a2, t2 = list(sizes), list(trans), list(sizes), list(trans)
    res_o, starts_o = o_sp.prepare_audio_context_and_start_positions(list(toks), LOC, a1, t1, PAD)
    res_p, starts_p = p_sp.prepare_audio_context_and_start_positions(list(toks), LOC, a2, t2, PAD)
    assert res_p == res_o and starts_p == starts_o and a1 == a2 and t1 == t2
    # one start per locator, in order; each locator became audio_size + transcription_size placeholders
    assert len(starts_p) == n_loc and starts_p == sorted(starts_p)
    assert len(res_p) == len(toks) - n_loc + sum(sizes[:n_loc]) + sum(trans[:n_loc])
    assert a2 == sizes[n_loc:] and t2 == trans[n_loc:]                  # consumed front to back (pop(0))
    for k, s in enumerate(starts_p):
        span = sizes[k] + trans[k]
        assert all(x == PAD for x in res_p[s:s + span])
    kept = [x for x in res_p if x != PAD]
    assert kept == [x for x in toks if x != LOC]                        # the text tokens survive in order
    # left padding shifts every start by the number of pad positions of the row
    mask = [0] * 3 + [1] * len(res_p)
    assert p_sp.apply_left_pad_offset(starts_p, mask) == [s + 3 for s in starts_p] == o_sp.left_pad_offset(starts_o, mask)


@settings(max_examples=30, deadline=None)
@given(st.lists(st.integers(0, 3), min_size=0, max_size=4), st.integers(0, 2 ** 31 - 1))
def test_pad_or_truncate_properties(kinds, seed):
    rng = np.random.default_rng(seed)
    lens = [(0, 1, N_SAMPLES, N_SAMPLES + 17)[k] if k != 1 else int(rng.integers(1, 5000)) for k in kinds]
    clips = [rng.standard_normal(n).astype(np.float32) for n in lens]
    out, lengths = pad_or_truncate_batch(clips)
    assert out.shape == (len(clips), N_SAMPLES) and out.dtype == np.float32
    for i, c in enumerate(clips):
        m = min(len(c), N_SAMPLES)
        assert lengths[i] == m and np.array_equal(out[i, :m], c[:m]) and not out[i, m:].any()
    # idempotent: padding the padded batch changes nothing
    out2, lengths2 = pad_or_truncate_batch(out) if len(clips) else (out, lengths)
    assert np.array_equal(out2, out)


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 5), st.integers(1, 4), st.sampled_from([8000, 16000, 22050, 44100, 48000]), st.integers(0, 300),
       st.booleans(), st.lists(st.integers(0, 23), min_size=0, max_size=3), st.integers(0, 120))
def test_wav_header_walk_product_equals_oracle(kind, ch, sr, frames, extensible, junk_sizes, cut):
    """RIFF chunk walk of the product (desta_b200.audio.read_wav_info, file based) against the oracle (bytes based) on
    random layouts: unknown chunks of odd and even sizes before `fmt ` and `data`, EXTENSIBLE headers, truncated payloads."""
    import io
    import struct

    import wavutil as W
    from desta_b200 import audio
    from oracle import audio_io as A
    import pytest
    x = np.zeros((frames, ch))
    data = W.wav_bytes(x, sr, kind, extensible=extensible and kind != W.U8)
    # splice junk chunks in front of the fmt chunk (12-byte RIFF header first)
    junk = b"".join(b"JUNK" + struct.pack("<I", n) + bytes(n) + (b"\x00" if n & 1 else b"") for n in junk_sizes)
    data = data[:12] + junk + data[12:]
    payload = frames * ch * (1, 2, 3, 4, 4, 8)[kind]
    payload += payload & 1                                  # RIFF pad byte after an odd data chunk
    data = data[:max(0, len(data) - cut)]                   # (a cut longer than the file leaves an empty file, not a negative slice)
    if cut > payload:
        # the cut reaches the chunk headers (data header, fmt body, junk ...): both walks refuse the file with ValueError,
        # which `_read_pcm` turns into the reference's "could not be decoded" exception (desta/utils/audio.py:344-346)
        with pytest.raises(ValueError):
            A.parse_wav(data)
        with pytest.raises(ValueError):
            audio.read_wav_info(io.BytesIO(data), len(data))
        return
    ref = A.parse_wav(data)
    info = audio.read_wav_info(io.BytesIO(data), len(data))
    assert (info.fmt, info.channels, info.sample_rate, info.data_offset, info.n_frames) == ref
    assert ref[0] == kind and ref[1] == ch and ref[2] == sr and ref[4] <= frames

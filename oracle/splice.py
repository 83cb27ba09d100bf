"""Placeholder expansion + embedding splice oracle (numpy, integer / bit-exact).  TEST INFRASTRUCTURE ONLY.

Follows ref: desta/models/modeling_desta25.py:99-123 (_prepare_audio_context_and_start_positions),
:1645-1651 and desta/trainer/data/simple_dataset.py:228-236 (left-pad offset),
:940-1050 (_prepare_inputs_for_llm: gathers :975-982,1009, scatter loop :1014-1045)."""
import numpy as np


def prepare_audio_context_and_start_positions(token_list, audio_locator, audio_size_list,
                                              transcription_size_list, placeholder_token):
    """Consumes (pop(0)) both size lists, exactly like the reference (:99-123)."""
    assert len(audio_size_list) == len(transcription_size_list)
    result, starts = [], []
    for tok in token_list:
        if tok == audio_locator:
            t = transcription_size_list.pop(0)
            a = audio_size_list.pop(0)
            starts.append(len(result))
            result.extend([placeholder_token] * a)
            result.extend([placeholder_token] * t)
        else:
            result.append(tok)
    return result, starts


def left_pad_offset(starts, attention_mask_row):
    """start += total_length - attention_mask.sum()  (:1645-1651)."""
    pad = len(attention_mask_row) - int(np.sum(attention_mask_row))
    return [s + pad for s in starts]


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 bit pattern (uint16); NaN kept quiet.  This is the implicit
    cast of the slice assignment at :1040-1041."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    nan = np.isnan(x)
    if nan.any():
        rounded = np.where(nan, ((u >> 16) | 0x40).astype(np.uint16), rounded)
    return rounded


def splice(input_ids, table_bits, audio_feats_f32, transcription_ids, start_positions, prompt_size=None):
    """input_ids [B,S] int64; table_bits [V,d] uint16 (bf16 bit patterns); audio_feats_f32 [N,Q,d] float32;
    transcription_ids: list of N 1-D int arrays; start_positions: list of N (b, start).
    -> inputs_embeds bit patterns [B,S,d] uint16.  Later clips overwrite earlier ones (loop order)."""
    input_ids = np.asarray(input_ids)
    out = table_bits[input_ids]                       # embed_tokens(input_ids) (:1009)
    N = len(start_positions)
    assert N == len(transcription_ids) == audio_feats_f32.shape[0]   # (:1005)
    Q = audio_feats_f32.shape[1] if prompt_size is None else prompt_size
    for i in range(N):
        b, s = int(start_positions[i][0]), int(start_positions[i][1])
        t = np.asarray(transcription_ids[i]).reshape(-1)
        rows = np.concatenate([f32_to_bf16_bits(audio_feats_f32[i, :Q]), table_bits[t]], axis=0)
        out[b, s: s + rows.shape[0]] = rows           # raises on length mismatch like torch does
    return out

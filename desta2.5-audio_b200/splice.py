"""Placeholder expansion (host integers) and the `<|AUDIO|>` splice (boundary b3).

ref: desta/models/modeling_desta25.py:99-123 (_prepare_audio_context_and_start_positions), :1645-1651 (left-pad offset),
:940-1050 (_prepare_inputs_for_llm).  The splice itself is one launch of desta_splice (libdesta_b200.so)."""
from typing import List, Sequence

import torch

from . import ops


def prepare_audio_context_and_start_positions(token_list, audio_locator, audio_size_list, transcription_size_list,
                                              placeholder_token):
    """Same semantics as the reference, including consuming (pop(0)) the two size lists."""
    assert len(audio_size_list) == len(transcription_size_list), (
        f"audio_size_list and transcription_size_list must have the same length, audio_size_list: {audio_size_list}, "
        f"transcription_size_list: {transcription_size_list}")
    result, start_positions = [], []
    for x in token_list:
        if x == audio_locator:
            transcription_size = transcription_size_list.pop(0)
            audio_size = audio_size_list.pop(0)
            start_positions.append(len(result))
            result.extend([placeholder_token] * (audio_size + transcription_size))
        else:
            result.append(x)
    return result, start_positions


def apply_left_pad_offset(start_positions: Sequence[int], attention_mask_row) -> List[int]:
    total = len(attention_mask_row)
    pad = total - int(sum(int(v) for v in attention_mask_row))
    return [s + pad for s in start_positions]


class _SpliceFn(torch.autograd.Function):
    """inputs_embeds = splice(table[input_ids], audio_features); gradient flows to the audio rows only
    (the embedding table is frozen in the reference, :1439-1449, and transcript rows come from it under no_grad :975)."""

    @staticmethod
    def forward(ctx, feats, input_ids, table, descs, trans_ids, Q):
        ctx.descs, ctx.Q, ctx.shape, ctx.dtype = descs, Q, feats.shape, feats.dtype
        return ops.splice(input_ids, table, descs, trans_ids, feats.contiguous(), Q)

    @staticmethod
    def backward(ctx, grad_out):
        g = torch.empty(ctx.shape, dtype=torch.float32, device=grad_out.device)
        ops.splice_bwd(grad_out.to(torch.bfloat16), list(ctx.descs), ctx.Q, g)
        live = _last_writer(ctx.descs, ctx.Q)
        if not all(live):
            for i, ok in enumerate(live):
                if not ok:
                    g[i].zero_()
        return g.to(ctx.dtype), None, None, None, None, None


def _last_writer(descs, Q):
    """A clip's audio rows receive gradient only where no later clip overwrote them; overlapping spans never come out
    of prepare_audio_context_and_start_positions, so whole-clip granularity is enough (overlap -> conservative False)."""
    live = [True] * len(descs)
    for i, (b, s, t, _o) in enumerate(descs):
        for (b2, s2, t2, _o2) in descs[i + 1:]:
            if b2 == b and s2 < s + Q and s < s2 + Q + t2:
                live[i] = False
    return live


def prepare_inputs_for_llm(embed_tokens_weight: torch.Tensor, perception, input_ids, attention_mask, batch_features,
                           batch_transcription_ids, batch_start_positions):
    """Drop-in body of DeSTA25AudioModel._prepare_inputs_for_llm for connector_mode == "qformer_1"."""
    table = embed_tokens_weight
    device = table.device
    N_audio = len(batch_start_positions)
    input_ids = input_ids.to(device)
    if N_audio == 0:                                                  # :962-967
        return ops.splice(input_ids, table, [], None, None, perception.dims.prompt_size)
    feats, lengths = perception(input_features=batch_features)
    assert len(batch_start_positions) == len(batch_transcription_ids) == feats.size(0) == len(lengths), (
        "batch_start_positions, batch_transcription_ids, audio_features, speech_feature_lengths must have the same length.")
    Q = lengths[0]
    descs, flat, off = [], [], 0
    for i in range(N_audio):
        b, s = batch_start_positions[i]
        t = batch_transcription_ids[i].reshape(-1)
        descs.append((int(b), int(s), int(t.numel()), off))
        off += int(t.numel())
        flat.append(t)
    trans_ids = torch.cat(flat).to(device=device, dtype=torch.int64) if off > 0 else None
    if feats.requires_grad:
        return _SpliceFn.apply(feats, input_ids, table, tuple(descs), trans_ids, Q)
    return ops.splice(input_ids, table, descs, trans_ids, feats.contiguous(), Q)

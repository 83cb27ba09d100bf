"""Train-mode dropout of the Q-Former (BertConfig hidden_dropout_prob = attention_probs_dropout_prob = 0.1; reference
sites: $TF/models/bert/modeling_bert.py:198-201, 279, 296, 354).  The kernels regenerate a counter-based Philox mask
from (seed, site, element); parity is checked by materialising exactly that mask through the C-ABI test hook and
feeding it to the oracle restatement (whose site placement is pinned against the reference by
tests/test_oracle_cpu.py::test_train_mode_dropout_sites_vs_golden).  Tolerances as in the eval-mode tests."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from desta_b200 import ops

from test_path_gpu import KINDS, _cos, _rel, build_perception  # noqa: E402


def test_mask_rate_determinism_and_sites():
    d = ops.make_dropout(0.1, 0x1234567890ABCDEF, 3)
    m = ops.dropout_mask(d, 4096, 1280, "cuda")
    rate = m.float().mean().item()
    assert abs(rate - 0.9) < 2e-3, rate                       # 5.2 M Bernoulli(0.9) draws: sigma = 1.3e-4
    assert torch.equal(m, ops.dropout_mask(d, 4096, 1280, "cuda"))
    m2 = ops.dropout_mask(ops.make_dropout(0.1, 0x1234567890ABCDEF, 4), 4096, 1280, "cuda")
    m3 = ops.dropout_mask(ops.make_dropout(0.1, 0x1234567890ABCDEE, 3), 4096, 1280, "cuda")
    for other in (m2, m3):                                      # another site / seed: independent mask
        agree = (m == other).float().mean().item()
        assert abs(agree - 0.82) < 5e-3, agree                  # 0.9^2 + 0.1^2
    # no visible structure along rows / columns / inside the groups of 8 that share a Philox call
    assert (m.float().mean(0) - 0.9).abs().max().item() < 0.03
    assert (m.float().mean(1) - 0.9).abs().max().item() < 0.05
    assert (m.view(4096, 160, 8).float().mean((0, 1)) - 0.9).abs().max().item() < 3e-3
    # ragged column count: rows are padded to whole groups internally
    mr = ops.dropout_mask(ops.make_dropout(0.25, 7, 0), 100, 1500, "cuda")
    assert mr.shape == (100, 1500) and abs(mr.float().mean().item() - 0.75) < 6e-3
    assert ops.make_dropout(0.0, 1, 1) is None


@pytest.mark.parametrize("M,N", [(1000, 1280), (64, 384), (4096, 3072)])
def test_dropout_add(M, N):
    x = torch.randn(M, N, device="cuda")
    r = torch.randn(M, N, device="cuda")
    d = ops.make_dropout(0.1, 99, 17)
    keep = ops.dropout_mask(d, M, N, "cuda")
    want = torch.where(keep, x * d.scale, torch.zeros_like(x))
    out = torch.empty_like(x)
    out_bf = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.dropout_add(x, None, d, out_f32=out, out_bf16=out_bf)
    assert torch.equal(out, want) and torch.equal(out_bf, want.to(torch.bfloat16))
    ops.dropout_add(x, r, d, out_f32=out)
    assert (out - (want + r)).abs().max().item() <= 1e-6 * (want + r).abs().max().item()
    x0 = x.clone()
    ops.dropout_add(x, r, None, out_f32=x)                     # p = 0, in place: plain residual add
    assert torch.equal(x, x0 + r)


@pytest.mark.parametrize("B,H,Tq,Tk,scale", [(2, 2, 64, 64, 0.125), (3, 2, 64, 1500, 0.125), (2, 6, 64, 300, 0.125),
                                             (2, 20, 64, 1500, 0.125), (1, 3, 100, 260, 0.125)])
def test_attention_dropout_fwd_bwd(B, H, Tq, Tk, scale):
    g = torch.Generator(device="cuda").manual_seed(Tk)
    d = H * 64
    q = torch.randn(B * Tq, d, device="cuda", generator=g).to(torch.bfloat16)
    k = torch.randn(B * Tk, d, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn(B * Tk, d, device="cuda", generator=g).to(torch.bfloat16)
    do = torch.randn(B * Tq, d, device="cuda", generator=g).to(torch.bfloat16)
    drop = ops.make_dropout(0.1, 4242 + Tk, 9)
    keep = ops.dropout_mask(drop, B * H * Tq, Tk, "cuda").view(B, H, Tq, Tk)
    o = torch.empty_like(q)
    lse = torch.empty(B * H * Tq, device="cuda")
    ops.attention(q, k, v, o, B, H, Tq, Tk, scale, lse=lse, drop=drop)
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    ops.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, B, H, Tq, Tk, scale, drop=drop)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    qh = qf.reshape(B, Tq, H, 64).transpose(1, 2)
    kh = kf.reshape(B, Tk, H, 64).transpose(1, 2)
    vh = vf.reshape(B, Tk, H, 64).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    a = torch.softmax(s, -1) * keep.float() * drop.scale
    ref = (a @ vh).transpose(1, 2).reshape(B * Tq, d)
    assert _rel(o, ref) < 2e-2 and _cos(o, ref) > 0.999, (_rel(o, ref), _cos(o, ref))
    # the result differs from the undropped attention (the mask really is applied)
    o0 = torch.empty_like(q)
    ops.attention(q, k, v, o0, B, H, Tq, Tk, scale)
    assert _rel(o0, ref) > 5e-2
    ref.backward(do.float())
    for name, x, y in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        assert _rel(x, y) < 3e-2 and _cos(x, y) > 0.999, f"{name}: rel {_rel(x, y)} cos {_cos(x, y)}"
    # row_keep_ratio: per query row, the sum of the dropped, rescaled probabilities over their plain sum (same mask, same output)
    rr = torch.full((B * H * Tq,), -1.0, device="cuda")
    o2 = torch.empty_like(q)
    ops.attention(q, k, v, o2, B, H, Tq, Tk, scale, drop=ops.make_dropout(0.1, 4242 + Tk, 9, rr))
    assert torch.equal(o2, o)
    ratio_ref = a.detach().sum(-1)                            # [B, H, Tq]: softmax rows sum to 1
    assert (rr.view(B, H, Tq) - ratio_ref).abs().max().item() < 2e-3
    # ... which turns the value-bias gradient into a reduction over dO:  sum_k dV[k] = sum_q ratio_q dO_q  per head
    got = torch.zeros(d, device="cuda")
    ops.colsum_heads(do, rr, B, H, Tq, got)
    want = vf.grad.reshape(B, Tk, d).sum((0, 1))
    assert _rel(got, want) < 2e-2 and _cos(got, want) > 0.9995, (_rel(got, want), _cos(got, want))
    plain = torch.zeros(d, device="cuda")
    ops.colsum_heads(do, None, B, H, Tq, plain)               # no weights: a plain column sum of dO
    assert _rel(plain, do.float().sum(0)) < 1e-3


@pytest.mark.parametrize("name,n_clips", [("micro", 3), ("debug", 2), ("l3_slim", 1), ("l3_slim", 2)])     # l3_slim: flagship layer shapes
def test_connector_train_mode_vs_oracle_with_kernel_masks(name, n_clips):
    from desta_b200.autograd import DropPlan
    from oracle import encoder as o_enc, logmel as o_mel, qformer as o_qf
    from oracle.configs import CONFIGS
    from oracle.synth import synth_waves
    from oracle.weights import make_weights
    cfg = CONFIGS[name]
    W = make_weights(cfg, seed=11, with_table=False)
    waves = synth_waves(seed=7, kinds=KINDS)[:n_clips]
    mel = torch.from_numpy(o_mel.logmel(waves, cfg.n_mels))
    gout = torch.randn(n_clips, cfg.prompt_size, cfg.d_llm, generator=torch.Generator().manual_seed(5))
    p = build_perception(cfg, W).train()
    for prm in p.whisper.parameters():
        prm.requires_grad_(False)
    wave = torch.from_numpy(np.stack([o_mel.pad_or_truncate(w) for w in waves])).cuda()
    torch.manual_seed(77)
    feats, _ = p(input_features=wave)
    plan = p.last_drop_plan
    assert isinstance(plan, DropPlan) and plan.p_hidden == pytest.approx(0.1) and plan.p_attn == pytest.approx(0.1)
    (feats * gout.cuda()).sum().backward()
    torch.manual_seed(77)                                   # same torch seed -> same step seed -> same output
    feats_again, _ = p(input_features=wave)
    assert torch.equal(feats_again, feats)
    feats_other, _ = p(input_features=wave)                  # next step draws a fresh seed
    assert not torch.equal(feats_other, feats)
    with torch.no_grad():                                    # eval mode: nn.Dropout semantics, nothing dropped
        feats_eval, _ = p.eval()(input_features=wave)
    p.train()

    # the masks the kernels used, per (tap, layer, site), in the oracle's tensor layouts
    K, N, Q, H, T, d = len(cfg.tap_layers), n_clips, cfg.prompt_size, cfg.n_heads, cfg.max_pos, cfg.d_model
    masks, scales = {}, {}
    for l in range(cfg.qf_layers):
        for which in range(5):
            dr = plan.site(l, which)
            scales[which] = dr.scale
            if which == o_qf.SELF_PROBS:
                m = ops.dropout_mask(dr, K * N * H * Q, Q, "cuda").view(K, N, H, Q, Q)
            elif which == o_qf.CROSS_PROBS:
                m = ops.dropout_mask(dr, K * N * H * Q, T, "cuda").view(K, N, H, Q, T)
            else:
                m = ops.dropout_mask(dr, K * N * Q, d, "cuda").view(K, N, Q, d)
            for k in range(K):
                masks[(k, l, which)] = m[k].cpu()
    with torch.no_grad():
        taps = o_enc.encoder_taps(W, cfg, mel)
    Wg = {kk: (v.clone().requires_grad_(True) if kk.startswith("perception.connector.") else v) for kk, v in W.items()}
    out_ref = o_qf.connector(Wg, cfg, taps, drop=o_qf.mask_dropout(masks, scales))
    (out_ref * gout).sum().backward()
    rel, cos = _rel(feats, out_ref), _cos(feats, out_ref)
    assert rel <= 2e-2 and cos >= 0.999, f"train-mode connector out: rel {rel:.4f} cos {cos:.6f}"
    assert _rel(feats_eval, out_ref) > 2e-2                   # and it is not the eval-mode result
    gmax = max(Wg["perception.connector." + pn].grad.abs().max().item() for pn, _ in p.connector.named_parameters())
    bad = []
    for pname, prm in p.connector.named_parameters():
        ref = Wg["perception.connector." + pname].grad
        if ref.abs().max().item() < 1e-3 * gmax:
            # ill-conditioned tensor: a flat softmax (pure-noise clip) makes dS = P o (dP - D) cancel almost exactly, so the
            # query / key gradients of that attention are ~1e-4 of the largest gradient and dominated by the bf16 rounding
            # of O inside D = rowsum(dO o O) - the same cancellation limits the reference's own bf16-autocast backward.
            # Such tensors are held to an absolute bound on the scale of the whole gradient instead of a relative one.
            if (prm.grad.float().cpu() - ref).abs().max().item() > 2e-3 * gmax:
                bad.append(f"{pname}: |delta| {(prm.grad.float().cpu() - ref).abs().max().item():.3e} vs gmax {gmax:.3e}")
            continue
        r_, c_ = _rel(prm.grad, ref), _cos(prm.grad, ref)
        if not (r_ <= 5e-2 and c_ >= 0.995):
            bad.append(f"{pname}: rel {r_:.4f} cos {c_:.5f}")
    assert not bad, "\n".join(bad)

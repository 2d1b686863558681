"""GPU tests of the training-step runtime around the kernels: multi-tensor weight preparation, B200FusedAdamW (+Noam) against
torch.optim.AdamW driven by the restated NoamAnnealing rate, the device-side RNG counter and CapturedStep (CUDA graph)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from nemo_b200 import _lib as L
    from nemo_b200.encoder import B200ConformerEncoder
    from nemo_b200.graph import CapturedStep, seed_counter
    from nemo_b200.optim import B200FusedAdamW, noam_lr
    from nemo_b200.parallel import FlatGradients
    from oracle import conformer_oracle as O
    DEV = torch.device("cuda:0")

CFG = dict(n_layers=2, d_model=128, n_heads=2, conv_kernel_size=9, subsampling_kind="striding", subsampling_factor=4)


def _enc(cfg=CFG, p=0.0, seed=5):
    sd = O.init_state_dict(seed=seed, **cfg)
    enc = B200ConformerEncoder(feat_in=80, n_layers=cfg["n_layers"], d_model=cfg["d_model"], n_heads=cfg["n_heads"],
                               conv_kernel_size=cfg["conv_kernel_size"], subsampling=cfg["subsampling_kind"],
                               subsampling_factor=cfg["subsampling_factor"],
                               subsampling_conv_channels=cfg.get("subsampling_conv_channels", -1), dropout=p,
                               dropout_pre_encoder=p, dropout_emb=0.0, dropout_att=p)
    enc.load_state_dict(sd)
    return enc.to(DEV).train()


@pytest.mark.parametrize("kind", ["striding", "dw_striding"])
def test_prep_weights_layouts(kind):
    cfg = dict(CFG) if kind == "striding" else dict(n_layers=1, d_model=128, n_heads=2, conv_kernel_size=9, subsampling_kind="dw_striding",
                                                    subsampling_factor=8, subsampling_conv_channels=64)
    enc = _enc(cfg)
    W = enc._weights()
    torch.cuda.synchronize()
    pe, d = enc.pre_encode, enc.d_model
    C, F2 = pe._conv_channels, pe._f_out
    bf = lambda t: t.detach().to(torch.bfloat16)
    w16 = torch.zeros(C, 16, device=DEV)
    w16[:, :9] = pe.conv[0].weight.detach().reshape(C, 9)
    assert torch.equal(W["conv1"], bf(w16))
    if kind == "striding":
        w2 = pe.conv[2].weight.detach().reshape(C, C, 9)
        assert torch.equal(W["conv2"], bf(w2.permute(0, 2, 1).reshape(C, 9 * C)))      # [co, tap, ci]
        assert torch.equal(W["conv2_d"], bf(w2.permute(1, 2, 0).reshape(C, 9 * C)))    # [ci, tap, co]
    else:
        assert torch.equal(W["pw3"], bf(pe.conv[3].weight.reshape(C, C)))
        assert torch.equal(W["pw6"], bf(pe.conv[6].weight.reshape(C, C)))
    assert torch.equal(W["out"], bf(pe.out.weight.detach().view(d, C, F2).permute(0, 2, 1).reshape(d, F2 * C)))
    lyr = enc.layers[0]
    at, cv = lyr.self_attn, lyr.conv
    assert torch.equal(W["0.qkv"], bf(torch.cat([at.linear_q.weight, at.linear_k.weight, at.linear_v.weight], 0)))
    assert torch.equal(W["0.qkv_b"], torch.cat([at.linear_q.bias, at.linear_k.bias, at.linear_v.bias], 0).detach())
    assert torch.equal(W["0.ff1.w1"], bf(lyr.feed_forward1.linear1.weight)) and torch.equal(W["0.ff2.w2"], bf(lyr.feed_forward2.linear2.weight))
    assert torch.equal(W["0.pos"], bf(at.linear_pos.weight)) and torch.equal(W["0.o"], bf(at.linear_out.weight))
    w1 = cv.pointwise_conv1.weight.detach().view(2, d, d)
    assert torch.equal(W["0.pw1"], bf(w1.permute(1, 0, 2).reshape(2 * d, d)))           # rows (2j, 2j+1) = (value_j, gate_j)
    assert torch.equal(W["0.pw1_b"], cv.pointwise_conv1.bias.detach().view(2, d).t().reshape(2 * d))
    assert torch.equal(W["0.pw2"], bf(cv.pointwise_conv2.weight.view(d, d)))
    # weights edited through .data (no version bump): training refills every forward, eval needs the explicit invalidate
    with torch.no_grad():
        at.linear_pos.weight.data.mul_(2.0)
    assert torch.equal(enc._weights()["0.pos"], bf(at.linear_pos.weight))
    enc.eval()
    enc._weights()
    with torch.no_grad():
        at.linear_pos.weight.data.mul_(0.5)
    assert not torch.equal(enc._weights()["0.pos"], bf(at.linear_pos.weight))  # documented: stale until invalidated
    enc.invalidate_weight_cache()
    assert torch.equal(enc._weights()["0.pos"], bf(at.linear_pos.weight))


def test_fused_adamw_noam_matches_torch_adamw():
    """5 steps of B200FusedAdamW (device-side Noam rate) vs torch.optim.AdamW whose lr is set each step to the restated
    NoamAnnealing value (lr_scheduler.py:555-575), on identical gradients."""
    torch.manual_seed(0)
    mod = torch.nn.ModuleDict({"a": torch.nn.Linear(96, 130), "b": torch.nn.Linear(130, 7), "n": torch.nn.LayerNorm(33)}).to(DEV)
    ref = torch.nn.ModuleDict({"a": torch.nn.Linear(96, 130), "b": torch.nn.Linear(130, 7), "n": torch.nn.LayerNorm(33)}).to(DEV)
    ref.load_state_dict(mod.state_dict())
    fg = FlatGradients(mod)
    d_model, warm, min_lr, base = 512, 3, 1e-6, 2.0
    opt = B200FusedAdamW(fg, lr=base, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-3, noam_d_model=d_model, noam_warmup_steps=warm,
                         noam_min_lr=min_lr)
    topt = torch.optim.AdamW(ref.parameters(), lr=base, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-3)
    g = torch.Generator(device="cuda").manual_seed(1)
    for k in range(1, 6):
        fg.zero()
        for (n, p), (_, q) in zip(mod.named_parameters(), ref.named_parameters()):
            gr = torch.randn(p.shape, generator=g, device=DEV) * 0.1
            p.grad.copy_(gr)
            q.grad = gr.clone()
        lr = noam_lr(base, d_model, warm, min_lr, k)
        for grp in topt.param_groups:
            grp["lr"] = lr
        opt.step()
        topt.step()
        assert abs(opt.last_lr - lr) <= 1e-6 * lr, (k, opt.last_lr, lr)
        for (n, p), (_, q) in zip(mod.named_parameters(), ref.named_parameters()):
            err = (p - q).abs().max().item()
            assert err <= 2e-6 * max(1.0, q.abs().max().item()), (k, n, err)
    assert float(opt.state.item()) == 5.0
    # Noam shape: linear warm-up to the peak at step = warmup, then step^-0.5
    assert noam_lr(2.0, 512, 10000, 1e-6, 10001) == pytest.approx(2.0 * 512 ** -0.5 * 10000 ** -0.5)
    assert noam_lr(2.0, 512, 10000, 1e-6, 2) == pytest.approx(2.0 * 512 ** -0.5 * 1 * 10000 ** -1.5)


def _mini_step(enc, fg, opt, x, ln, r):
    fg.zero()
    y, _ = enc(audio_signal=x, length=ln)
    loss = (y.float() * r).sum() / y.shape[0]
    loss.backward()
    fg.finish()
    opt.step()
    return loss.detach()


def test_seed_counter_changes_masks_and_captured_step_equals_eager():
    """(1) the device counter alone (same by-value seed) changes the dropout masks; (2) a captured training step replayed k
    times lands on the same weights as k eager steps with the same counter values; (3) replays draw fresh masks."""
    ctr = seed_counter()
    x = torch.randn(2, 80, 160, generator=torch.Generator().manual_seed(12)).to(DEV)
    ln = torch.tensor([160, 111], device=DEV)

    def make():
        enc = _enc(p=0.1)
        enc._seed = lambda: 1234  # freeze the by-value seed exactly as a captured graph does
        fg = FlatGradients(enc)
        # eps = 1: the update is ~lr * m (SGD-like).  With Adam's default eps a parameter whose gradient is rounding noise
        # (analytically zero: depthwise_conv.bias under BatchNorm) moves by +-lr with a random sign, in eager and captured alike
        opt = B200FusedAdamW(fg, lr=1e-2, betas=(0.9, 0.98), eps=1.0, weight_decay=1e-3)
        return enc, fg, opt

    enc, fg, opt = make()
    ctr.fill_(100)
    with torch.no_grad():
        y1, _ = enc(audio_signal=x, length=ln)
        y1b, _ = enc(audio_signal=x, length=ln)
        ctr.fill_(101)
        y2, _ = enc(audio_signal=x, length=ln)
    assert torch.equal(y1, y1b) and not torch.equal(y1, y2)

    r = torch.randn(y1.shape, generator=torch.Generator().manual_seed(13)).to(DEV)
    # eager: 3 steps at counter values 201, 202, 203
    enc_e, fg_e, opt_e = make()
    losses_e = []
    for k in range(3):
        ctr.fill_(201 + k)
        losses_e.append(_mini_step(enc_e, fg_e, opt_e, x, ln, r).item())
    # captured: warm-up runs mutate weights / optimizer state, so snapshot and restore around the capture
    enc_g, fg_g, opt_g = make()
    snap = {k: v.clone() for k, v in enc_g.state_dict().items()}
    step = CapturedStep(lambda: _mini_step(enc_g, fg_g, opt_g, x, ln, r), warmup=2)
    assert step.launches_per_replay > 50
    enc_g.load_state_dict(snap)
    opt_g.exp_avg.zero_(); opt_g.exp_avg_sq.zero_(); opt_g.state.zero_()
    ctr.fill_(200)  # the graph's first node advances it to 201
    losses_g = [step().item() for _ in range(3)]
    torch.cuda.synchronize()
    assert int(ctr.item()) == 203 and float(opt_g.state.item()) == 3.0
    assert len(set(losses_g)) == 3  # fresh dropout masks every replay
    for a, b in zip(losses_e, losses_g):
        assert abs(a - b) <= 2e-3 * max(1e-6, abs(a)), (losses_e, losses_g)
    worst = 0.0
    for (n, p), (_, q) in zip(enc_e.named_parameters(), enc_g.named_parameters()):
        worst = max(worst, ((p - q).norm() / p.norm().clamp_min(1e-12)).item())
    print("captured vs eager weights after 3 steps: worst rel-L2", worst)
    # split-K atomics reorder fp32 sums (a last-bit difference that can flip a bf16 rounding downstream); everything else
    # is identical.  Observed 7e-6 ... 1e-3 depending on the data; a wrong mask / stale weight would be O(1).
    assert worst < 5e-3
    ctr.zero_()


@pytest.mark.parametrize("reduction", ["mean_batch", "mean_volume", "none"])
def test_ctc_loss_vs_oracle(reduction):
    """B200CTCLoss (csrc/ctc.cu) against the oracle restatement of nemo/collections/asr/losses/ctc.py (torch CTC on CPU, fp32):
    ragged inputs, repeated labels, an empty target and a BPE-sized vocabulary."""
    from nemo_b200.losses import B200CTCLoss

    g = torch.Generator().manual_seed(9)
    B, T, V = 5, 188, 1024
    logits = torch.randn(B, T, V + 1, generator=g)
    in_len = torch.tensor([188, 150, 97, 40, 188])
    tgt_len = torch.tensor([60, 33, 20, 0, 7])
    tgt = torch.randint(0, V, (B, 60), generator=g)
    tgt[0, 10:14] = tgt[0, 10]  # repeated labels
    x = logits.to(DEV).requires_grad_(True)
    lp = torch.log_softmax(x, dim=-1)
    loss = B200CTCLoss(num_classes=V, reduction=reduction)(log_probs=lp, targets=tgt.to(DEV), input_lengths=in_len.to(DEV),
                                                           target_lengths=tgt_len.to(DEV))
    xr = logits.clone().requires_grad_(True)
    ref = O.ctc_loss(torch.log_softmax(xr, dim=-1), tgt, in_len, tgt_len, blank=V, reduction=reduction)
    w = torch.randn(ref.shape, generator=g) if reduction == "none" else torch.ones(())
    (loss * w.to(DEV)).sum().backward()
    (ref * w).sum().backward()
    e = ((loss.cpu() - ref).abs() / ref.abs().clamp_min(1e-6)).max().item()
    eg = ((x.grad.cpu() - xr.grad).norm() / xr.grad.norm()).item()
    print(f"[ctc {reduction}] loss rel err {e:.2e}; dlogits rel-L2 {eg:.2e}")
    assert e < 1e-5 and eg < 1e-4


def test_ctc_loss_golden_reference_fixture(golden_dir):
    """fixture produced by the unmodified reference ConvASRDecoder + CTCLoss arithmetic (tools/make_golden.py)"""
    import os

    import numpy as np

    from nemo_b200.losses import B200CTCLoss

    z = np.load(os.path.join(golden_dir, "ctc_head.npz"))
    lp = torch.from_numpy(z["logp"]).to(DEV).requires_grad_(True)
    crit = B200CTCLoss(num_classes=30, reduction="mean_batch")
    loss = crit(log_probs=lp, targets=torch.from_numpy(z["tgt"]).to(DEV), input_lengths=torch.from_numpy(z["in_len"]).to(DEV),
                target_lengths=torch.from_numpy(z["tgt_len"]).to(DEV))
    assert abs(loss.item() - float(z["loss"])) < 1e-4 * abs(float(z["loss"]))
    # gradient w.r.t. the log-probs vs torch's (the reference's) on the same log-probs
    lr = torch.from_numpy(z["logp"]).requires_grad_(True)
    O.ctc_loss(lr, torch.from_numpy(z["tgt"]), torch.from_numpy(z["in_len"]), torch.from_numpy(z["tgt_len"]), blank=30).backward()
    loss.backward()
    assert (lp.grad.cpu() - lr.grad).abs().max().item() < 1e-5
    # infeasible utterance (target longer than the input): inf, or 0 with zero_infinity and a zero gradient
    lp2 = torch.log_softmax(torch.randn(1, 3, 31, device=DEV), -1).requires_grad_(True)
    t2, il2, tl2 = torch.randint(0, 30, (1, 8), device=DEV), torch.tensor([3], device=DEV), torch.tensor([8], device=DEV)
    assert torch.isinf(B200CTCLoss(30, reduction="none")(log_probs=lp2, targets=t2, input_lengths=il2, target_lengths=tl2)).all()
    l0 = B200CTCLoss(30, zero_infinity=True, reduction="none")(log_probs=lp2, targets=t2, input_lengths=il2, target_lengths=tl2)
    l0.sum().backward()
    assert l0.item() == 0.0 and lp2.grad.abs().max().item() == 0.0

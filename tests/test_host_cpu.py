"""CPU tests of the host-side logic: drop-in boundary (constructor kwargs, state_dict layout, error behaviour),
data-parallel gradient plumbing (gloo, world_size 2) and the benchmark's work model.  No CUDA compute."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nemo_b200.encoder import B200ConformerEncoder
from nemo_b200.parallel import FlatGradients, shard_batch
from nemo_b200.preprocessor import B200MelPreprocessor, slaney_mel_filterbank
from oracle import conformer_oracle as O
from oracle import ref_loader as rl

needs_ref = pytest.mark.skipif(not rl.available(), reason="/root/reference not present")


def test_encoder_state_dict_layout_matches_oracle_init():
    for kind, fac, ch in (("striding", 4, -1), ("dw_striding", 8, 64)):
        enc = B200ConformerEncoder(feat_in=80, n_layers=2, d_model=128, n_heads=2, conv_kernel_size=9, subsampling=kind,
                                   subsampling_factor=fac, subsampling_conv_channels=ch)
        sd = O.init_state_dict(n_layers=2, d_model=128, n_heads=2, conv_kernel_size=9, subsampling_kind=kind,
                               subsampling_factor=fac, subsampling_conv_channels=ch)
        res = enc.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        for k, v in enc.state_dict().items():
            assert tuple(v.shape) == tuple(sd[k].shape), k


@needs_ref
def test_encoder_is_a_drop_in_for_the_reference_module():
    ref = rl.make_encoder(feat_in=80, n_layers=2, d_model=128, n_heads=2, conv_kernel_size=31, dropout=0.1, dropout_att=0.1)
    new = B200ConformerEncoder.from_reference(ref)
    assert list(new.state_dict().keys()) == list(ref.state_dict().keys())  # same keys, same order
    for (k, a), (_, b) in zip(new.state_dict().items(), ref.state_dict().items()):
        assert torch.equal(a, b), k
    assert (new._feat_out, new.d_model, new.subsampling_factor, len(new.layers)) == (ref._feat_out, ref.d_model, ref.subsampling_factor, len(ref.layers))
    assert new.dropout == 0.1 and new.dropout_att == 0.1 and hasattr(new, "pre_encode")
    # reference error behaviour (conformer_encoder.py:569-573) and no silent CPU fallback
    with pytest.raises(ValueError):
        new(audio_signal=torch.zeros(1, 79, 50), length=torch.tensor([50]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        new(audio_signal=torch.zeros(1, 80, 50), length=torch.tensor([50]))


def test_encoder_rejects_configurations_outside_the_hot_path():
    base = dict(feat_in=80, n_layers=1, d_model=128, n_heads=2)
    for bad in (dict(self_attention_model="abs_pos"), dict(subsampling="vggnet"), dict(conv_norm_type="layer_norm"),
                dict(att_context_size=[64, 0]), dict(n_heads=4), dict(causal_downsampling=True),
                dict(stochastic_depth_drop_prob=0.1), dict(reduction="pooling", reduction_factor=2, reduction_position=0)):
        with pytest.raises(NotImplementedError):
            B200ConformerEncoder(**{**base, **bad})
    enc = B200ConformerEncoder(**base)
    with pytest.raises(NotImplementedError):
        enc.cache_aware_stream_step()
    with pytest.raises(NotImplementedError):
        enc(audio_signal=torch.zeros(1, 80, 10), length=None, cache_last_channel=torch.zeros(1))


def test_preprocessor_boundary(golden_dir):
    pre = B200MelPreprocessor(window_size=0.025, window_stride=0.01, features=80, n_fft=512, pad_to=0)
    assert list(pre.state_dict().keys()) == ["featurizer.window", "featurizer.fb"]
    z = np.load(os.path.join(golden_dir, "mel.npz"))  # buffers of the reference FilterbankFeatures
    assert np.array_equal(pre.featurizer.fb[0].numpy(), z["fb"])
    assert np.array_equal(pre.featurizer.window.numpy(), z["window"])
    assert pre.filter_banks.shape == (1, 80, 257) and pre._sample_rate == 16000
    assert (pre.featurizer.dither, pre.featurizer.pad_to) == (1e-5, 0)
    with pytest.raises(ValueError):
        B200MelPreprocessor(window_size=0.025, n_window_size=400)
    with pytest.raises(NotImplementedError):
        B200MelPreprocessor(window_size=0.025, window_stride=0.01, n_fft=1024)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pre(input_signal=torch.zeros(1, 1600), length=torch.tensor([1600]))
    assert np.array_equal(slaney_mel_filterbank(16000, 512, 80), O.slaney_mel_filterbank())


def test_shard_batch_partitions_the_global_batch():
    for n, w in ((32, 1), (32, 8), (64, 4), (16, 8), (10, 3)):
        spans = [shard_batch(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _dp_worker(rank, world, port, q, bucket_bytes=96 << 20, defer=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    enc = B200ConformerEncoder(feat_in=80, n_layers=2, d_model=128, n_heads=2, conv_kernel_size=9)
    fg = FlatGradients(enc, bucket_bytes=bucket_bytes, defer=defer)
    calls = []
    real = dist.all_reduce

    def counting(t, *a, **k):
        calls.append(t.numel())
        return real(t, *a, **k)

    dist.all_reduce = counting
    assert all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in enc.parameters())
    # emulate the backward orchestration: last layer first, each layer's slice all-reduced as soon as it is "ready"
    g = torch.Generator().manual_seed(100 + rank)
    local = torch.randn(fg.flat.numel(), generator=g)
    for pfx in ("layers.1.", "layers.0.", "pre_encode."):
        lo, hi = fg.span(pfx)
        fg.flat[lo:hi] += local[lo:hi]
        fg.ready(lo, hi)
    fg.finish()
    expect = sum(torch.randn(fg.flat.numel(), generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
    ok = torch.allclose(fg.flat, expect, atol=1e-6)
    covered = sorted(fg.span(p) for p in ("pre_encode.", "layers.0.", "layers.1."))
    contiguous = covered[0][0] == 0 and covered[-1][1] == fg.flat.numel() and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    # every element is reduced exactly once, whatever the bucketing
    q.put((rank, bool(ok), bool(contiguous), fg.loss_scale(), len(calls), sum(calls) == fg.flat.numel()))
    dist.all_reduce = real
    dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes,defer,n_calls", [(96 << 20, False, 1), (1, False, 3), (96 << 20, True, 1)])
def test_flat_gradient_allreduce_gloo_world2(bucket_bytes, defer, n_calls):
    """layer slices finished by the backward are coalesced into buckets of >= bucket_bytes (one call for this small model, one per
    layer with bucket_bytes = 1); defer=True reduces the whole buffer once in finish() (gradient accumulation)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (7 if defer else 0) + (13 if bucket_bytes == 1 else 0)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q, bucket_bytes, defer)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, contiguous, scale, calls, once in res:
        assert ok and contiguous and scale == 0.5 and calls == n_calls and once, (rank, ok, contiguous, scale, calls, once)


def test_bench_work_model_matches_baseline_md():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    assert abs(bench.flops_fwd_bwd(32, 751) / 1e12 - 27.34) < 0.02  # 27.34 TFLOP / step at 32x30 s
    assert abs(bench.flops_fwd_bwd(1, 751) / 6 / 1e9 * 2 - 298.47) < 12  # ~298.5 GFLOP fwd for 1x30 s (pos term is batch-indep.)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port timed on the host cores — the one other place besides tests / smoke
    that may execute oracle/) prints ONE JSON line with the keys the driver reads, without touching a GPU."""
    import json
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "audio-sec/s" and j["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "e2e",
                "cpu_baseline"):
        assert key in j, key
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert j["value"] > 0


@pytest.mark.parametrize("time_width", [10, 0.05])
def test_spec_augment_host_draws_match_the_oracle(time_width):
    """the span boundaries B200SpecAugment hands to its CUDA kernel (host logic, no GPU needed) reproduce the masks of the
    oracle restatement of the reference class, in the device-RNG and in the host-RNG mode"""
    import random

    from nemo_b200.spec_augment import B200SpecAugment

    B, F, T = 3, 80, 301
    x = torch.randn(B, F, T)
    ln = torch.tensor([301, 250, 17])

    def mask_from(t0, t1, f0, f1):
        t, f = torch.arange(T), torch.arange(F)
        mt = ((t[None, None, :] >= t0[:, :, None]) & (t[None, None, :] < t1[:, :, None])).any(1)
        mf = ((f[None, None, :] >= f0[:, :, None]) & (f[None, None, :] < f1[:, :, None])).any(1)
        return mt[:, None, :] | mf[:, :, None]

    sa = B200SpecAugment(freq_masks=2, time_masks=5, freq_width=27, time_width=time_width)
    torch.manual_seed(7)
    t0, t1 = sa._bounds_vectorized(B, T, 5, ln, time_width, sa.TIME_AXIS, x.device)
    f0, f1 = sa._bounds_vectorized(B, F, 2, ln, 27, sa.FREQ_AXIS, x.device)
    torch.manual_seed(7)
    want = O.spec_augment_vectorized(x, ln, 2, 5, 27, time_width) == 0
    assert torch.equal(mask_from(t0, t1, f0, f1), want)

    sa = B200SpecAugment(freq_masks=2, time_masks=5, freq_width=27, time_width=time_width, rng=random.Random(3),
                         use_vectorized_code=False)
    fs, fe, ts, te = sa._bounds_legacy(B, F, ln.numpy())
    want = O.spec_augment_legacy(x, ln, random.Random(3), 2, 5, 27, time_width) == 0
    got = mask_from(torch.tensor(ts), torch.tensor(te), torch.tensor(fs), torch.tensor(fe))
    assert torch.equal(got, want)


def test_nemo_checkpoint_round_trip(tmp_path):
    """.nemo import / export (save_restore_connector.py:44-45,599-607,700-759): a reference-layout checkpoint — written here
    exactly the way `_make_nemo_file_from_folder` does (tar 'w:', arcname '.') from live reference modules when
    /root/reference is present, else from oracle-initialised weights — loads into the B200 modules, and what `save_nemo`
    writes has the reference's member names / key layout and loads back bit-identically (also from a gzip'ed tar)."""
    import tarfile

    import yaml

    from nemo_b200 import checkpoint as ck
    from nemo_b200.decoder import B200ConvASRDecoder
    from nemo_b200.encoder import B200ConformerEncoder
    from nemo_b200.preprocessor import B200MelPreprocessor
    from oracle import conformer_oracle as O
    from oracle import ref_loader as rl

    enc_cfg = dict(feat_in=80, n_layers=2, d_model=128, n_heads=2, conv_kernel_size=9, subsampling="striding", subsampling_factor=4,
                   subsampling_conv_channels=-1, dropout=0.1, dropout_pre_encoder=0.1, dropout_emb=0.0, dropout_att=0.1)
    if rl.available():
        ref = rl.load()
        enc_ref = rl.make_encoder(seed=3, **enc_cfg)
        dec_ref = ref.ConvASRDecoder(feat_in=128, num_classes=64)
        fb_ref = ref.FilterbankFeatures(sample_rate=16000, n_window_size=400, n_window_stride=160, n_fft=512, nfilt=80, dither=1e-5, pad_to=0)
        sd = {f"encoder.{k}": v for k, v in enc_ref.state_dict().items()}
        sd.update({f"decoder.{k}": v for k, v in dec_ref.state_dict().items()})
        sd.update({f"preprocessor.featurizer.{k}": v for k, v in fb_ref.state_dict().items()})
    else:
        sd = {f"encoder.{k}": v for k, v in O.init_state_dict(seed=3, n_layers=2, d_model=128, n_heads=2, conv_kernel_size=9).items()}
        sd["decoder.decoder_layers.0.weight"] = torch.randn(65, 128, 1)
        sd["decoder.decoder_layers.0.bias"] = torch.randn(65)
    conf = {"sample_rate": 16000, "encoder": dict(enc_cfg, _target_="nemo.collections.asr.modules.ConformerEncoder"),
            "decoder": {"_target_": "nemo.collections.asr.modules.ConvASRDecoder", "feat_in": 128, "num_classes": 64}}
    src = tmp_path / "src"
    src.mkdir()
    torch.save(sd, src / "model_weights.ckpt")
    (src / "model_config.yaml").write_text(yaml.safe_dump(conf))
    ref_file = str(tmp_path / "ref.nemo")
    with tarfile.open(ref_file, "w:") as tar:  # save_restore_connector.py:605-606
        tar.add(str(src), arcname=".")
    cfg, got = ck.load_nemo(ref_file)
    assert cfg["encoder"]["d_model"] == 128 and set(got) == set(sd)
    enc = ck.encoder_from_config(cfg["encoder"])
    dec = B200ConvASRDecoder(feat_in=128, num_classes=64)
    pre = B200MelPreprocessor(window_size=0.025, window_stride=0.01, features=80, n_fft=512, pad_to=0)
    ck.restore(ref_file, preprocessor=pre if rl.available() else None, encoder=enc, decoder=dec)
    for k, v in enc.state_dict().items():
        assert torch.equal(v, sd[f"encoder.{k}"]), k
    assert torch.equal(dec.state_dict()["decoder_layers.0.weight"], sd["decoder.decoder_layers.0.weight"])
    ours = ck.save_nemo(str(tmp_path / "ours.nemo"), conf, preprocessor=pre, encoder=enc, decoder=dec)
    with tarfile.open(ours, "r:") as tar:
        names = sorted(m.name for m in tar.getmembers())
    with tarfile.open(ref_file, "r:") as tar:
        ref_names = sorted(m.name for m in tar.getmembers())
    assert names == ref_names == [".", "./model_config.yaml", "./model_weights.ckpt"]
    cfg2, sd2 = ck.load_nemo(ours)
    assert cfg2 == cfg
    assert [k for k in sd2 if k.startswith("encoder.")] == [k for k in sd if k.startswith("encoder.")]  # key ORDER too
    for k in sd:
        if k.startswith(("encoder.", "decoder.")):
            assert torch.equal(sd2[k], sd[k]), k
    if rl.available():  # what we wrote loads into the unmodified reference modules, strictly
        parts = ck.split_state_dict(sd2)
        enc_ref.load_state_dict(parts["encoder"], strict=True)
        dec_ref.load_state_dict(parts["decoder"], strict=True)
    gz = str(tmp_path / "old.nemo")
    with tarfile.open(gz, "w:gz") as tar:  # pre-1.7 checkpoints are gzip'ed
        tar.add(str(src), arcname=".")
    assert set(ck.load_nemo(gz)[1]) == set(sd)
    evil = str(tmp_path / "evil.nemo")
    with tarfile.open(evil, "w:") as tar:
        tar.add(str(src / "model_weights.ckpt"), arcname="../model_weights.ckpt")
    with pytest.raises(FileNotFoundError):
        ck.load_nemo(evil)


def test_grad_layout_keeps_qkv_adjacent_and_block_views():
    """parallel.grad_layout: inside every attention block linear_q/k/v weights (and biases) are adjacent in the flat gradient
    buffer (the fused q|k|v weight gradient is one [3d, d] GEMM output written in place); a frozen member breaks the block."""
    from nemo_b200.encoder import B200ConformerEncoder
    from nemo_b200.parallel import FlatGradients, grad_layout

    enc = B200ConformerEncoder(feat_in=80, n_layers=2, d_model=128, n_heads=2, conv_kernel_size=9)
    names = [n for n, _ in grad_layout(enc.named_parameters())]
    assert sorted(names) == sorted(n for n, _ in enc.named_parameters())
    i = names.index("layers.1.self_attn.linear_q.weight")
    assert names[i:i + 6] == [f"layers.1.self_attn.linear_{x}.{y}" for y in ("weight", "bias") for x in ("q", "k", "v")]
    fg = FlatGradients(enc)
    blk = fg.block("layers.0.self_attn.linear_q.weight", 3)
    assert blk is not None and blk.numel() == 3 * 128 * 128
    assert blk.data_ptr() == enc.layers[0].self_attn.linear_q.weight.grad.data_ptr()
    assert enc.layers[0].self_attn.linear_v.weight.grad.data_ptr() == blk.data_ptr() + 2 * 128 * 128 * 4
    lo, hi = fg.span("layers.1.")
    assert hi - lo == sum(p.numel() for p in enc.layers[1].parameters())  # a layer stays one contiguous all-reduce span
    enc2 = B200ConformerEncoder(feat_in=80, n_layers=1, d_model=128, n_heads=2, conv_kernel_size=9)
    enc2.layers[0].self_attn.linear_k.weight.requires_grad_(False)
    fg2 = FlatGradients(enc2)
    b2 = fg2.block("layers.0.self_attn.linear_q.weight", 3)
    assert b2 is None or b2.numel() != 3 * 128 * 128


def test_noam_restatement_matches_the_reference_scheduler():
    """optim.noam_lr (what B200FusedAdamW evaluates on the device) vs NoamAnnealing._noam_annealing of the reference
    (nemo/core/optim/lr_scheduler.py:566-575), extracted from the source file when /root/reference is present."""
    import ast

    from nemo_b200.optim import noam_lr

    path = "/root/reference/nemo/core/optim/lr_scheduler.py"
    if not os.path.exists(path):
        pytest.skip("/root/reference not present (GPU box)")
    src = open(path).read()
    tree = ast.parse(src)
    fn = next(n for c in tree.body if isinstance(c, ast.ClassDef) and c.name == "NoamAnnealing" for n in c.body
              if isinstance(n, ast.FunctionDef) and n.name == "_noam_annealing")
    ns = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)  # the reference's own method body, unmodified

    class S:  # the attributes the method reads (NoamAnnealing.__init__, :530-546)
        pass

    for d_model, warm, min_lr in ((512, 10000, 1e-6), (1024, 25000, 1e-6), (176, 0, 1e-5)):
        s = S()
        s._normalize, s.warmup_steps, s.min_lr = d_model ** -0.5, warm, min_lr
        for opt_step in (1, 2, 3, 100, 9999, 10000, 10001, 10002, 50000, 10 ** 7):
            want = ns["_noam_annealing"](s, initial_lr=2.0, step=max(1, opt_step - 1))  # get_lr: step = max(1, last_epoch)
            assert noam_lr(2.0, d_model, warm, min_lr, opt_step) == pytest.approx(want, rel=1e-12), (d_model, opt_step)


@pytest.mark.parametrize("rel,params_m", [("examples/asr/conf/conformer/conformer_ctc_bpe.yaml", 121.44),
                                          ("examples/asr/conf/fastconformer/fast-conformer_ctc_bpe.yaml", None),
                                          ("examples/asr/conf/conformer/conformer_transducer_bpe.yaml", None)])
def test_modules_instantiate_from_the_reference_training_yamls(rel, params_m):
    """the encoder / preprocessor sections of the reference's shipped training configs (with their `${...}` interpolations
    resolved) are accepted kwarg-for-kwarg by the B200 modules; Conformer-CTC-Large has the 121.44 M encoder parameters of
    BASELINE.md §2"""
    import yaml

    from nemo_b200 import checkpoint as ck

    path = os.path.join("/root/reference", rel)
    if not os.path.exists(path):
        pytest.skip("/root/reference not present (GPU box)")
    cfg = ck.resolve_config(yaml.safe_load(open(path)))
    m = cfg["model"]
    assert m["encoder"]["feat_in"] == 80 and m["preprocessor"]["sample_rate"] == 16000
    enc = ck.encoder_from_config(m["encoder"])
    pre = ck.preprocessor_from_config(m["preprocessor"])
    assert pre.featurizer.nfilt == 80 and pre.featurizer.n_fft == 512 and pre.featurizer.hop_length == 160
    assert enc.n_layers == m["encoder"]["n_layers"] and enc.subsampling == m["encoder"]["subsampling"]
    if params_m is not None:
        assert abs(sum(p.numel() for p in enc.parameters()) / 1e6 - params_m) < 0.01

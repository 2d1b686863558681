"""2-rank data-parallel correctness ON HARDWARE: gradients after the layer-wise overlapped NCCL all-reduce of
FlatGradients == single-rank gradients of the concatenated batch.  BatchNorm runs on running statistics (eval) so the
utterances are independent — with batch statistics the two differ by construction (SyncBatchNorm is deliberately not
reproduced).  With >= 2 GPUs the ranks sit on different devices and reduce over NCCL (`gpurun --gpus 2`); on a
one-GPU box both ranks share cuda:0 and reduce over gloo (NCCL refuses two ranks on one device) — the kernels, the flat
buffer, the bucketing and the all-reduce calls are the same."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(n_layers=2, d_model=128, n_heads=2, conv_kernel_size=9, subsampling_kind="striding", subsampling_factor=4)


def _build(dev):
    from nemo_b200.encoder import B200ConformerEncoder
    from oracle import conformer_oracle as O

    sd = O.init_state_dict(seed=5, **CFG)
    enc = B200ConformerEncoder(feat_in=80, n_layers=2, d_model=128, n_heads=2, conv_kernel_size=9, dropout=0.0, dropout_pre_encoder=0.0,
                               dropout_emb=0.0, dropout_att=0.0)
    enc.load_state_dict(sd)
    enc = enc.to(dev).eval()
    with torch.no_grad():
        for lyr in enc.layers:
            lyr.conv.batch_norm.running_mean.normal_(0, 0.1, generator=torch.Generator(device=dev).manual_seed(1))
            lyr.conv.batch_norm.running_var.uniform_(0.75, 1.25, generator=torch.Generator(device=dev).manual_seed(2))
    return enc


def _data():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 80, 203, generator=g)
    ln = torch.tensor([203, 160, 97, 203])
    for b in range(4):
        x[b, :, ln[b]:] = 0
    r = torch.randn(4, 128, 51, generator=g)
    return x, ln, r


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from nemo_b200.parallel import FlatGradients, shard_batch

    ngpu = torch.cuda.device_count()
    dev = torch.device("cuda", rank % ngpu)
    torch.cuda.set_device(dev)
    if ngpu >= world:
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        enc = _build(dev)
        fg = FlatGradients(enc)
        x, ln, r = _data()
        lo, hi = shard_batch(4, rank, world)
        fg.zero()
        y, _ = enc(audio_signal=x[lo:hi].to(dev), length=ln[lo:hi].to(dev))
        ((y.float() * r[lo:hi].to(dev)).sum() / 4.0).backward()
        fg.finish()
        torch.cuda.synchronize()
        if rank == 0:
            out["flat"] = fg.flat.detach().cpu()
            out["names"] = list(fg.names)
            out["offsets"] = dict(fg.offsets)
        # a second backward without zero() must be refused (it would reduce the first micro-batch twice)
        y, _ = enc(audio_signal=x[lo:hi].to(dev), length=ln[lo:hi].to(dev))
        try:
            ((y.float() * r[lo:hi].to(dev)).sum() / 4.0).backward()
            refused = False
        except RuntimeError:
            refused = True
        if rank == 0:
            out["refused"] = refused
        dist.barrier()
    except Exception:  # noqa: BLE001
        import traceback

        out[f"error{rank}"] = traceback.format_exc()
        raise
    finally:
        try:
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass


def test_two_rank_allreduced_gradients_equal_single_rank_global_batch():
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    out = mgr.dict()
    try:
        mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    except Exception:  # noqa: BLE001
        for k in ("error0", "error1"):
            if k in out:
                print(k, out[k])
        raise
    dev = torch.device("cuda", 0)
    enc = _build(dev)
    x, ln, r = _data()
    y, _ = enc(audio_signal=x.to(dev), length=ln.to(dev))
    ((y.float() * r.to(dev)).sum() / 4.0).backward()
    flat = out["flat"]
    worst = 0.0
    rms_all = max(p.grad.float().pow(2).mean().sqrt().item() for p in enc.parameters())
    for n, p in enc.named_parameters():
        lo, hi = out["offsets"][n]
        got, want = flat[lo:hi].view_as(p), p.grad.detach().cpu()
        if want.pow(2).mean().sqrt().item() < 1e-4 * rms_all:  # analytically-zero gradients: rounding noise on both sides
            continue
        e = ((got - want).norm() / want.norm()).item()
        worst = max(worst, e)
        # speed mode rounds the summed position-side score gradient (and a few other reduced tensors) to bf16 once PER RANK
        # before the last GEMM, the single-rank run once for the whole batch: bf16-level differences (4e-3 observed on
        # linear_pos.weight); a sharding / all-reduce bug would be O(1)
        assert e < 1e-2, (n, e)
    print("2-rank all-reduced vs single-rank gradients: worst rel-L2", worst)
    assert out["refused"] is True

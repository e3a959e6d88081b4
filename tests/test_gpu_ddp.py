"""GPU parity of the data-parallel training path (BASELINE configs[3]; reference: DDP over NCCL,
configs/_base_/trainers/base.py:30-41): gradients of 2 ranks x per-rank batch b, averaged by the overlapped bucketed
all-reduce launched from inside the native backward (train.GradSync), equal the gradients of ONE process on the
concatenated batch.  Needs 2 GPUs (skipped otherwise); NCCL over 127.0.0.1."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(mel_channels=64, d_encoder=64, residual_channels=128, residual_layers=5, use_linear_bias=True, dilation_cycle=4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev, seed=0):
    from fish_diffusion_b200 import DIFFUSIONS, synthetic
    diff = DIFFUSIONS.build(dict(type="GaussianDiffusion", denoiser=dict(type="WaveNetDenoiser", **CFG),
                                 mel_channels=CFG["mel_channels"], noise_loss="smoothed-l1", sampler_interval=10,
                                 spec_min=[-5.0], spec_max=[0.0])).to(dev)
    diff.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in synthetic.wavenet_weights(seed, **CFG).items()})
    return diff.train()


def _batch(world, B, T):
    g = torch.Generator().manual_seed(100)
    return (torch.randn(world * B, T, CFG["d_encoder"], generator=g), torch.rand(world * B, T, CFG["mel_channels"], generator=g) * 5 - 5,
            torch.randint(0, 1000, (world * B,), generator=g), torch.randn(world * B, CFG["mel_channels"], T, generator=g))


def _worker(rank, world, port, B, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    torch.distributed.init_process_group("nccl", rank=rank, world_size=world)
    from fish_diffusion_b200.train import DenoiserTrainer
    feats, mel, t, noise = _batch(world, B, T)
    sl = slice(rank * B, (rank + 1) * B)
    diff = _build(dev)
    tr = DenoiserTrainer(diff, device=dev, sync="bucketed", bucket_layers=2, clip=0)
    assert tr.sync is not None
    loss = tr.module(feats[sl].to(dev), mel[sl].to(dev), t=t[sl].to(dev), noise=noise[sl].to(dev))
    loss.backward()
    tr.sync.wait()
    tr._reduce_rest()
    assert diff.denoise_fn._synced_in_backward and tr.sync.bytes > 0
    if rank == 0:
        q.put({k: p.grad.detach().cpu() for k, p in diff.named_parameters()})
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_gradients_equal_single_process_on_the_global_batch():
    import torch.multiprocessing as mp
    B, T, world = 3, 200, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port() if r == 0 else 0, B, T, q)) for r in range(world)]
    port = procs[0]._args[2]
    for p in procs:
        p._args = p._args[:2] + (port,) + p._args[3:]
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    dev = torch.device("cuda", 0)
    feats, mel, t, noise = _batch(world, B, T)
    ref = _build(dev)
    from fish_diffusion_b200.train import TrainStepModule
    TrainStepModule(ref)(feats.to(dev), mel.to(dev), t=t.to(dev), noise=noise.to(dev)).backward()
    worst = 0.0
    for k, p in ref.named_parameters():
        e = float((got[k].to(dev) - p.grad).norm() / p.grad.norm().clamp_min(1e-30))
        worst = max(worst, e)
        # fp32 tensor-core accumulation is order dependent: the split changes the partial sums
        assert e < 1e-3, (k, e)
    print("worst relative gradient difference 2 ranks vs 1:", worst)

"""CPU tests of the N>1 path: batch sharding over a world_size-2 gloo group (no data-path collective)."""
import os

import numpy as np
import torch
import torch.multiprocessing as mp

from fish_diffusion_b200.dist import gather_batch, item_seed, max_over_ranks, shard_range


def test_shard_range_partitions():
    for n in (1, 2, 5, 16, 32, 33):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert item_seed(3, 10) == item_seed(3, 10) != item_seed(3, 11)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    B = 5
    full = torch.arange(B * 3, dtype=torch.float32).reshape(B, 3)
    lo, hi = shard_range(B, rank, world)
    # per-item work whose RNG stream depends on the GLOBAL item index only
    local = torch.stack([full[i] * 2 + torch.Generator().manual_seed(item_seed(7, i) % (2 ** 31)).initial_seed() % 5
                         for i in range(lo, hi)])
    out = gather_batch(local, B)
    want = torch.stack([full[i] * 2 + item_seed(7, i) % (2 ** 31) % 5 for i in range(B)])
    ok = torch.equal(out, want)
    t = max_over_ranks(1.0 + rank)
    ret[rank] = (ok, t)
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world))
    assert all(ret[r][1] == 2.0 for r in range(world))          # max over ranks


def _sync_worker(rank, world, port, ret):
    """GradSync (the bucketed all-reduce the native backward launches) and DenoiserTrainer._reduce_rest on gloo/CPU."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from fish_diffusion_b200.train import GradSync
    sync = GradSync(bucket_layers=2)
    a = torch.full((2, 4, 3), float(rank + 1))
    b = torch.full((2, 4), 10.0 * (rank + 1))
    sync.reduce_async(a, b)                      # asynchronous: the caller keeps working, joins later
    c = torch.full((5,), float(rank))
    sync.reduce_async(c)
    sync.wait()
    ok = bool((a == 1.5).all() and (b == 15.0).all() and (c == 0.5).all()) and not sync.handles and sync.bytes > 0
    ret[rank] = ok
    torch.distributed.destroy_process_group()


def test_grad_sync_buckets_average_over_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_sync_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def _voc_sync_worker(rank, world, port, ret):
    """vocoder_gan.average_gradients (the DDP role of the vocoder trainer, configs/vocoder_nsf_hifigan.py:25) on gloo/CPU:
    per-network averaging, parameters without a gradient are left alone."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from fish_diffusion_b200.vocoder_gan import DiscriminatorP, average_gradients
    torch.manual_seed(0)                                         # same weights on every rank
    d = DiscriminatorP(3)
    x = torch.randn(2, 1, 300, generator=torch.Generator().manual_seed(100 + rank))    # different data per rank
    out, _ = d(x)
    out.square().mean().backward()
    local = [p.grad.clone() for p in d.parameters()]
    extra = torch.nn.Parameter(torch.ones(3))                    # never used: no gradient, must stay None
    average_gradients(list(d.parameters()) + [extra])
    gathered = [None] * world
    torch.distributed.all_gather_object(gathered, [g.numpy() for g in local])
    want = [sum(torch.from_numpy(gathered[r][i]) for r in range(world)) / world for i in range(len(local))]
    ok = all(torch.allclose(p.grad, w, rtol=1e-6, atol=1e-8) for p, w in zip(d.parameters(), want)) and extra.grad is None
    ret[rank] = bool(ok)
    torch.distributed.destroy_process_group()


def test_vocoder_trainer_gradient_average_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_voc_sync_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def _voc_step_worker(rank, world, port, ret):
    """One whole HifiGanTrainer.training_step per rank on different data with the gradient-averaging hook (gloo): the
    replicas stay identical.  Device primitives are emulated (tests/native_emu.py) -- this checks the host-side data-parallel
    plumbing of the vocoder trainer, the role of DDPStrategy in configs/vocoder_nsf_hifigan.py:25."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    import n4_util as nu
    from native_emu import emulated_native
    from fish_diffusion_b200.vocoder_gan import HifiGanTrainer, average_gradients
    h = dict(nu.train_config(), discriminator_periods=[2])
    tr = HifiGanTrainer(h, precision="f16")
    sd = torch.load(os.path.join(here, "golden", "ref_generator_small.ckpt"), map_location="cpu")["generator"]
    tr.generator.load_state_dict({k: v / 3 for k, v in sd.items()})
    nu.fill_discriminators(tr.mpd, tr.msd)
    tr.train()
    batch = nu.make_batch()
    S = 4096                                                     # the 4096-point loss mel needs more than 2048 samples
    batch = dict(pitches=batch["pitches"][rank:rank + 1, :, :S // 64], audio=batch["audio"][rank:rank + 1, :, :S],
                 audio_lens=torch.tensor([S]))
    mel = torch.randn(1, 32, S // 64, generator=torch.Generator().manual_seed(7 + rank)) - 2.0
    batch["mels"] = mel
    with emulated_native():
        out = tr.training_step(batch, reduce_grads=average_gradients)
    digest = torch.cat([p.detach().double().reshape(-1)[:64] for p in tr.parameters()])
    gathered = [None] * world
    torch.distributed.all_gather_object(gathered, (digest.numpy(), out["loss_gen"]))
    same = all(float(abs(gathered[0][0] - g[0]).max()) == 0.0 for g in gathered)
    ret[rank] = (bool(same), gathered[0][1] != gathered[1][1])    # identical replicas, different local losses
    torch.distributed.destroy_process_group()


def test_vocoder_training_step_two_ranks_stay_identical_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 35500 + (os.getpid() % 2000)
    mp.spawn(_voc_step_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world)), dict(ret)
    assert all(ret[r][1] for r in range(world))

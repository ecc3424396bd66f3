"""Host-side logic of the data-parallel trainer on CPU with the gloo backend, world_size 2 (no GPU, no kernels):
flat [decoder | LoRA] parameter buffer, two contiguous gradient buckets, AVG all-reduce, identical clip + update on every
rank, tile sharding of an inference sweep."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from miphei_vit_b200.generators.mipheivit import get_vitmatte
        from miphei_vit_b200.trainer import FlatParams, lr_lambda, shard_tiles

        torch.manual_seed(0)  # same weights everywhere
        m = get_vitmatte("hoptimus0", 128, 3, use_lora=True, embed_dim=128, depth=2, num_heads=2, hidden=256)
        flat = FlatParams(m, device="cpu")
        # layout: decoder segment first, LoRA after; every parameter is a view of the flat buffer
        assert flat.n_dec > 0 and flat.total > flat.n_dec
        names = [n for n, _ in flat.order]
        first_lora = next(i for i, n in enumerate(names) if not n.startswith("decoder."))
        assert all(n.startswith("decoder.") for n in names[:first_lora])
        assert all(".lora_" in n for n in names[first_lora:])
        for _, p in flat.order:
            assert p.data.untyped_storage().data_ptr() == flat.flat.untyped_storage().data_ptr()
            assert p.grad.untyped_storage().data_ptr() == flat.gflat.untyped_storage().data_ptr()
        # the LoRA segment is laid out block by block, 32 * D floats each: the trainer reduces it in block-range slices
        # (upper half of the blocks while the lower half is still in backward) and mv_lora_refresh reads it that way
        D, off = 128, flat.n_dec
        offsets = {}
        for n, p in flat.order:
            offsets[n] = (p.data.data_ptr() - flat.flat.data_ptr()) // 4
        for i in range(2):
            for j, t in enumerate(("lora_q.A", "lora_q.B", "lora_v.A", "lora_v.B")):
                assert offsets["encoder.vit.blocks.%d.attn.qkv.%s" % (i, t)] == off + i * 32 * D + j * 8 * D
        assert flat.total == off + 2 * 32 * D
        # rank-dependent fake gradients -> bucketed AVG all-reduce -> identical on both ranks, equal to the mean
        g = torch.Generator().manual_seed(100 + rank)
        flat.gflat.copy_(torch.randn(flat.total, generator=g))
        mine = flat.gflat.clone()
        flat.allreduce_bucket(0, async_op=False)
        flat.allreduce_bucket(1, async_op=False)
        other = torch.randn(flat.total, generator=torch.Generator().manual_seed(100 + (1 - rank)))
        assert torch.allclose(flat.gflat, (mine + other) / 2, atol=1e-6)
        gathered = [torch.empty_like(flat.gflat) for _ in range(world)]
        dist.all_gather(gathered, flat.gflat)
        assert torch.equal(gathered[0], gathered[1])
        # clip coefficient from the reduced gradient is the same everywhere
        norm = flat.gflat.double().norm().float()
        coef = torch.clamp(1.0 / (norm + 1e-6), max=1.0)
        allc = [torch.empty_like(coef) for _ in range(world)]
        dist.all_gather(allc, coef)
        assert torch.equal(allc[0], allc[1])
        # tile sharding: disjoint, complete, balanced
        idx = shard_tiles(1001, rank, world)
        cnt = torch.tensor([len(idx)])
        dist.all_reduce(cnt)
        assert int(cnt) == 1001 and abs(len(idx) - 1001 / world) <= 1
        every = [None] * world
        dist.all_gather_object(every, list(idx))
        assert sorted(every[0] + every[1]) == list(range(1001))
        assert lr_lambda(0, 1000) == 0.0 and lr_lambda(400, 1000) == 1.0 and abs(lr_lambda(750, 1000) - 0.5) < 1e-12
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


def test_flat_buckets_allreduce_and_tile_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", "rank %d: %s" % (rank, msg)

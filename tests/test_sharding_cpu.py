"""CPU: batch sharding logic, single-process and across two gloo ranks (the N>1 host path; no GPU involved)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nice_diffusion_b200.sharding import global_indices, output_names, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 16, 255, 256):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(total, r, world)
                assert 0 <= lo <= hi <= total
                seen += list(range(lo, hi))
            assert seen == list(range(total))
            sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert list(global_indices(2, 256, 3, 8)) == list(range(2 * 256 + 96, 2 * 256 + 128))


def test_output_names_match_reference_rule():
    assert output_names([3, 3, 5, 3, None]) == ['3_sample0.jpg', '3_sample1.jpg', '5_sample0.jpg', '3_sample2.jpg',
                                                 'sample0.jpg']


def _fake_sample(gid):
    """Stand-in for 'the image of global sample id gid': depends on the id only."""
    g = torch.Generator().manual_seed(1000 + gid)
    return torch.randn(3, 4, 4, generator=g)


def _worker(rank, world, port, batch, n_batches, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        mine = []
        for i in range(n_batches):
            mine += [(g, _fake_sample(g)) for g in global_indices(i, batch, rank, world)]
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)       # test-only gather; the data path itself has no collective
        if rank == 0:
            flat = sorted([x for part in gathered for x in part], key=lambda t: t[0])
            torch.save(flat, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('batch', [6, 7])
def test_two_rank_sharding_reproduces_single_process(tmp_path, batch):
    world, n_batches = 2, 3
    out = str(tmp_path / 'gathered.pt')
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, batch, n_batches, out), nprocs=world, join=True)
    flat = torch.load(out)
    assert [g for g, _ in flat] == list(range(batch * n_batches))
    for g, img in flat:
        assert torch.equal(img, _fake_sample(g))

"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (goldrush_b200/multi.py).
Pass-1 sharding + bitwise-OR all-reduce of partial bit vectors, the tile shares of the pass-2 query,
the unique-id broadcast shape and the replica-agreement check.  No compute call is made here; the
bit positions come from the CPU oracle's hashes of synthetic reads."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_util as ou
from goldrush_b200 import multi

SEED22 = "1011011110110111101101"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partial_words(reads, lo, hi, seeds, bits):
    words = np.zeros((bits + 63) // 64, dtype=np.uint64)
    for seq in reads[lo:hi]:
        hv = ou.hash_sequence(seq, seeds).reshape(-1)
        pos = hv % np.uint64(bits)
        np.bitwise_or.at(words, (pos >> np.uint64(6)).astype(np.int64),
                         np.uint64(1) << (pos & np.uint64(63)))
    return words


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        reads = ["".join(rng.choice(list("ACGT"), size=int(n))).encode()
                 for n in rng.integers(60, 400, size=9)]
        seeds = ou.make_seed_pattern(SEED22, 22, 16, 3)
        bits = 4096 + 64
        lo, hi = multi.shard_range(len(reads), rank, world)
        mine = torch.from_numpy(_partial_words(reads, lo, hi, seeds, bits).view(np.int64).copy())
        multi.or_allreduce(mine)
        whole = _partial_words(reads, 0, len(reads), seeds, bits).view(np.int64)
        assert np.array_equal(mine.numpy(), whole), "OR-reduced shards differ from the whole fill"

        # replica agreement: equal decisions pass, different decisions raise on every rank
        multi.assert_replicas_agree(np.arange(12, dtype=np.uint32))
        with pytest.raises(RuntimeError):
            multi.assert_replicas_agree(np.arange(12, dtype=np.uint32) + rank)

        # the unique-id broadcast carries 128 bytes from rank 0
        class FakeApi:
            def comm_unique_id(self):
                return bytes(range(128))

            def comm_init(self, ident, r, w, device):
                self.got = (ident, r, w, device)

        fa = FakeApi()
        assert multi.init_comm(0, api=fa) == (rank, world)
        assert fa.got == (bytes(range(128)), rank, world, 0)
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_world2_gloo_pass1_or_reduce_and_plumbing(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_polish_batches_are_dealt_to_ranks_exactly_once():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 64):
            dealt = sorted(b for r in range(world) for b in multi.polish_batch_share(n, r, world))
            assert dealt == list(range(n))


def test_shares_cover_and_do_not_overlap():
    for n in (0, 1, 7, 128, 3201):
        for world in (1, 2, 3, 4, 8):
            spans = [multi.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
            tiles = [multi.tile_share(n, r, world) for r in range(world)]
            assert tiles[0][0] == 0 and tiles[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(tiles, tiles[1:]))
            c = multi.tile_chunk(n, world)
            assert all(h - l <= c for l, h in tiles) and c * world >= n

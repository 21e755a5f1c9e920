"""Multi-GPU plumbing of the GoldRush-Path engine: one process per GPU, `torch.distributed` for
rendezvous and the collectives, the C ABI for everything that touches a filter.

What shards (SURVEY.md 8e / DESIGN.md 6):

* pass 1 (goldrush_path.cpp:235-339): reads are independent and the bit OR is commutative, so rank r
  hashes reads [r*n/W, (r+1)*n/W) into its own zeroed bit vector and the partial vectors are
  OR-reduced.  NCCL and gloo have no bitwise-OR reduction: all-gather the partial vectors and OR
  them locally (`or_allreduce`); on a GPU the local OR is the library's `grb_or_words` kernel.
* pass 2 speculative query (goldrush_path.cpp:544-626): the tiles of one batch are cut into W equal
  chunks (`tile_chunk`), rank r queries chunk r inside `grb_select_reads`, and the per-tile results
  are all-gathered over NVLink by the library's own NCCL communicator (`init_comm`, csrc/comm.cuh).
  With that communicator in place `grb_build_bitvector` also shards and OR-reduces pass 1 natively;
  `or_allreduce` / `build_bitvector_sharded` below remain for callers that bring their own
  collectives (and for the gloo tests).
* the ordered commit (goldrush_path.cpp:1229-1256) does not shard: it is replicated, integer-only and
  deterministic, so every replica ends each batch with the same filter and the same decisions
  (`assert_replicas_agree`).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous share of n items for `rank`: [lo, hi).  Shares differ by at most one item."""
    return rank * n // world, (rank + 1) * n // world


def polish_batch_share(n_batches, rank, world):
    """(f4) GoldPolish Bloom-filter batches are independent objects (one OpenMP task each in the
    reference, goldpolish_targeted_bfs.cpp:181-196): rank r serves batches r, r + W, r + 2W, ... with
    its own grb_polish_serve_batches / grb_polish_fill_batches call and writes their filters itself.
    No collective."""
    return list(range(rank, n_batches, world))


def tile_chunk(n_tiles, world):
    """Tiles per rank of one batch: the library pads a batch to world * tile_chunk tiles so that
    the all-gather is uniform; rank r owns tiles [r*chunk, min((r+1)*chunk, n_tiles))."""
    return -(-n_tiles // world)


def tile_share(n_tiles, rank, world):
    c = tile_chunk(n_tiles, world)
    return min(rank * c, n_tiles), min((rank + 1) * c, n_tiles)


def or_allreduce(words, group=None, or_into=None):
    """Bitwise-OR all-reduce of an int64 tensor, in place.  `or_into(dst, src)` does dst |= src;
    default is torch.bitwise_or (the CPU / gloo test path); on a GPU pass the engine's kernel."""
    world = dist.get_world_size(group)
    if world == 1:
        return words
    rank = dist.get_rank(group)
    flat = torch.empty(world * words.numel(), dtype=words.dtype, device=words.device)
    dist.all_gather_into_tensor(flat, words.contiguous().view(-1), group=group)
    gathered = flat.view((world,) + tuple(words.shape))
    for r in range(world):
        if r == rank:
            continue
        if or_into is None:
            words.bitwise_or_(gathered[r])
        else:
            or_into(words, gathered[r])
    return words


class _DevMem:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<i8",
                                         "data": (ptr, False), "version": 2}


def device_words(ptr, nbytes, device):
    """int64 tensor view of raw device memory owned by the library."""
    return torch.as_tensor(_DevMem(ptr, nbytes), device=torch.device("cuda", device))


def build_bitvector_sharded(eng, n_reads, device, stream, group=None):
    """Pass 1 on W GPUs: hash this rank's share of the reads, OR-reduce the bit vectors."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        eng.build_bitvector()
        return
    lo, hi = shard_range(n_reads, rank, world)
    eng.build_bitvector(lo, hi - lo)
    ptr, nbytes = eng.bitvector_device()
    mine = device_words(ptr, nbytes, device)
    with torch.cuda.stream(stream):
        or_allreduce(mine, group,
                     or_into=lambda d, s: eng.or_words(d.data_ptr(), s.data_ptr(), d.numel()))
    eng.sync()


def init_comm(device, group=None, api=None):
    """Creates the library's process-wide NCCL communicator (one process per GPU): rank 0 makes the
    NCCL unique id, torch.distributed carries it to the other ranks (works over nccl or gloo).
    Call before creating the Engine / calling run_path; returns (rank, world).  `api` is the
    module providing comm_unique_id / comm_init (goldrush_b200.api; injectable for CPU tests)."""
    if api is None:
        from . import api
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return rank, world
    ident = api.comm_unique_id() if rank == 0 else bytes(128)
    dev = torch.device("cuda", device) if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(list(ident), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=group)
    api.comm_init(bytes(t.cpu().tolist()), rank, world, device)
    return rank, world


def assert_replicas_agree(decisions, group=None):
    """Every replica must have taken the same decisions (the commit is replicated)."""
    world = dist.get_world_size(group)
    if world == 1:
        return
    import hashlib
    dig = hashlib.sha256(np.ascontiguousarray(decisions).tobytes()).digest()
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor(list(dig), dtype=torch.uint8, device=dev)
    flat = torch.empty(world * 32, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(flat, mine, group=group)
    allv = flat.view(world, 32)
    if not bool((allv == allv[0]).all().item()):
        raise RuntimeError("replicas disagree on the selection decisions")

"""Multi-GPU plumbing: prompts shard by batch across ranks, one process per GPU, no collective on the hot path.

The reference has no multi-GPU inference at all (single process, `CUDA_VISIBLE_DEVICES=$which_gpu`,
inference.sh:13); clips are independent end to end, so the only communication is gathering the int16 waveforms
(320 KB / clip) for output and, optionally, a 2-float all-reduce that reproduces the reference's *batch-global*
waveform centring (hifigan/utilities.py:84-85) across shards.  torch.distributed (NCCL on GPUs, gloo in the CPU
tests) is plumbing only.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size, rank):
    """Contiguous, balanced split: ranks [0, n % world) get one extra item.  Returns (start, stop)."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard(seq, world_size=None, rank=None):
    world_size = world_size if world_size is not None else dist.get_world_size()
    rank = rank if rank is not None else dist.get_rank()
    a, b = shard_bounds(len(seq), world_size, rank)
    return seq[a:b]


def sharded_noise(n_global, seed, world_size, rank, shape=(8, 256, 16)):
    """Rows [start, stop) of the noise a single process would draw for the whole batch with `seed` (so the
    concatenation over ranks equals the single-process tensor bit for bit)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    full = torch.randn((n_global,) + tuple(shape), generator=g)
    a, b = shard_bounds(n_global, world_size, rank)
    return full[a:b].clone()


def global_minmax(minmax):
    """All-reduce of the (min, max) pair used for waveform centring; no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        lo, hi = minmax[0:1].clone(), minmax[1:2].clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return torch.cat([lo, hi])
    return minmax


def gather_waveforms(local, n_global=None, sizes=None):
    """All-gathers int16 [B_local, T] shards (possibly unequal B_local) into [n_global, T] on every rank.
    `sizes` (rows per rank) skips the size exchange and its host synchronisation when the split is known, e.g.
    [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if sizes is None:
        sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device))
        sizes = [int(s.item()) for s in sizes]
    assert len(sizes) == world and sizes[dist.get_rank()] == local.shape[0]
    m = max(sizes)
    # neither NCCL nor gloo transports int16: ship the raw bytes
    row_bytes = local.element_size()
    for d in local.shape[1:]:
        row_bytes *= d
    raw = local.contiguous().view(torch.uint8).reshape(local.shape[0], row_bytes)
    if local.shape[0] < m:
        raw = torch.cat([raw, raw.new_zeros((m - local.shape[0], raw.shape[1]))])
    bufs = [torch.empty_like(raw) for _ in range(world)]
    dist.all_gather(bufs, raw.contiguous())
    out = torch.cat([b[:n] for b, n in zip(bufs, sizes)]).view(local.dtype).reshape((-1,) + tuple(local.shape[1:]))
    if n_global is not None:
        assert out.shape[0] == n_global
    return out

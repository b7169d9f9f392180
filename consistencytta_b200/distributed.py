"""Multi-GPU plumbing: prompts shard by batch across ranks, one process per GPU, no collective on the hot path.

The reference has no multi-GPU inference at all (single process, `CUDA_VISIBLE_DEVICES=$which_gpu`,
inference.sh:13); clips are independent end to end, so the only communication is collecting the int16 waveforms
(320 KB / clip) on rank 0 for output (`WaveformGatherer`: a gather to ONE rank on a side stream, overlapped with the next
step's compute; `gather_waveforms` is the all-ranks variant) and, optionally, a 2-float all-reduce that reproduces the reference's *batch-global*
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


class WaveformGatherer:
    """Collects every rank's int16 clips [B_local, T] on rank 0 — the only rank that writes files (inference.py:221-222)
    — without stalling the compute stream.

    `submit(local)` copies the clips into one of two staging buffers on the caller's stream (the engine's output buffer is
    overwritten by the next replay) and issues `dist.gather(dst=0)` on a side stream; the caller's stream never waits
    for the transfer, only — two submits later — for the staging buffer to be free again.  `result()` makes the caller's
    stream wait for the last transfer and returns rank 0's [n_global, T] tensor (None elsewhere).  Rows are padded to
    the largest shard so unequal shards work; per-rank waveform centring stays the default (each rank is an independent
    `vocoder_infer` batch, hifigan/utilities.py:84-86), `global_minmax` being the opt-in for single-process equality."""

    def __init__(self, sizes, row_shape, device, dtype=torch.int16, rank=None, world=None):
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank() if dist.is_initialized() else 0)
        self.sizes = list(sizes)
        assert len(self.sizes) == self.world
        self.m = max(self.sizes)
        self.row_shape = tuple(row_shape)
        self.dtype = dtype
        self.device = torch.device(device)
        row_bytes = torch.empty((), dtype=dtype).element_size()
        for d in self.row_shape:
            row_bytes *= d
        self.row_bytes = row_bytes
        self.cuda = self.device.type == "cuda"
        # neither NCCL nor gloo transports int16: ship the raw bytes
        self.stage = [torch.zeros(self.m, row_bytes, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.recv = None
        if self.rank == 0 and self.world > 1:
            # one contiguous [world, m, row] buffer per slot: with equal shards the result is a VIEW of it (no torch.cat)
            self.recv = [torch.empty(self.world, self.m, row_bytes, dtype=torch.uint8, device=self.device)
                         for _ in range(2)]
        self.side = torch.cuda.Stream(device=self.device) if self.cuda else None
        self.done = [None, None]
        self.k = 0
        self.last = None

    def submit(self, local):
        assert local.shape[0] == self.sizes[self.rank] and tuple(local.shape[1:]) == self.row_shape
        slot = self.k & 1
        self.k += 1
        if self.world == 1:
            self.last = (slot, local.contiguous().view(torch.uint8).reshape(local.shape[0], self.row_bytes))
            return
        # typed view of the staging slot: a strided `local` (e.g. clips[:, :160000]) is copied exactly once
        stage = self.stage[slot].view(self.dtype).reshape((self.m,) + self.row_shape)
        if self.cuda:
            cur = torch.cuda.current_stream(self.device)
            if self.done[slot] is not None:
                cur.wait_event(self.done[slot])          # the transfer that last used this staging slot has finished
            stage[: local.shape[0]].copy_(local, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(self.side):
                self.side.wait_event(ready)
                dist.gather(self.stage[slot], list(self.recv[slot].unbind(0)) if self.rank == 0 else None, dst=0)
                ev = torch.cuda.Event()
                ev.record(self.side)
            self.done[slot] = ev
        else:
            stage[: local.shape[0]].copy_(local)
            dist.gather(self.stage[slot], list(self.recv[slot].unbind(0)) if self.rank == 0 else None, dst=0)
        self.last = (slot, None)

    def to_host(self, host_out):
        """Rank 0, equal shards: copies the LAST gathered batch into pinned host memory on the side stream (overlapped with
        whatever the caller's stream does next) — the files are written from there (inference.py:221-222)."""
        if self.rank != 0 or self.last is None:
            return
        slot, raw = self.last
        if self.world == 1:
            host_out.view(torch.uint8).reshape(-1, self.row_bytes).copy_(raw, non_blocking=True)
            return
        assert all(n == self.m for n in self.sizes), "to_host needs equal shards"
        src = self.recv[slot].view(self.world * self.m, self.row_bytes)
        dst = host_out.view(torch.uint8).reshape(self.world * self.m, self.row_bytes)
        if self.cuda:
            with torch.cuda.stream(self.side):
                dst.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.side)
            self.done[slot] = ev
        else:
            dst.copy_(src)

    def result(self):
        """Rank 0: [n_global, *row_shape] of the LAST submit (ordered after it on the current stream); others: None."""
        if self.last is None:
            return None
        slot, raw = self.last
        if self.world == 1:
            return raw.view(self.dtype).reshape((-1,) + self.row_shape)
        if self.cuda and self.done[slot] is not None:
            torch.cuda.current_stream(self.device).wait_event(self.done[slot])
        if self.rank != 0:
            return None
        if all(n == self.m for n in self.sizes):
            out = self.recv[slot].view(self.world * self.m, self.row_bytes)
        else:
            out = torch.cat([b[:n] for b, n in zip(self.recv[slot].unbind(0), self.sizes)])
        return out.view(self.dtype).reshape((-1,) + self.row_shape)

"""Multi-GPU sampling: samples are independent, so the sampling batch is split
contiguously over the ranks of one node (one process per GPU, weights replicated, no per-step traffic) and the
final frames are gathered with ONE all-gather (NCCL over NVLink/NVSwitch; gloo in the CPU tests).

Mirrors the reference's data-parallel sampling (HF Accelerate ``split_batches=True`` DataLoader sharding,
tools/evaluation/sample_and_save_cond.py:33-38,63); the reference never gathers (each rank writes its own
files, :157-159), so only the concatenation order is a contract: rank r owns samples [r*B/N, (r+1)*B/N).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split; the first ``total % world`` ranks take one extra sample."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch_dict(batch: dict, rank: int, world: int) -> dict:
    """Slice every batched tensor of a conditioning dict (layout boxes, masks, concat_cond ...)."""
    n = next(v.shape[0] for v in batch.values() if torch.is_tensor(v))
    lo, hi = shard_range(n, rank, world)
    return {k: (v[lo:hi] if torch.is_tensor(v) and v.shape[0] == n else v) for k, v in batch.items()}


def all_gather_samples(x_local: torch.Tensor, total: int) -> torch.Tensor:
    """One collective at the end of sample(): [b_local, ...] per rank -> [total, ...] on every rank (rank order)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x_local
    world = dist.get_world_size()
    sizes = [shard_range(total, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = x_local
    if x_local.shape[0] < nmax:      # ragged last shards: pad to the common size for the collective
        pad = torch.cat([x_local, x_local.new_zeros(nmax - x_local.shape[0], *x_local.shape[1:])])
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous())
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


@torch.inference_mode()
def sample_sharded(ddpm, total_batch: int, num_steps: int, batch_dict: dict | None = None, rng=None,
                   mode: str = "ddim", ddim_eta: float = 0.0, progress: bool = False) -> torch.Tensor:
    """Data-parallel ``ddpm.sample``: every rank samples its shard, then one all-gather.  ``rng`` may be a list of
    ``total_batch`` per-sample generators (the reference's ``setup_rng`` list, utils/inference.py:460-461); each
    rank uses its slice, so the result is independent of the number of GPUs."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(total_batch, rank, world)
    local_rng = rng[lo:hi] if isinstance(rng, list) else rng
    kw = dict(num_steps=num_steps, progress=progress, rng=local_rng, mode=mode, ddim_eta=ddim_eta)
    if batch_dict is None:
        x = ddpm.sample(batch_size=hi - lo, **kw)
    else:
        x = ddpm.sample(shard_batch_dict(batch_dict, rank, world), batch_size=hi - lo, **kw)
    return all_gather_samples(x, total_batch)


@torch.inference_mode()
def generate_sharded(sampler, scenes: list, num_frames: int, num_steps: int, rng=None, **kw) -> torch.Tensor:
    """Data-parallel ``TemporalSampler.generate`` (clips are independent; frames inside a clip are sequential,
    tools/evaluation/sample_and_save_temporal.py:203-333): rank r generates the clips of scenes [lo, hi), then ONE all-gather
    of the [b_local, num_frames, 5, H, W] clips.  ``rng``: one generator per scene (sliced like the scenes) or None."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    total = len(scenes)
    lo, hi = shard_range(total, rank, world)
    local_rng = rng[lo:hi] if isinstance(rng, list) else rng
    if hi > lo:
        clips = sampler.generate(scenes[lo:hi], num_frames=num_frames, num_steps=num_steps, rng=local_rng, **kw)
    else:       # more ranks than clips: this rank only takes part in the collective
        clips = torch.zeros(0, num_frames, 5, sampler.H, sampler.W, device=sampler.device)
    return all_gather_samples(clips, total)

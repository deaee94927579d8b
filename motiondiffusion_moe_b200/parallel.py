"""Data-parallel CFG sampling across the GPUs of one node (BASELINE.json configs[2]).

The reference's only multi-GPU code is a DDP wrapper around training whose gradient all-reduce never
fires; sampling is single-GPU (trainers/ddpm_trainer.py:145-174).  Sequences are
independent in every op of the denoiser, so the batch shards by sequence with NO collective inside the
1000-step loop: each rank owns rows [lo, hi) of the global batch, weights are replicated, the global
initial / per-step noise is drawn from one seed on every rank and sliced, so that the G-GPU result equals
the 1-GPU result for the same seed, and one all_gather assembles the samples at the end.
"""
from typing import Callable, Optional, Tuple

import torch


def shard_range(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first batch % world ranks get one extra sequence."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _step_seed(seed: int, step: int) -> int:
    return (int(seed) * 1000003 + int(step) + 1) & 0x7FFFFFFFFFFFFFFF


def global_noise(shape, seed: int, step: int, device) -> torch.Tensor:
    """The [B, T, F] noise of reverse step `step` (step = -1: the initial x_T), identical on every rank.
    Drawn on the CPU generator so that it does not depend on the device type or the rank's RNG state."""
    g = torch.Generator(device="cpu").manual_seed(_step_seed(seed, step))
    return torch.randn(*shape, generator=g).to(device)


class DeviceNoise:
    """The same contract as global_noise, drawn ON the GPU: every rank seeds its own CUDA Philox generator with the
    step's seed and draws the GLOBAL [B, T, F] tensor (3.3 M normals: microseconds), then slices its rows.  Same seed,
    same shape => the same values on every rank and for every world size, so the G-GPU sample still equals the 1-GPU
    sample bit for bit, and nothing crosses PCIe inside the loop (the CPU draw + upload of global_noise costs ~15 ms
    per step: more than the whole step at 8 sequences per GPU)."""

    def __init__(self, shape, seed: int, device):
        self.shape, self.seed = tuple(shape), int(seed)
        self.gen = torch.Generator(device=device)
        self.buf = torch.empty(self.shape, device=device, dtype=torch.float32)

    def __call__(self, step: int) -> torch.Tensor:
        self.gen.manual_seed(_step_seed(self.seed, step))
        return self.buf.normal_(generator=self.gen)


def slice_kwargs(model_kwargs: dict, lo: int, hi: int) -> dict:
    out = {}
    for k, v in model_kwargs.items():
        if isinstance(v, torch.Tensor) and v.dim() > 0:
            out[k] = v[lo:hi].contiguous()
        elif isinstance(v, (list, tuple)):
            out[k] = list(v[lo:hi])
        else:
            out[k] = v
    return out


def sample_dp(make_stepper: Callable, shape, model_kwargs: dict, num_timesteps: int, seed: int = 0,
              group=None, num_steps: Optional[int] = None, device=None, schedule=None, noise: str = "host",
              layout=None, stepper=None) -> torch.Tensor:
    """Sharded p_sample_loop_with_cfg.  make_stepper(local_shape, local_kwargs) must return an object with
    `.x` (the local state tensor) and `.step(t, noise)` (GaussianDiffusion.make_cfg_stepper does).  Returns
    the full [B, T, F] sample on every rank.  schedule (optional): the (t, t_prev) pairs of a strided DDIM loop
    (zip(*GaussianDiffusion.ddim_timesteps(n)) with a stepper built with sampler="ddim"): each step is then
    `.step(t, noise, ts_prev=t_prev)`.  noise: "host" (global_noise: CPU generator, device independent) or "device"
    (DeviceNoise: the same global tensor drawn by every rank on its GPU; CUDA steppers only).  layout=(world, rank)
    overrides the process group's (e.g. (1, 0): the unsharded loop inside a multi-rank job, for comparisons);
    stepper: reuse an already built (captured) stepper of this rank's shard shape instead of calling make_stepper."""
    import torch.distributed as dist
    if layout is not None:
        world, rank = layout
    else:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = shape[0]
    lo, hi = shard_range(B, world, rank)
    st = stepper if stepper is not None else make_stepper((hi - lo,) + tuple(shape[1:]), slice_kwargs(model_kwargs, lo, hi))
    dev = device if device is not None else st.x.device
    if noise == "device":
        dn = DeviceNoise(shape, seed, dev)
        draw = lambda step: dn(step)[lo:hi]
    elif noise == "host":
        draw = lambda step: global_noise(shape, seed, step, dev)[lo:hi]
    else:
        raise ValueError("noise must be 'host' or 'device'")
    st.x.copy_(draw(-1))
    if schedule is not None:
        for t, t_prev in schedule:
            st.step(t, draw(t), ts_prev=t_prev)
    else:
        steps = list(reversed(range(num_timesteps)))
        if num_steps is not None:
            steps = steps[:num_steps]
        for t in steps:
            st.step(t, draw(t))
    local = st.x.contiguous()
    if world == 1:
        return local.clone()
    # ragged all_gather: pad every shard to the largest one
    mx = (B + world - 1) // world
    pad = torch.zeros((mx,) + tuple(shape[1:]), dtype=local.dtype, device=local.device)
    pad[:hi - lo] = local
    if pad.is_cuda and dist.get_backend(group) == "gloo":     # gloo job driving GPUs (tests on a shared GPU): stage on the host
        hp = pad.cpu()
        hparts = [torch.empty_like(hp) for _ in range(world)]
        dist.all_gather(hparts, hp, group=group)
        parts = [h.to(pad.device) for h in hparts]
    else:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
    out = []
    for r in range(world):
        a, b = shard_range(B, world, r)
        out.append(parts[r][:b - a])
    return torch.cat(out)

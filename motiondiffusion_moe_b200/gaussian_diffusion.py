"""GaussianDiffusion — drop-in for the sampling / q-sample surface of the reference
(text2motion/models/gaussian_diffusion.py:375-1141) that the trainer uses
(trainers/ddpm_trainer.py:43-50,161): linear-beta DDPM, epsilon prediction, FIXED_SMALL variance,
classifier-free guidance.

What is B200-native here: the conditional and unconditional branches of p_sample_with_cfg run as ONE
batched MotionTransformer forward (2B sequences, per-sequence text lengths), the whole
eps -> x0 -> guidance -> posterior mean -> noise update is one kernel (mdm_cfg_update) with fp32 schedule
tables resident on the device, the text-side projections are computed once per loop, and the 1000-step
loop replays a CUDA graph of the step.
"""
import enum
import math

import numpy as np
import torch

from . import ops
from ._lib import MdmError


class ModelMeanType(enum.Enum):      # gaussian_diffusion.py:346-352
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):       # :355-362
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):           # :365-372
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """gaussian_diffusion.py:19-55 (float64)."""
    n = num_diffusion_timesteps
    if schedule_name == "linear":
        scale = 1000 / n
        return np.linspace(scale * 0.0001, scale * 0.02, n, dtype=np.float64)
    if schedule_name == "cosine":
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        return np.array([min(1 - f((i + 1) / n) / f(i / n), 0.999) for i in range(n)], dtype=np.float64)
    if schedule_name == "sqrt":
        a = np.linspace(1.0, 0.0, n, dtype=np.float64)
        b = 1 - a ** 2
        b = (b - b.min()) / (b.max() - b.min())
        return b * (0.02 - 0.0001) + 0.0001
    raise NotImplementedError("unknown beta schedule: %s" % schedule_name)


class GaussianDiffusion:
    def __init__(self, *, betas, model_mean_type=ModelMeanType.EPSILON, model_var_type=ModelVarType.FIXED_SMALL,
                 loss_type=LossType.MSE, rescale_timesteps=False, cfg_scale=7.5):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.rescale_timesteps = rescale_timesteps
        self.cfg_scale = cfg_scale
        betas = np.array(betas, dtype=np.float64)                     # :396-431
        assert betas.ndim == 1 and (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self._dev_tables = {}
        self._graphs = {}

    # ------------------------------------------------------------------ device tables
    def _tables(self, device):
        """fp32 copies of the float64 tables (the reference gathers then .float(): :338), uploaded once."""
        key = str(device)
        if key not in self._dev_tables:
            step = np.stack([self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod,
                             self.posterior_mean_coef1, self.posterior_mean_coef2,
                             self.posterior_log_variance_clipped]).astype(np.float32)
            q = np.stack([self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod]).astype(np.float32)
            self._dev_tables[key] = (torch.from_numpy(step).to(device).contiguous(),
                                     torch.from_numpy(q).to(device).contiguous())
        return self._dev_tables[key]

    def _ddim_tables(self, device):
        """[4, T] fp32: sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, alphas_cumprod, alphas_cumprod_prev."""
        key = "ddim:" + str(device)
        if key not in self._dev_tables:
            tab = np.stack([self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod, self.alphas_cumprod,
                            self.alphas_cumprod_prev]).astype(np.float32)
            self._dev_tables[key] = torch.from_numpy(tab).to(device).contiguous()
        return self._dev_tables[key]

    def _check_supported(self):
        if self.model_mean_type != ModelMeanType.EPSILON or self.model_var_type != ModelVarType.FIXED_SMALL:
            raise NotImplementedError("the CUDA sampler implements the trainer's configuration only: "
                                      "EPSILON mean, FIXED_SMALL variance (trainers/ddpm_trainer.py:43-50)")
        if self.rescale_timesteps:
            raise NotImplementedError("rescale_timesteps=True is not used by the reference trainer")

    # ------------------------------------------------------------------ q(x_t | x_0)
    def q_sample(self, x_start, t, noise=None):
        """:449-460."""
        if noise is None:
            noise = torch.randn_like(x_start)
        assert noise.shape == x_start.shape
        _, q = self._tables(x_start.device)
        out = torch.empty_like(x_start, dtype=torch.float32)
        ops.q_sample(x_start.float().contiguous(), noise.float().contiguous(), t.to(torch.int64).contiguous(), q,
                     self.num_timesteps, out)
        return out

    # ------------------------------------------------------------------ p(x_{t-1} | x_t)
    def _extract(self, arr, t, shape):
        """_extract_into_tensor, :329-341 (gather, .float(), broadcast); plumbing on small tables."""
        res = torch.from_numpy(arr).to(device=t.device)[t].float()
        while res.dim() < len(shape):
            res = res[..., None]
        return res.expand(shape)

    def _p_mean_variance(self, model, x, t, clip_denoised, denoised_fn, model_kwargs, noise=None):
        self._check_supported()
        if denoised_fn is not None:
            raise NotImplementedError("denoised_fn is not supported by the fused update")
        if model_kwargs is None:
            model_kwargs = {}
        B = x.shape[0]
        assert t.shape == (B,)
        x = x.float().contiguous()
        t = t.to(torch.int64).contiguous()
        eps = model(x, t, **model_kwargs)                               # :493
        assert eps.shape == x.shape
        step, _ = self._tables(x.device)
        mean, x0 = torch.empty_like(x), torch.empty_like(x)
        sample = torch.empty_like(x) if noise is not None else None
        ops.p_mean_variance(x, eps.contiguous(), t, step, self.num_timesteps, clip_denoised, mean=mean, x0=x0,
                            noise=None if noise is None else noise.float().contiguous(), sample=sample)
        out = {"mean": mean, "variance": self._extract(self.posterior_variance, t, x.shape),
               "log_variance": self._extract(self.posterior_log_variance_clipped, t, x.shape), "pred_xstart": x0}
        if sample is not None:
            out["sample"] = sample
        return out

    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        """:481-552 for the trainer's configuration (EPSILON mean, FIXED_SMALL variance): one model forward,
        then x0 and the posterior mean in one kernel (mdm_p_mean_variance).  Returns the reference's dict."""
        return self._p_mean_variance(model, x, t, clip_denoised, denoised_fn, model_kwargs)

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                 noise_fn=None, noise=None):
        """:582-614.  The reference's default noise_fn=th.randn_like is called as noise_fn(shape, device=...,
        dtype=...) and raises TypeError; here noise_fn=None means torch.randn_like(x), a
        shape-style callable (e.g. torch.randn) is called the way the reference calls it, and `noise`
        (keyword-only extra) injects a tensor."""
        if cond_fn is not None:
            raise NotImplementedError("cond_fn (classifier guidance) is not used by the reference trainer")
        if noise is None:
            noise = torch.randn_like(x, dtype=torch.float32) if noise_fn is None else \
                noise_fn(x.shape, device=x.device, dtype=torch.float32)
        out = self._p_mean_variance(model, x, t, clip_denoised, denoised_fn, model_kwargs, noise=noise)
        return {"sample": out["sample"], "pred_xstart": out["pred_xstart"]}

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=False, noise_fn=None):
        """:616-697 (plain ancestral sampling, one forward per step, no guidance)."""
        if device is None:
            device = next(model.parameters()).device
        img = torch.randn(*shape, device=device) if noise is None else noise.to(device).float()
        steps = list(reversed(range(self.num_timesteps)))
        if progress:
            from tqdm.auto import tqdm
            steps = tqdm(steps)
        for i in steps:
            t = torch.full((shape[0],), i, dtype=torch.int64, device=device)
            img = self.p_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                                model_kwargs=model_kwargs, noise_fn=noise_fn)["sample"]
        return img

    # ------------------------------------------------------------------ CFG sampling
    def _cfg_inputs(self, model, B, model_kwargs, device):
        """Conditional + unconditional text batched as 2B sequences with per-sequence token counts.
        The unconditional branch is text="" with xf_* re-encoded (:1059-1062)."""
        if model_kwargs is None or "text" not in model_kwargs:
            raise KeyError("text")                                     # reference: model_kwargs["text"], :1060
        xf_proj, xf_out = model_kwargs.get("xf_proj"), model_kwargs.get("xf_out")
        if xf_proj is None or xf_out is None:
            xf_proj, xf_out = model.encode_text(model_kwargs["text"], device)
        u_proj, u_out = model.encode_text([""] * len(model_kwargs["text"]), device)
        nc, nu = xf_out.shape[1], u_out.shape[1]
        nmax = max(nc, nu)
        Dt = xf_out.shape[2]
        xo = torch.zeros(2 * B, nmax, Dt, device=device, dtype=torch.float32)
        xo[:B, :nc] = xf_out
        xo[B:, :nu] = u_out
        xp = torch.cat([xf_proj.float(), u_proj.float()])
        nt = torch.cat([torch.full((B,), nc, dtype=torch.int32, device=device),
                        torch.full((B,), nu, dtype=torch.int32, device=device)])
        return model.prepare_text(xp, xo, nt)

    def _cfg_step(self, model, ctx, x, t, length2, noise, cfg_scale, clip, x_out, x0_out):
        B = x.shape[0]
        eps = model(torch.cat([x, x]), torch.cat([t, t]), length2, text_ctx=ctx)
        step, _ = self._tables(x.device)
        ops.cfg_update(x, eps[:B], eps[B:], noise, t, step, self.num_timesteps, cfg_scale, clip, x_out, x0_out)

    def make_cfg_stepper(self, model, shape, model_kwargs, cfg_scale=7.5, clip_denoised=True, device=None,
                         use_cuda_graph=True):
        """Static-shape CFG step runner: text context prepared once, device buffers allocated once and
        (optionally) the whole step (batched cond+uncond forward + update) captured in a CUDA graph."""
        return CFGStepper(self, model, shape, model_kwargs, cfg_scale, clip_denoised, device, use_cuda_graph)

    def p_sample_with_cfg(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None,
                          cfg_scale=7.5, noise=None, text_ctx=None):
        """:1042-1098.  `noise` (optional) replaces the internal torch.randn_like(x) draw."""
        self._check_supported()
        if denoised_fn is not None:
            raise NotImplementedError("denoised_fn is not supported by the fused CFG update")
        if model_kwargs is None:
            model_kwargs = {}
        B = x.shape[0]
        x = x.float().contiguous()
        t = t.to(torch.int64).contiguous()
        assert t.shape == (B,)
        ctx = text_ctx if text_ctx is not None else self._cfg_inputs(model, B, model_kwargs, x.device)
        length = model_kwargs["length"].reshape(-1).to(torch.int64)
        if noise is None:
            noise = torch.randn_like(x)                                # :1094
        sample, x0 = torch.empty_like(x), torch.empty_like(x)
        self._cfg_step(model, ctx, x, t, torch.cat([length, length]).contiguous(), noise.float().contiguous(),
                       cfg_scale, clip_denoised, sample, x0)
        return {"sample": sample, "pred_xstart": x0}

    def p_sample_loop_with_cfg(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                               model_kwargs=None, device=None, progress=False, cfg_scale=7.5, *,
                               use_cuda_graph=True, num_steps=None, step_noise=None):
        """:1100-1141.  Extra keyword-only arguments: use_cuda_graph (replay one captured step),
        num_steps (run only the first num_steps of the num_timesteps reverse steps: benchmarking),
        step_noise (callable t -> noise tensor, for parity tests with injected noise)."""
        self._check_supported()
        if denoised_fn is not None:
            raise NotImplementedError("denoised_fn is not supported by the fused CFG update")
        if device is None:
            device = next(model.parameters()).device
        st = CFGStepper(self, model, shape, model_kwargs, cfg_scale, clip_denoised, device, use_cuda_graph)
        st.x.copy_(torch.randn(*shape, device=device) if noise is None else noise.to(device).float())
        steps = list(reversed(range(self.num_timesteps)))
        if num_steps is not None:
            steps = steps[:num_steps]
        if progress:
            from tqdm.auto import tqdm
            steps = tqdm(steps, desc="Sampling")
        for ts in steps:
            st.step(ts, None if step_noise is None else step_noise(ts))
        out = st.x.clone()
        model.check_health()             # expert parallelism: barrier time-outs / overflows are raised here (synchronises)
        return out

    # ------------------------------------------------------------------ DDIM
    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                    eta=0.0, *, noise=None):
        """:699-742: one model forward, then pred_xstart / eps / sigma / sample in one kernel (mdm_ddim_update,
        torch-eager arithmetic order).  Like the reference, a randn_like(x) is drawn even when eta == 0 (same RNG
        consumption); `noise` (keyword-only extra) injects the tensor instead."""
        self._check_supported()
        if denoised_fn is not None or cond_fn is not None:
            raise NotImplementedError("denoised_fn / cond_fn are not used by the reference's samplers")
        if model_kwargs is None:
            model_kwargs = {}
        B = x.shape[0]
        assert t.shape == (B,)
        x = x.float().contiguous()
        t = t.to(torch.int64).contiguous()
        eps = model(x, t, **model_kwargs)                               # p_mean_variance, :493
        assert eps.shape == x.shape
        if noise is None:
            noise = torch.randn_like(x)                                 # :735
        sample, x0 = torch.empty_like(x), torch.empty_like(x)
        ops.ddim_update(x, eps.contiguous(), t, self._ddim_tables(x.device), self.num_timesteps, float(eta),
                        clip_denoised, sample, noise=noise.float().contiguous(), x0=x0)
        return {"sample": sample, "pred_xstart": x0}

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                     cond_fn=None, model_kwargs=None, device=None, progress=False, eta=0.0):
        """:777-815 (all num_timesteps steps, t = T-1 .. 0)."""
        if device is None:
            device = next(model.parameters()).device
        img = torch.randn(*shape, device=device) if noise is None else noise.to(device).float()
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        for i in indices:
            t = torch.full((shape[0],), i, dtype=torch.int64, device=device)
            out = self.ddim_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                   cond_fn=cond_fn, model_kwargs=model_kwargs, eta=eta)
            yield out
            img = out["sample"]

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0):
        """:744-775."""
        final = None
        for sample in self.ddim_sample_loop_progressive(model, shape, noise=noise, clip_denoised=clip_denoised,
                                                        denoised_fn=denoised_fn, cond_fn=cond_fn,
                                                        model_kwargs=model_kwargs, device=device, progress=progress,
                                                        eta=eta):
            final = sample
        return final["sample"]

    def ddim_timesteps(self, num_inference_steps):
        """Strided schedule for DDIM with fewer steps: `num_inference_steps` timesteps spread uniformly over
        [0, T), returned in sampling order (descending) together with each one's predecessor (-1 after the last)."""
        S = int(num_inference_steps)
        if not 1 <= S <= self.num_timesteps:
            raise ValueError("num_inference_steps must be in [1, %d]" % self.num_timesteps)
        ts = sorted(set(int(round(i * self.num_timesteps / S)) for i in range(S)))
        order = ts[::-1]
        return order, order[1:] + [-1]

    def _cfg_ddim_step(self, model, ctx, x, t, t_prev, length2, noise, cfg_scale, eta, clip, x_out, x0_out):
        B = x.shape[0]
        eps = model(torch.cat([x, x]), torch.cat([t, t]), length2, text_ctx=ctx)
        ops.ddim_update(x, eps[:B], t, self._ddim_tables(x.device), self.num_timesteps, float(eta), clip, x_out,
                        eps_u=eps[B:], cfg_scale=cfg_scale, noise=noise, t_prev=t_prev, x0=x0_out)

    def ddim_sample_with_cfg(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None,
                             cfg_scale=7.5, eta=0.0, *, t_prev=None, noise=None, text_ctx=None):
        """The DDIM update (:721-742) on the classifier-free-guided pred_xstart of p_sample_with_cfg (:1065-1079):
        cond + uncond as one batched forward.  t_prev [B] (optional): previous timestep of a strided schedule."""
        self._check_supported()
        if denoised_fn is not None:
            raise NotImplementedError("denoised_fn is not supported by the fused update")
        B = x.shape[0]
        x = x.float().contiguous()
        t = t.to(torch.int64).contiguous()
        ctx = text_ctx if text_ctx is not None else self._cfg_inputs(model, B, model_kwargs, x.device)
        length = model_kwargs["length"].reshape(-1).to(torch.int64)
        if noise is None and eta != 0.0:
            noise = torch.randn_like(x)
        sample, x0 = torch.empty_like(x), torch.empty_like(x)
        self._cfg_ddim_step(model, ctx, x, t, None if t_prev is None else t_prev.to(torch.int64).contiguous(),
                            torch.cat([length, length]).contiguous(),
                            None if noise is None else noise.float().contiguous(), cfg_scale, eta, clip_denoised,
                            sample, x0)
        return {"sample": sample, "pred_xstart": x0}

    def ddim_sample_loop_with_cfg(self, model, shape, noise=None, clip_denoised=True, model_kwargs=None, device=None,
                                  progress=False, cfg_scale=7.5, eta=0.0, num_inference_steps=50, *,
                                  use_cuda_graph=True):
        """CFG sampling in `num_inference_steps` DDIM steps (strided schedule, `ddim_timesteps`) on the same
        CUDA-graph step runner as p_sample_loop_with_cfg: 1000 / num_inference_steps times fewer forwards."""
        self._check_supported()
        if device is None:
            device = next(model.parameters()).device
        st = CFGStepper(self, model, shape, model_kwargs, cfg_scale, clip_denoised, device, use_cuda_graph,
                        sampler="ddim", eta=eta)
        st.x.copy_(torch.randn(*shape, device=device) if noise is None else noise.to(device).float())
        order, prev = self.ddim_timesteps(num_inference_steps)
        steps = list(zip(order, prev))
        if progress:
            from tqdm.auto import tqdm
            steps = tqdm(steps, desc="DDIM sampling")
        for ts, tp in steps:
            st.step(ts, ts_prev=tp)
        out = st.x.clone()
        model.check_health()
        return out

    # ------------------------------------------------------------------ training losses
    def training_losses(self, model, x_start, t, model_kwargs=None, noise=None):
        """:923-992, MSE branch: {"mse", "target", "pred", "moe_loss"}.
        The reference back-propagates through these tensors with torch autograd; here the backward is hand-written: when the
        model has a training engine attached (training.TrainEngine, created by DDPMTrainer with is_train=True) the forward
        keeps its activations and the returned dict carries them under the extra key "saved", to be handed to
        `engine.backward(terms["saved"], dLoss/dpred)` (DDPMTrainer.update does).  Without an engine: forward values only."""
        self._check_supported()
        if self.loss_type not in (LossType.MSE, LossType.RESCALED_MSE):
            raise NotImplementedError(self.loss_type)
        if model_kwargs is None:
            model_kwargs = {}
        if noise is None:
            noise = torch.randn_like(x_start)
        x_t = self.q_sample(x_start, t, noise=noise)
        model.reset_all_moe_counters(model)                           # :935
        eng = getattr(model, "_train_engine", None)
        saved = None
        if eng is not None:
            xf_proj, xf_out = model_kwargs.get("xf_proj"), model_kwargs.get("xf_out")
            if xf_proj is None or xf_out is None:                     # transformer.py:311-312
                xf_proj, xf_out = model.encode_text(model_kwargs["text"], x_start.device)
            with torch.cuda.device(eng.dev):
                out, saved = eng.forward_train(x_t, t, model_kwargs["length"], xf_proj, xf_out)
        else:
            out = model(x_t, t, **model_kwargs)                       # :950
        assert out.shape == noise.shape == x_start.shape
        terms = {"mse": ((noise - out) ** 2).mean(dim=list(range(1, out.dim()))).view(-1),
                 "target": noise, "pred": out, "moe_loss": 0.0 + model.get_moe_loss(model)}
        if saved is not None:
            terms["saved"] = saved
        return terms


class CFGStepper:
    """One reverse-diffusion step x_t -> x_{t-1} with classifier-free guidance on static device buffers.

    `x` is updated in place by `step(t)`.  With use_cuda_graph the step is captured once (after one
    eager warm-up step on scratch data, so lazily created workspaces exist) and replayed afterwards."""

    def __init__(self, diffusion, model, shape, model_kwargs, cfg_scale=7.5, clip_denoised=True, device=None,
                 use_cuda_graph=True, sampler="ddpm", eta=0.0):
        diffusion._check_supported()
        if sampler not in ("ddpm", "ddim"):
            raise ValueError(sampler)
        self.sampler, self.eta = sampler, float(eta)
        self.d, self.model = diffusion, model
        self.device = torch.device(device if device is not None else next(model.parameters()).device)
        if self.device.type != "cuda":
            raise MdmError("CFGStepper needs a CUDA device: there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.B = shape[0]
        self.cfg_scale, self.clip = cfg_scale, clip_denoised
        self.ctx = diffusion._cfg_inputs(model, self.B, model_kwargs, self.device)
        length = model_kwargs["length"].reshape(-1).to(device=self.device, dtype=torch.int64)
        self.length2 = torch.cat([length, length]).contiguous()
        # `length` is fixed over the loop: sort once, FastAttention starts the longest sequences first
        self.ctx.seq_order = torch.argsort(self.length2, descending=True, stable=True).to(torch.int32).contiguous()
        self.x = torch.zeros(*shape, device=self.device, dtype=torch.float32)
        self.x0 = torch.zeros_like(self.x)
        self.noise = torch.zeros_like(self.x)
        self.t = torch.zeros(self.B, dtype=torch.int64, device=self.device)
        self.t_prev = torch.zeros(self.B, dtype=torch.int64, device=self.device)
        self.use_graph = use_cuda_graph
        self.graph = None

    def _run(self):
        with torch.cuda.device(self.device):
            self._run_on_device()

    def _run_on_device(self):
        if self.sampler == "ddim":
            self.d._cfg_ddim_step(self.model, self.ctx, self.x, self.t, self.t_prev, self.length2, self.noise,
                                  self.cfg_scale, self.eta, self.clip, self.x, self.x0)
            return
        self.d._cfg_step(self.model, self.ctx, self.x, self.t, self.length2, self.noise, self.cfg_scale, self.clip,
                         self.x, self.x0)

    def _capture(self):
        with torch.cuda.device(self.device):
            self._capture_on_device()

    def _capture_on_device(self):
        pk = self.model._packed or self.model._pack()
        keep = (self.x.clone(), pk["usage"].clone(), pk["importance"].clone())
        self._run()                                   # warm-up: allocates workspaces, sets kernel attributes
        torch.cuda.synchronize(self.device)
        self.model.check_health()                     # expert parallelism: a barrier time-out / overflow is fatal here
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._run()
        self.x.copy_(keep[0])
        pk["usage"].copy_(keep[1])
        pk["importance"].copy_(keep[2])

    def step(self, ts, noise=None, ts_prev=None):
        """Advance x from timestep ts to ts-1 (DDIM: to ts_prev, default ts-1).  noise=None draws torch's
        normal_ (== randn_like, :1094); a DDIM step with eta == 0 uses no noise."""
        with torch.cuda.device(self.device):
            return self._step(ts, noise, ts_prev)

    def _step(self, ts, noise, ts_prev):
        if self.use_graph and self.graph is None:
            self._capture()
        self.t.fill_(int(ts))
        self.t_prev.fill_(int(ts) - 1 if ts_prev is None else int(ts_prev))
        if noise is not None:
            self.noise.copy_(noise)
        elif self.sampler == "ddpm" or self.eta != 0.0:
            self.noise.normal_()
        if self.use_graph:
            self.graph.replay()
        else:
            self._run()
        return self.x

    # ------------------------------------------------------------------ host-buffer interface (pipelined copies)
    def step_host(self, x_host, ts, out_host, noise=None, ts_prev=None):
        """One step on HOST data: x_host (pinned, [B,T,F] fp32) -> device, step(ts), result -> out_host (pinned).
        The copies run on two side streams through double-buffered device staging tensors, so the upload of the next
        call and the download of the previous result overlap the compute of the current step (PCIe and the SMs work
        concurrently; a call returns as soon as its work is enqueued).  Call flush() before reading out_host.
        x_host is read asynchronously: it must not be modified until `input_consumed()` has returned (or flush()),
        which waits for the upload of the most recent call."""
        if not (x_host.is_pinned() and out_host.is_pinned()):
            raise ValueError("step_host needs pinned host tensors (the copies are asynchronous)")
        with torch.cuda.device(self.device):
            self._step_host(x_host, ts, out_host, noise, ts_prev)

    def _step_host(self, x_host, ts, out_host, noise, ts_prev):
        if not hasattr(self, "_hp"):
            mk = lambda: [torch.empty_like(self.x) for _ in range(2)]
            ev = lambda: [torch.cuda.Event() for _ in range(2)]
            self._hp = {"xin": mk(), "xout": mk(), "s_in": torch.cuda.Stream(self.device), "s_out": torch.cuda.Stream(self.device),
                        "in_ready": ev(), "in_free": ev(), "out_ready": ev(), "out_free": ev(), "i": 0}
            if self.use_graph and self.graph is None:
                self._capture()
        hp = self._hp
        k = hp["i"] & 1
        first = hp["i"] < 2
        hp["i"] += 1
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(hp["s_in"]):
            if not first:
                hp["s_in"].wait_event(hp["in_free"][k])          # the step two calls ago has consumed xin[k]
            hp["xin"][k].copy_(x_host, non_blocking=True)
            hp["in_ready"][k].record(hp["s_in"])
        main.wait_event(hp["in_ready"][k])
        self.x.copy_(hp["xin"][k])
        hp["in_free"][k].record(main)
        self.step(ts, noise=noise, ts_prev=ts_prev)
        if not first:
            main.wait_event(hp["out_free"][k])                    # the download two calls ago has left xout[k]
        hp["xout"][k].copy_(self.x)
        hp["out_ready"][k].record(main)
        with torch.cuda.stream(hp["s_out"]):
            hp["s_out"].wait_event(hp["out_ready"][k])
            out_host.copy_(hp["xout"][k], non_blocking=True)
            hp["out_free"][k].record(hp["s_out"])

    def input_consumed(self):
        """Block the host until the uploads of every step_host() call so far have read their x_host."""
        hp = getattr(self, "_hp", None)
        if hp is not None:
            for k in range(min(2, hp["i"])):
                hp["in_ready"][k].synchronize()

    def flush(self):
        """Make the current stream wait for every download issued by step_host()."""
        hp = getattr(self, "_hp", None)
        if hp is not None:
            main = torch.cuda.current_stream(self.device)
            for k in range(min(2, hp["i"])):
                main.wait_event(hp["out_free"][k])

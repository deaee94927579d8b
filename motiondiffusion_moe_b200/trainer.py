"""The inference surface of the reference's DDPMTrainer (trainers/ddpm_trainer.py) over the B200 kernels: what
tools/evaluation.py and tools/visualization.py call to turn captions into motions.

    trainer = DDPMTrainer(opt, encoder)        # opt.device, opt.diffusion_steps, opt.is_train (False), opt.cfg_scale
    trainer.load(path)                         # the reference's ckpt_e###.tar          (:277-289)
    motions = trainer.generate(captions, m_lens, dim_pose, batch_size=8)               (:176-199)

`generate_batch` (:145-174) is `p_sample_loop_with_cfg` on the CUDA-graph step runner.  Extras (keyword-only /
optional attributes, absent from the reference): `sampler="ddim"` with `num_inference_steps` runs the strided DDIM
loop instead of the 1000-step ancestral one.  `forward` / `backward_G` (:97-143, 201-225) compute the training-step
VALUES (noise prediction, masked reconstruction loss, MoE balance loss) on the kernels; `update` / `train` (:227-360)
are not built: there are no backward / optimizer kernels yet, and they say so loudly."""
from collections import OrderedDict

import numpy as np
import torch

from .gaussian_diffusion import GaussianDiffusion, LossType, ModelMeanType, ModelVarType, get_named_beta_schedule
from .postprocess import load_reference_checkpoint, save_reference_checkpoint


class DDPMTrainer(object):
    def __init__(self, args, encoder, *, sampler="ddpm", num_inference_steps=50, eta=0.0):
        self.opt = args
        self.device = args.device
        self.encoder = encoder
        self.diffusion_steps = args.diffusion_steps
        betas = get_named_beta_schedule("linear", self.diffusion_steps)                    # :41-43
        self.diffusion = GaussianDiffusion(betas=betas, model_mean_type=ModelMeanType.EPSILON,
                                           model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
        self.sampler_name = "uniform"
        if getattr(args, "is_train", False):
            raise NotImplementedError("the DDPM training step (backward + Adam) has no CUDA kernels yet: "
                                      "construct the trainer with is_train=False for sampling")
        if sampler not in ("ddpm", "ddim"):
            raise ValueError(sampler)
        self.sampling, self.num_inference_steps, self.eta = sampler, num_inference_steps, eta
        self.to(self.device)
        self.cfg_scale = getattr(args, "cfg_scale", 7.5)                                    # :67

    def _model(self):                                                                      # :69-77
        return self.encoder.module if hasattr(self.encoder, "module") else self.encoder

    def maybe_reset_all_moe_counters(self):                                               # :79-85
        m = self._model()
        if hasattr(m, "reset_all_moe_counters"):
            m.reset_all_moe_counters(m)

    def to(self, device):                                                                  # :246-251
        self._model().to(device)

    def train_mode(self):
        self._model().train()

    def eval_mode(self):                                                                   # :257-258
        self._model().eval()

    def save(self, file_name, ep, total_it):                                               # :260-275
        save_reference_checkpoint(self._model(), file_name, ep=ep, total_it=total_it)

    def load(self, model_dir):                                                             # :277-289
        ep, total_it, _, _ = load_reference_checkpoint(self._model(), model_dir, map_location="cpu")
        self._model().to(self.device)
        return ep, total_it

    def generate_batch(self, caption, m_lens, dim_pose):                                   # :145-174
        m = self._model()
        xf_proj, xf_out = m.encode_text(caption, self.device) if hasattr(m, "encode_text") else (None, None)
        m_lens = torch.as_tensor(m_lens).to(self.device)
        T = int(min(int(m_lens.max()), m.num_frames)) if hasattr(m, "num_frames") else int(m_lens.max())
        B = len(caption)
        kw = {"xf_proj": xf_proj, "xf_out": xf_out, "length": m_lens, "text": caption}
        if self.sampling == "ddim":
            return self.diffusion.ddim_sample_loop_with_cfg(m, (B, T, dim_pose), clip_denoised=False, model_kwargs=kw,
                                                            cfg_scale=self.cfg_scale, eta=self.eta, device=self.device,
                                                            num_inference_steps=self.num_inference_steps)
        return self.diffusion.p_sample_loop_with_cfg(m, (B, T, dim_pose), clip_denoised=False, progress=False,
                                                     model_kwargs=kw, cfg_scale=self.cfg_scale, device=self.device)

    def generate(self, caption, m_lens, dim_pose, batch_size=8):                           # :176-199
        N = len(caption)
        cur_idx = 0
        self.eval_mode()
        all_output = []
        while cur_idx < N:
            end_idx = min(cur_idx + batch_size, N)
            output = self.generate_batch(caption[cur_idx:end_idx], m_lens[cur_idx:end_idx], dim_pose)
            for i in range(output.shape[0]):
                all_output.append(output[i])
            cur_idx += batch_size
        return all_output

    # ------------------------------------------------------------------ training forward (loss VALUES; no backward yet)
    def forward(self, batch_data, eval_mode=False):                                       # :97-143
        """q_sample at a uniformly drawn timestep, one model forward, noise prediction vs noise, MoE balance loss, mask:
        the values the reference computes before `update` back-propagates.  Uses numpy's global RNG for the timesteps
        (UniformSampler.sample, models/gaussian_diffusion.py:108-133) and torch's for the noise, like the reference."""
        caption, motions, m_lens = batch_data
        motions = motions.detach().to(self.device).float()
        self.caption, self.motions = caption, motions
        B, T = motions.shape[:2]
        cur_len = torch.LongTensor([min(T, int(m)) for m in m_lens]).to(self.device)
        p = np.ones([self.diffusion.num_timesteps]) / self.diffusion.num_timesteps
        t = torch.from_numpy(np.random.choice(len(p), size=(B,), p=p)).long().to(self.device)
        output = self.diffusion.training_losses(model=self._model(), x_start=motions, t=t,
                                                model_kwargs={"text": caption, "length": cur_len})
        self.real_noise, self.fake_noise = output["target"], output["pred"]
        self.moe_loss = output.get("moe_loss", 0.0)
        self.cur_len = cur_len
        self.src_mask = self._model().generate_src_mask(T, cur_len).to(motions.device)

    def backward_G(self):                                                                 # :201-225 (values only)
        """loss_mot_rec = sum(mse_per_frame * src_mask) / sum(src_mask) (one kernel: mdm_masked_mse) + moe_loss."""
        pred, target = self.fake_noise.float().contiguous(), self.real_noise.float().contiguous()
        dev = pred.device
        if not hasattr(self, "_loss_ws") or self._loss_ws[0].numel() < pred.shape[0]:
            self._loss_ws = (torch.zeros(pred.shape[0], device=dev), torch.zeros(1, dtype=torch.int32, device=dev),
                             torch.zeros(1, device=dev))
        partial, counter, loss = self._loss_ws
        from . import ops
        ops.masked_mse(pred, target, self.cur_len.contiguous(), partial, counter, loss)
        loss_mot_rec = loss[0].clone()
        total = loss_mot_rec + (self.moe_loss if isinstance(self.moe_loss, torch.Tensor) else 0.0)
        self.loss_mot_rec = total
        return OrderedDict({"loss_mot_rec": loss_mot_rec.item(),
                            "loss_moe": float(self.moe_loss) if isinstance(self.moe_loss, torch.Tensor) else self.moe_loss,
                            "loss_total": total.item()})

    # ------------------------------------------------------------------ optimisation: not built
    def _no_training(self, *a, **kw):
        raise NotImplementedError("DDPMTrainer.update / train need the backward and optimizer kernels of the training "
                                  "step, which are not built (DESIGN.md section 7)")

    update = train = _no_training

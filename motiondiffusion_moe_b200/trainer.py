"""The inference surface of the reference's DDPMTrainer (trainers/ddpm_trainer.py) over the B200 kernels: what
tools/evaluation.py and tools/visualization.py call to turn captions into motions.

    trainer = DDPMTrainer(opt, encoder)        # opt.device, opt.diffusion_steps, opt.is_train (False), opt.cfg_scale
    trainer.load(path)                         # the reference's ckpt_e###.tar          (:277-289)
    motions = trainer.generate(captions, m_lens, dim_pose, batch_size=8)               (:176-199)

`generate_batch` (:145-174) is `p_sample_loop_with_cfg` on the CUDA-graph step runner.  Extras (keyword-only /
optional attributes, absent from the reference): `sampler="ddim"` with `num_inference_steps` runs the strided DDIM
loop instead of the 1000-step ancestral one.

Training (args.is_train): `forward` / `backward_G` / `update` / `train` (:97-143, 201-244, 290-360) run on the hand-written
backward and optimizer kernels through `training.TrainEngine` (flat fp32 master parameters + bf16 operand mirror, fused
global-norm clipping + Adam).  With torch.distributed initialised and world size > 1, `update` all-reduces the flat
gradient buffer (average) before the optimizer step: the data-parallel semantics the reference intends with its DDP
wrapper, whose reducer never fires because the trainer calls the unwrapped module."""
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
        if sampler not in ("ddpm", "ddim"):
            raise ValueError(sampler)
        self.sampling, self.num_inference_steps, self.eta = sampler, num_inference_steps, eta
        self.to(self.device)
        self.cfg_scale = getattr(args, "cfg_scale", 7.5)                                    # :67
        self.engine = None
        if getattr(args, "is_train", False):                                                # :52-53, :297 (Adam, lr = opt.lr)
            from .training import TrainEngine
            self.engine = TrainEngine(self._model(), lr=getattr(args, "lr", 2e-4), max_grad_norm=1.0)

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
        opt_state = None
        if self.engine is not None:     # Adam moments in the flat layout of TrainEngine (the reference stores torch's opt.state_dict())
            opt_state = {"mdm_b200_flat_adam": True, "step": self.engine.step_count, "offset": dict(self.engine.offset),
                         "exp_avg": self.engine.exp_avg.cpu(), "exp_avg_sq": self.engine.exp_avg_sq.cpu()}
        save_reference_checkpoint(self._model(), file_name, ep=ep, total_it=total_it, opt_state=opt_state)

    def load(self, model_dir):                                                             # :277-289
        import torch as _t
        ckpt = _t.load(model_dir, map_location="cpu", weights_only=False) if not isinstance(model_dir, dict) else model_dir
        ep, total_it, _, _ = load_reference_checkpoint(self._model(), ckpt, map_location="cpu")
        if self.engine is None:
            self._model().to(self.device)
        else:                            # parameters are views of the engine's flat buffer: loaded in place; re-derive the mirrors
            opt = ckpt.get("opt_encoder") or {}
            if opt.get("mdm_b200_flat_adam") and opt.get("offset") == self.engine.offset:
                self.engine.exp_avg.copy_(opt["exp_avg"])
                self.engine.exp_avg_sq.copy_(opt["exp_avg_sq"])
                self.engine.step_count = int(opt["step"])
            self.engine.refresh()
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

    # ------------------------------------------------------------------ training step
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
        # training_losses (gaussian_diffusion.py:923-992); with a training engine attached the forward keeps its activations
        # for the backward kernels (output["saved"])
        output = self.diffusion.training_losses(model=self._model(), x_start=motions, t=t,
                                                model_kwargs={"text": caption, "length": cur_len})
        self.real_noise, self.fake_noise = output["target"], output["pred"]
        self.moe_loss = output.get("moe_loss", 0.0)
        self._saved = output.get("saved")
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

    # ------------------------------------------------------------------ optimisation
    def update(self):                                                                     # :227-244
        """zero_grad, backward_G, backward (hand-written kernels), clip_grad_norm_(1.0), Adam step."""
        if self.engine is None:
            raise RuntimeError("DDPMTrainer.update needs a trainer constructed with args.is_train = True")
        if getattr(self, "_saved", None) is None:
            raise RuntimeError("call forward(batch_data) before update()")
        from . import train_ops as T
        eng = self.engine
        with torch.cuda.device(eng.dev):
            eng.zero_grad()
            loss_logs = self.backward_G()
            d_pred = T.masked_mse_grad(self.fake_noise.float().contiguous(), self.real_noise.float().contiguous(),
                                       self.cur_len.contiguous())
            # the MoE balance loss carries no gradient.  Data parallel: every decoder layer's gradient bucket is
            # all-reduced on a side stream as soon as its backward is done, overlapped with the layers still to come
            eng.backward(self._saved, d_pred, grad_ready=self._bucket_all_reduce if self._dp_world() > 1 else None)
            self._saved = None
            self._finish_all_reduce()
            eng.optimizer_step()
        return loss_logs

    def _dp_world(self):
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def _bucket_all_reduce(self, ranges):
        """Average the given ranges of the flat gradient buffer over the ranks, on a side stream, ordered after the kernels
        that produced them (bucketed gradient all-reduce overlapped with the backward; the intended semantics of the
        reference's DDP wrapper, tools/train.py:140-145)."""
        import torch.distributed as dist
        eng = self.engine
        if not hasattr(self, "_ar_stream"):
            self._ar_stream = torch.cuda.Stream(eng.dev)
            self._ar_avg = dist.get_backend() == "nccl"              # gloo (tests on a shared GPU) has no AVG
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(eng.dev))
        self._ar_stream.wait_event(ready)
        with torch.cuda.stream(self._ar_stream):
            for lo, hi in ranges:
                buf = eng.grad[lo:hi]
                if self._ar_avg:
                    dist.all_reduce(buf, op=dist.ReduceOp.AVG)
                else:
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
                    buf.mul_(1.0 / dist.get_world_size())

    def _finish_all_reduce(self):
        if hasattr(self, "_ar_stream") and self._dp_world() > 1:
            torch.cuda.current_stream(self.engine.dev).wait_stream(self._ar_stream)

    def _all_reduce_gradients(self):
        """Data-parallel training (BASELINE.json configs[4] at N > 1): average the flat gradient buffer over the ranks.
        One collective over one contiguous buffer (NCCL reduces in-switch over NVLS on an NVSwitch node)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.engine.grad, op=dist.ReduceOp.SUM)
            self.engine.grad.mul_(1.0 / dist.get_world_size())

    def train(self, train_dataset):                                                       # :290-360
        """Main loop of the reference: per batch one update with the captions and one with empty captions (the
        unconditional branch of classifier-free guidance); MoE counters reset per epoch; checkpoints as the reference."""
        from os.path import join as pjoin
        import os
        it, cur_epoch = 0, 0
        latest = pjoin(self.opt.model_dir, "latest.tar")
        if os.path.exists(latest):
            cur_epoch, it = self.load(latest)
            self.engine.refresh()
        logs = OrderedDict()
        for epoch in range(cur_epoch, self.opt.num_epochs):
            self.train_mode()
            self.maybe_reset_all_moe_counters()
            for i, batch_data in enumerate(train_dataset):
                self.forward(batch_data)
                loss_logs = self.update()
                caption, motions, m_lens = batch_data
                self.forward(([""] * len(caption), motions, m_lens))
                for k, v in self.update().items():
                    loss_logs["uncond_" + k] = v
                for k, v in loss_logs.items():
                    logs[k] = logs.get(k, 0.0) + v
                it += 1
                if it % getattr(self.opt, "log_every", 50) == 0:
                    print("epoch %d it %d " % (epoch, it) + " ".join("%s: %.4f" % (k, v / self.opt.log_every) for k, v in logs.items()))
                    logs = OrderedDict()
                if it % getattr(self.opt, "save_latest", 500) == 0:
                    self.save(latest, epoch, it)
            self.save(latest, epoch, it)
            if epoch % getattr(self.opt, "save_every_e", 5) == 0:
                self.save(pjoin(self.opt.model_dir, "ckpt_e%03d.tar" % epoch), epoch, total_it=it)

"""Continuous-time Gaussian diffusion sampler on the B200 plan (drop-in for
lidargen/models/diffusion/{base,continuous_time,continuous_time_cond}.py).

Public surface kept from the reference: constructor kwargs, ``sampling_shape``, ``device``, ``randn`` /
``randn_like`` (None | Generator | per-sample list), ``log_snr``, ``q_step_from_x_0``, ``q_step``,
``p_step`` and ``sample`` with the exact reference signatures; ``p_sample_loop`` is an alias of
``sample`` (BASELINE.json's wording).  ``forward`` / ``p_loss`` EVALUATE the training loss (no gradient:
the kernel plans have no backward; training itself is SURVEY section 8f rank 4, out of scope).

One denoiser step = the model's static kernel plan + one fused sampler-update kernel; ``sample``
captures that step in a CUDA graph and replays it ``num_steps`` times.
"""
from __future__ import annotations

import math
from functools import partial
from typing import List, Literal

import torch
from torch import nn

from . import _lib


# ---- log-SNR schedules (continuous_time.py:14-63) --------------------------------------------
def _log(t, eps=1e-20):
    return torch.log(t.clamp(min=eps))


def _log_snr_schedule_linear(t: torch.Tensor) -> torch.Tensor:
    return -_log(torch.special.expm1(1e-4 + 10 * (t ** 2)))[:, None, None, None]


def _log_snr_schedule_cosine(t: torch.Tensor, logsnr_min: float = -15, logsnr_max: float = 15) -> torch.Tensor:
    t_min = math.atan(math.exp(-0.5 * logsnr_max))
    t_max = math.atan(math.exp(-0.5 * logsnr_min))
    return -2 * _log(torch.tan(t_min + t * (t_max - t_min)))[:, None, None, None]


def _log_snr_schedule_cosine_shifted(t, image_d, noise_d, logsnr_min=-15, logsnr_max=15):
    return _log_snr_schedule_cosine(t, logsnr_min, logsnr_max) + 2 * math.log(noise_d / image_d)


def _log_snr_schedule_cosine_interpolated(t, image_d, noise_d_low, noise_d_high, logsnr_min=-15, logsnr_max=15):
    lo = _log_snr_schedule_cosine_shifted(t, image_d, noise_d_low, logsnr_min, logsnr_max)
    hi = _log_snr_schedule_cosine_shifted(t, image_d, noise_d_high, logsnr_min, logsnr_max)
    return t * lo + (1 - t) * hi


def _log_snr_to_alpha_sigma(log_snr: torch.Tensor):
    return log_snr.sigmoid().sqrt(), (-log_snr).sigmoid().sqrt()


_OBJ = {"eps": 0, "v": 1, "x_0": 2}


class GaussianDiffusion(nn.Module):
    """diffusion/base.py:9-165 (sampling side)."""

    def __init__(self, model: nn.Module, condition_model: nn.Module = None, sampling: str = "ddpm",
                 prediction_type: str = "eps", loss_type="l2", num_training_steps: int | None = 1000,
                 noise_schedule: str = "linear", min_snr_loss_weight: bool = True, min_snr_gamma: float = 5.0,
                 sampling_resolution=None, clip_sample: bool = True, clip_sample_range: float = 1):
        super().__init__()
        self.model = model
        self.condition_model = condition_model
        self.sampling = sampling
        self.num_training_steps = num_training_steps
        self.objective = prediction_type
        self.noise_schedule = noise_schedule
        self.min_snr_loss_weight = min_snr_loss_weight
        self.min_snr_gamma = min_snr_gamma
        self.clip_sample = clip_sample
        self.clip_sample_range = clip_sample_range
        if prediction_type not in _OBJ:
            raise ValueError(f"invalid objective {prediction_type}")
        if not (isinstance(loss_type, nn.Module) or loss_type in ("l2", "l1", "huber")):
            raise ValueError(f"invalid criterion: {loss_type}")
        self.loss_type = loss_type
        if sampling_resolution is None:
            assert hasattr(self.model, "resolution") and hasattr(self.model, "in_channels")
            self.sampling_shape = (self.model.in_channels, *self.model.resolution)
        else:
            assert len(sampling_resolution) == 2 and hasattr(self.model, "in_channels")
            self.sampling_shape = (self.model.in_channels, *sampling_resolution)
        self.setup_parameters()
        self.register_buffer("_dummy", torch.tensor([]))

    @property
    def device(self):
        return self._dummy.device

    def randn(self, *shape, rng: List[torch.Generator] | torch.Generator | None = None, **kwargs) -> torch.Tensor:
        if rng is None:
            return torch.randn(*shape, **kwargs)
        elif isinstance(rng, torch.Generator):
            return torch.randn(*shape, generator=rng, **kwargs)
        elif isinstance(rng, list):
            assert len(rng) == shape[0]
            return torch.stack([torch.randn(*shape[1:], generator=r, **kwargs) for r in rng])
        raise ValueError(f"invalid rng: {rng}")

    def randn_like(self, x: torch.Tensor, rng=None) -> torch.Tensor:
        return self.randn(*x.shape, rng=rng, device=x.device, dtype=x.dtype)

    def setup_parameters(self) -> None:
        raise NotImplementedError

    # ---- loss EVALUATION (base.py:119-151): forward only -- the kernel plans have no backward, so this is the validation
    # loss of a checkpoint, not a training step (SURVEY 8f-4 stays out of scope) ----
    def _criterion(self, prediction: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if isinstance(self.loss_type, nn.Module):
            return self.loss_type(prediction, target)
        if self.loss_type == "l2":
            return (prediction - target) ** 2
        if self.loss_type == "l1":
            return (prediction - target).abs()
        return nn.functional.smooth_l1_loss(prediction, target, reduction="none")      # "huber" (base.py:45-46)

    @torch.inference_mode()
    def p_loss(self, x_0: torch.Tensor, steps: torch.Tensor, loss_mask: torch.Tensor | None = None) -> torch.Tensor:
        loss_mask = torch.ones_like(x_0) if loss_mask is None else loss_mask
        x_t, noise = self.q_step_from_x_0(x_0, steps)
        prediction = self.model(x_t, self.get_network_condition(steps))
        loss = self._criterion(prediction, self.get_target(x_0, steps, noise))       # (B,C,H,W)
        loss = (loss * loss_mask).flatten(1).sum(dim=1, keepdim=True)
        loss = loss / loss_mask.flatten(1).sum(dim=1, keepdim=True).add(1e-8)         # (B,1)
        return (loss * self.get_loss_weight(steps)).mean()

    @torch.inference_mode()
    def forward(self, x_0: torch.Tensor, loss_mask: torch.Tensor | None = None) -> torch.Tensor:
        steps = self.sample_timesteps(x_0.shape[0], x_0.device)
        return self.p_loss(x_0, steps, loss_mask)


class ContinuousTimeGaussianDiffusion(GaussianDiffusion):
    """diffusion/continuous_time.py:66-260."""

    def __init__(self, model: nn.Module, condition_model: nn.Module = None, prediction_type: str = "eps",
                 loss_type="l2", noise_schedule: str = "cosine", min_snr_loss_weight: bool = True,
                 min_snr_gamma: float = 5.0, sampling_resolution=None, clip_sample: bool = True,
                 clip_sample_range: float = 1, image_d: float = None, noise_d_low: float = None,
                 noise_d_high: float = None):
        self.image_d, self.noise_d_low, self.noise_d_high = image_d, noise_d_low, noise_d_high
        super().__init__(model=model, condition_model=condition_model, sampling="ddpm",
                         prediction_type=prediction_type, loss_type=loss_type, num_training_steps=None,
                         noise_schedule=noise_schedule, min_snr_loss_weight=min_snr_loss_weight,
                         min_snr_gamma=min_snr_gamma, sampling_resolution=sampling_resolution,
                         clip_sample=clip_sample, clip_sample_range=clip_sample_range)
        self.use_cuda_graph = True

    def setup_parameters(self) -> None:
        t_min, t_max = math.atan(math.exp(-0.5 * 15)), math.atan(math.exp(0.5 * 15))
        # (kind, t_min, t_span, shift_lo, shift_hi) of b200_sampler_coefficients -- the device mirror of self.log_snr
        if self.noise_schedule == "linear":
            self.log_snr = _log_snr_schedule_linear
            self._sched = (0, 0.0, 0.0, 0.0, 0.0)
        elif self.noise_schedule == "cosine":
            self.log_snr = _log_snr_schedule_cosine
            self._sched = (1, t_min, t_max - t_min, 0.0, 0.0)
        elif self.noise_schedule == "cosine_shifted":
            assert self.image_d is not None and self.noise_d_low is not None
            self.log_snr = partial(_log_snr_schedule_cosine_shifted, image_d=self.image_d, noise_d=self.noise_d_low)
            self._sched = (2, t_min, t_max - t_min, 2 * math.log(self.noise_d_low / self.image_d), 0.0)
        elif self.noise_schedule == "cosine_interpolated":
            assert self.image_d is not None and self.noise_d_low is not None and self.noise_d_high is not None
            self.log_snr = partial(_log_snr_schedule_cosine_interpolated, image_d=self.image_d,
                                   noise_d_low=self.noise_d_low, noise_d_high=self.noise_d_high)
            self._sched = (3, t_min, t_max - t_min, 2 * math.log(self.noise_d_low / self.image_d),
                           2 * math.log(self.noise_d_high / self.image_d))
        else:
            raise ValueError(f"invalid beta schedule: {self.noise_schedule}")

    def sample_timesteps(self, batch_size: int, device) -> torch.Tensor:
        return torch.rand(batch_size, device=device, dtype=torch.float32)

    def get_network_condition(self, steps):
        return self.log_snr(steps)[:, 0, 0, 0]

    def get_target(self, x_0, step_t, noise):
        """continuous_time.py:142-153"""
        if self.objective == "eps":
            return noise
        if self.objective == "x_0":
            return x_0
        alpha, sigma = _log_snr_to_alpha_sigma(self.log_snr(step_t))
        return alpha * noise - sigma * x_0

    def get_loss_weight(self, steps):
        """continuous_time.py:155-169 (min-SNR-gamma weighting); shape [B,1,1,1] like the reference"""
        snr = self.log_snr(steps).exp()
        clipped = snr.clamp(max=self.min_snr_gamma) if self.min_snr_loss_weight else snr
        if self.objective == "eps":
            return clipped / snr
        if self.objective == "x_0":
            return clipped
        return clipped / (snr + 1)

    def q_step_from_x_0(self, x_0, step_t, rng=None):
        noise = self.randn_like(x_0, rng=rng)
        alpha, sigma = _log_snr_to_alpha_sigma(self.log_snr(step_t))
        return x_0 * alpha + noise * sigma, noise

    def q_step(self, x_s, step_t, step_s, rng=None):
        alpha_t, sigma_t = _log_snr_to_alpha_sigma(self.log_snr(step_t))
        alpha_s, sigma_s = _log_snr_to_alpha_sigma(self.log_snr(step_s))
        alpha_ts = alpha_t / alpha_s
        var_noise = self.randn_like(x_s, rng=rng)
        var = sigma_t.pow(2) - alpha_ts.pow(2) * sigma_s.pow(2)
        return x_s * alpha_ts + var.sqrt() * var_noise

    # ---- sampler coefficients: [.., 8] = alpha_t, sigma_t, alpha_s, sigma_s, c1, c2, ddpm_c, 0 ----
    def _coefficients(self, step_t: torch.Tensor, step_s: torch.Tensor, ddim_eta: float, out=None):
        """-> (log-SNR(t) [B], coefficient rows [B,8]); ``out`` = (lt, coef) device buffers to write into."""
        if step_t.is_cuda:
            # one launch instead of the ~25 elementwise kernels of the expressions below (same fp32 formulas)
            B = step_t.shape[0]
            st, ss = step_t.float().contiguous(), step_s.float().contiguous()
            if out is not None:
                lt, coef = out
            else:
                lt = torch.empty(B, device=st.device, dtype=torch.float32)
                coef = torch.empty(B, 8, device=st.device, dtype=torch.float32)
            k, t_min, t_span, sh_lo, sh_hi = self._sched
            _lib.get_lib().sampler_coefficients(st.data_ptr(), ss.data_ptr(), k, t_min, t_span, sh_lo, sh_hi,
                                                float(ddim_eta), lt.data_ptr(), coef.data_ptr(), B,
                                                _lib.current_stream(st.device))
            return lt, coef
        lt = self.log_snr(step_t)[:, 0, 0, 0]
        ls = self.log_snr(step_s)[:, 0, 0, 0]
        a_t, s_t = _log_snr_to_alpha_sigma(lt)
        a_s, s_s = _log_snr_to_alpha_sigma(ls)
        c1 = ddim_eta * s_s / s_t * (1 - a_t ** 2 / a_s ** 2).sqrt()
        c2 = (1 - a_s ** 2 - c1 ** 2).sqrt()
        cc = -torch.special.expm1(lt - ls)
        coef = torch.stack([a_t, s_t, a_s, s_s, c1, c2, cc, torch.zeros_like(cc)], dim=-1).float().contiguous()
        lt = lt.float().contiguous()
        if out is not None:           # (emulator / CPU path) honour the caller's buffers like the kernel does
            out[0].copy_(lt)
            out[1].copy_(coef)
            return out
        return lt, coef

    def _predict(self, x_t: torch.Tensor, log_snr_t: torch.Tensor) -> torch.Tensor:
        return self.model(x_t, log_snr_t)

    @torch.inference_mode()
    def p_step(self, x_t: torch.Tensor, step_t: torch.Tensor, step_s: torch.Tensor, rng=None,
               mode: Literal["ddpm", "ddim"] = "ddpm", ddim_eta: float = 0.0) -> torch.Tensor:
        """continuous_time.py:194-234: one reverse step (model forward + fused update kernel)."""
        if mode not in ("ddpm", "ddim"):
            raise ValueError(f"invalid mode {mode}")
        plan = self.model.get_plan(x_t.shape[0]) if hasattr(self.model, "get_plan") else None
        return self._p_step_plan(plan, x_t, step_t, step_s, rng, mode, ddim_eta)

    def _p_step_plan(self, plan, x_t, step_t, step_s, rng, mode, ddim_eta, predict=None):
        """one reverse step on a kernel plan whose condition (if any) is already folded"""
        if plan is not None and self.use_cuda_graph and x_t.is_cuda:
            # the same captured step (model plan + fused update) that sample() replays: one graph launch per call
            entry = self._step_graph(plan, x_t.shape[0], mode)
            if entry["graph"] is not None:
                plan.x_in.copy_(x_t)
                # straight into the graph's inputs (step tensors follow x_t's device: a CPU step tensor must not leave the
                # captured coefficient buffers stale)
                self._coefficients(step_t.to(x_t.device), step_s.to(x_t.device), ddim_eta, out=(plan.t_in, entry["coef"]))
                noise = self.randn_like(x_t, rng=rng)      # drawn every step like the reference (RNG stream parity)
                if mode == "ddpm" or ddim_eta != 0.0:
                    entry["noise"].copy_(noise)
                entry["graph"].replay()
                return plan.x_in.clone()
        step_t, step_s = step_t.to(x_t.device), step_s.to(x_t.device)
        lt, coef = self._coefficients(step_t, step_s, ddim_eta)
        pred = (plan(x_t, lt) if plan is not None else (predict or self._predict)(x_t, lt)).contiguous()
        noise = self.randn_like(x_t, rng=rng).contiguous()
        x_t = x_t.contiguous()
        x_s = torch.empty_like(x_t)
        lib = _lib.get_lib()
        B = x_t.shape[0]
        lib.sampler_update(x_t.data_ptr(), pred.data_ptr(), noise.data_ptr(), coef.data_ptr(), x_s.data_ptr(), B,
                           x_t[0].numel(), 0 if mode == "ddim" else 1, _OBJ[self.objective],
                           float(self.clip_sample_range) if self.clip_sample else 0.0,
                           _lib.current_stream(x_t.device))
        return x_s

    # ---- fast path: one CUDA graph per (batch, mode) replayed num_steps times ----
    def _step_graph(self, plan, B: int, mode: str):
        # the captured step lives ON the plan (it references the plan's buffers): a plan dropped by the model -- new
        # weights, coords, precision, .to() -- takes its graphs and activation arena with it instead of leaking them
        graphs = plan.__dict__.setdefault("_step_graphs", {})
        key = (B, mode, self.objective, self.clip_sample, self.clip_sample_range, tuple(self.sampling_shape),
               bool(self.use_cuda_graph))
        if key in graphs:
            return graphs[key]
        dev = plan.dev
        lib = _lib.get_lib()
        coef = torch.zeros(B, 8, device=dev)
        coef[:, 0] = 1; coef[:, 1] = 1; coef[:, 2] = 1; coef[:, 3] = 1
        noise = torch.zeros(B, *self.sampling_shape, device=dev)
        n_per = int(noise[0].numel())
        mode_i = 0 if mode == "ddim" else 1
        clip = float(self.clip_sample_range) if self.clip_sample else 0.0

        def step():
            st = _lib.current_stream(dev)
            plan.launch(st)
            # in-place: x_in <- update(x_in, pred)
            lib.sampler_update(plan.x_in.data_ptr(), plan.pred.data_ptr(), noise.data_ptr(), coef.data_ptr(),
                               plan.x_in.data_ptr(), B, n_per, mode_i, _OBJ[self.objective], clip, st)

        entry = {"coef": coef, "noise": noise, "step": step, "graph": None}
        if self.use_cuda_graph and dev.type == "cuda":
            x_keep = plan.x_in.clone()
            s = torch.cuda.Stream(dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                step()  # warm-up outside capture (function attributes, lazy module load)
            torch.cuda.current_stream(dev).wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            plan.x_in.copy_(x_keep)
            entry["graph"] = g
        graphs[key] = entry
        return entry

    @torch.inference_mode()
    def sample(self, batch_size: int, num_steps: int, progress: bool = True, rng=None, return_all: bool = False,
               mode: Literal["ddpm", "ddim"] = "ddpm", ddim_eta: float = 0.0):
        """continuous_time.py:236-260."""
        if mode not in ("ddpm", "ddim"):
            raise ValueError(f"invalid mode {mode}")
        x = self.randn(batch_size, *self.sampling_shape, rng=rng, device=self.device)
        return self._sample_from(x, num_steps, progress, rng, return_all, mode, ddim_eta)

    p_sample_loop = sample

    # ---- RePaint (continuous_time.py:262-319): same p_step kernels + q_step / mask blend ----
    def _repaint_loop(self, p_step_fn, known, mask, num_steps, num_resample_steps, jump_length, progress, rng,
                      return_all):
        assert num_resample_steps > 0 and jump_length > 0
        B = known.shape[0]
        x_t = self.randn(B, *self.sampling_shape, rng=rng, device=self.device)
        steps = torch.linspace(1, 0, num_steps + 1, device=self.device)[None].repeat_interleave(B, dim=0)
        out = [x_t] if return_all else None
        interp = torch.linspace(0, 1, jump_length + 1, device=self.device)
        x_s = x_t
        for i in range(num_steps):
            r_steps = steps[:, [i]] + interp[None] * (steps[:, [i + 1]] - steps[:, [i]])
            for j in range(num_resample_steps):
                x = x_t
                for k in range(jump_length):            # t -> s: reverse diffusion, known region re-noised
                    known_s, _ = self.q_step_from_x_0(known, r_steps[:, k + 1], rng=rng)
                    unknown_s = p_step_fn(x, r_steps[:, k], r_steps[:, k + 1], rng)
                    x = mask * known_s + (1 - mask) * unknown_s
                x_s = x
                if return_all:
                    out.append(x_s)
                if i == num_steps - 1 or j == num_resample_steps - 1:
                    x_t = x
                    break
                for k in range(jump_length, 0, -1):     # s -> t: forward diffusion (resampling jump)
                    x = self.q_step(x, r_steps[:, k - 1], r_steps[:, k], rng=rng)
                x_t = x
        return torch.stack(out) if return_all else x_s

    @torch.inference_mode()
    def repaint(self, known: torch.Tensor, mask: torch.Tensor, num_steps: int, num_resample_steps: int = 1,
                jump_length: int = 1, progress: bool = True, rng=None, return_all: bool = False):
        """continuous_time.py:262-319 (RePaint, https://arxiv.org/abs/2201.09865); reverse steps are DDPM p_steps."""
        return self._repaint_loop(lambda x, t, s, r: self.p_step(x, t, s, rng=r), known, mask, num_steps,
                                  num_resample_steps, jump_length, progress, rng, return_all)

    def _sample_from(self, x, num_steps, progress, rng, return_all, mode, ddim_eta, plan=None):
        B = x.shape[0]
        dev = x.device
        if plan is None:
            plan = self.model.get_plan(B)
        entry = self._step_graph(plan, B, mode)
        steps = torch.linspace(1.0, 0.0, num_steps + 1, device=dev)
        # per-step tables (same fp32 torch math as the reference, evaluated once for the whole trajectory)
        lts, coefs = self._coefficients(steps[:-1].repeat_interleave(B), steps[1:].repeat_interleave(B), ddim_eta)
        lts, coefs = lts.view(num_steps, B), coefs.view(num_steps, B, 8)
        need_noise = mode == "ddpm" or ddim_eta != 0.0
        plan.x_in.copy_(x)
        out = [x] if return_all else None
        it = range(num_steps)
        if progress:
            try:
                from tqdm.auto import tqdm
                it = tqdm(it, desc="sampling", leave=False)
            except Exception:  # pragma: no cover
                pass
        for i in it:
            plan.t_in.copy_(lts[i])
            entry["coef"].copy_(coefs[i])
            # the reference draws noise every step (even for eta == 0): keep the RNG stream identical
            noise = self.randn_like(plan.x_in, rng=rng)
            if need_noise:
                entry["noise"].copy_(noise)
            if entry["graph"] is not None:
                entry["graph"].replay()
            else:
                entry["step"]()
            if return_all:
                out.append(plan.x_in.clone())
        return torch.stack(out) if return_all else plan.x_in.clone()


class CondContinuousTimeGaussianDiffusion(ContinuousTimeGaussianDiffusion):
    """diffusion/continuous_time_cond.py:66-281: layout-conditioned sampling (cond_mode='concat').

    ``sample(batch_dict, batch_size, num_steps, ...)`` runs ``condition_model(batch_dict)`` once, folds the condition
    into the denoiser plan (LayoutUnetPlan.set_condition) and replays the captured step graph."""

    def __init__(self, model: nn.Module, condition_model: nn.Module = None, prediction_type: str = "eps",
                 loss_type="l2", noise_schedule: str = "cosine", min_snr_loss_weight: bool = True,
                 min_snr_gamma: float = 5.0, sampling_resolution=None, clip_sample: bool = True,
                 clip_sample_range: float = 1, image_d: float = None, noise_d_low: float = None,
                 noise_d_high: float = None, cond_mode: str = "concat", w_loss_weight: float = 1.0):
        super().__init__(model=model, condition_model=condition_model, prediction_type=prediction_type,
                         loss_type=loss_type, noise_schedule=noise_schedule, min_snr_loss_weight=min_snr_loss_weight,
                         min_snr_gamma=min_snr_gamma, sampling_resolution=sampling_resolution,
                         clip_sample=clip_sample, clip_sample_range=clip_sample_range, image_d=image_d,
                         noise_d_low=noise_d_low, noise_d_high=noise_d_high)
        self.cond_mode = cond_mode
        self.w_loss_weight = w_loss_weight
        if self.cond_mode == "concat":
            self.sampling_shape = (self.model.in_channels - condition_model.out_channels, *self.sampling_shape[1:])

    @torch.inference_mode()
    def p_loss(self, input_dict: dict, steps: torch.Tensor, loss_mask: torch.Tensor | None = None) -> torch.Tensor:
        """continuous_time_cond.py:414-436, EVALUATED (forward only, no gradient: the kernel plans have no backward)"""
        x_0 = input_dict["x_0"]
        loss_mask = torch.ones_like(x_0) if loss_mask is None else loss_mask
        x_t, noise = self.q_step_from_x_0(x_0, steps)
        condition = self.get_network_condition(steps, input_dict)
        self._reject_tensor_condition(condition["other_condition"])
        prediction = self.model(x_t, condition)
        loss = self._criterion(prediction, self.get_target(x_0, steps, noise))
        loss = (loss * loss_mask).flatten(1).sum(dim=1, keepdim=True)
        loss = loss / loss_mask.flatten(1).sum(dim=1, keepdim=True).add(1e-8)
        return (loss * self.get_loss_weight(steps)).mean()

    @torch.inference_mode()
    def forward(self, input_dict: dict) -> torch.Tensor:
        """continuous_time_cond.py:438-455"""
        x_0 = input_dict["x_0"]
        steps = self.sample_timesteps(x_0.shape[0], x_0.device)
        loss_mask = None
        if self.w_loss_weight:
            loss_mask = input_dict.get("scene_loss_weight_map", None)        # [B,H,W]
            if loss_mask is not None:
                loss_mask = loss_mask.unsqueeze(1).repeat(1, x_0.shape[1], 1, 1)
        return self.p_loss(input_dict, steps, loss_mask)

    def _reject_tensor_condition(self, other):
        if self.cond_mode == "concat" and isinstance(other, torch.Tensor):
            # continuous_time_cond.py:223-226 calls model(cat([x_t, cond]), {"time_condition"}) here; neither denoiser of the
            # hot path accepts that call in the reference either (LayoutUnetV1.forward reads cond_dict["other_condition"],
            # layout_unet_v1.py:869; EfficientUNet.forward takes a timestep tensor) -- it serves the out-of-scope HDiT models
            raise NotImplementedError("tensor-valued concat condition: wrap it as {'concat_cond': tensor, ...} "
                                      "(every nuScenes layout config passes the encoder's dict)")

    def get_network_condition(self, steps=None, input_dict=None, only_custom_condition=False):
        other_condition = self.condition_model(input_dict)
        if only_custom_condition:
            return dict(other_condition=other_condition)
        return dict(time_condition=self.log_snr(steps)[:, 0, 0, 0], other_condition=other_condition)

    @torch.inference_mode()
    def p_step(self, x_t: torch.Tensor, condition_dict: dict, step_t: torch.Tensor, step_s: torch.Tensor, rng=None,
               mode: Literal["ddpm", "ddim"] = "ddpm", ddim_eta: float = 0.0) -> torch.Tensor:
        """continuous_time_cond.py:206-253: folds the condition into the denoiser plan (every call: the dict may have
        changed), then runs the captured step (model plan + fused update kernel)."""
        if mode not in ("ddpm", "ddim"):
            raise ValueError(f"invalid mode {mode}")
        other = condition_dict["other_condition"]
        self._reject_tensor_condition(other)
        condition_dict.update(dict(time_condition=self.log_snr(step_t)[:, 0, 0, 0]))     # the reference mutates the dict
        plan = self.model.get_plan(x_t.shape[0])
        plan.set_condition(other)
        return self._p_step_plan(plan, x_t, step_t, step_s, rng, mode, ddim_eta)

    @torch.inference_mode()
    def sample(self, batch_dict: dict, batch_size: int, num_steps: int, progress: bool = True, rng=None,
               return_all: bool = False, mode: Literal["ddpm", "ddim"] = "ddpm", ddim_eta: float = 0.0):
        """continuous_time_cond.py:255-281."""
        if mode not in ("ddpm", "ddim"):
            raise ValueError(f"invalid mode {mode}")
        x = self.randn(batch_size, *self.sampling_shape, rng=rng, device=self.device)
        condition_dict = self.get_network_condition(input_dict=batch_dict, only_custom_condition=True)
        plan = self.model.get_plan(batch_size)
        plan.set_condition(condition_dict["other_condition"])
        return self._sample_from(x, num_steps, progress, rng, return_all, mode, ddim_eta, plan=plan)

    p_sample_loop = sample

    @torch.inference_mode()
    def inpaint(self, known: torch.Tensor, mask: torch.Tensor, batch_dict: dict, num_steps: int,
                num_resample_steps: int = 1, jump_length: int = 1, progress: bool = True, rng=None,
                return_all: bool = False):
        """continuous_time_cond.py:283-353: layout-conditioned RePaint.  (The reference passes ``condition_dict`` as
        ``step_t`` to ``q_step`` in the resampling branch, :337 -- only reached for num_resample_steps > 1; here
        q_step gets the step tensors it is declared with.)"""
        cond = self.get_network_condition(input_dict=batch_dict, only_custom_condition=True)
        self._reject_tensor_condition(cond["other_condition"])
        plan = self.model.get_plan(known.shape[0])
        plan.set_condition(cond["other_condition"])          # folded ONCE for the whole loop (not per reverse step)
        return self._repaint_loop(lambda x, t, s, r: self._p_step_plan(plan, x, t, s, r, "ddpm", 0.0), known, mask,
                                  num_steps, num_resample_steps, jump_length, progress, rng, return_all)

    @torch.inference_mode()
    def repaint(self, known, mask, batch_dict, num_steps, **kw):
        return self.inpaint(known, mask, batch_dict, num_steps, **kw)

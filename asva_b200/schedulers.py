"""Samplers for the denoising loop with the diffusers scheduler protocol the reference pipeline drives
(`init_noise_sigma`, `set_timesteps`, `timesteps`, `scale_model_input`, `step(...).prev_sample`, `config`;
/root/reference/avgen/pipelines/pipeline_audio_cond_animation.py:259,325-327,337,364).

  DDIMScheduler  eta = 0, epsilon prediction  (the sampler BASELINE.json's north_star measures)
  PNDMScheduler  PLMS with skip_prk_steps      (what scripts/animation_demo.py:75 instantiates)

Defaults are SD-1.5's scheduler_config.json (scaled-linear betas 0.00085..0.012, 1000 train steps, steps_offset 1,
set_alpha_to_one False, clip_sample False) - the same betas as configs/audio-cond_animation/*.yaml:8-20.
The arithmetic restates diffusers 0.29.2 (schedulers/scheduling_ddim.py, scheduling_pndm.py).

`step()` is the protocol entry point for foreign callers and runs as ordinary tensor ops on whatever device its
inputs live on.  The pipeline's hot loop does not call it: it asks `step_plan()` for per-step scalar coefficients
and feeds them to the fused CFG + sampler CUDA kernel (asva_cfg_ddim_step / asva_cfg_plms_step)."""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch


class _Cfg(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


def _alphas_cumprod(num_train_timesteps, beta_start, beta_end, beta_schedule) -> torch.Tensor:
    if beta_schedule == "scaled_linear":
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    elif beta_schedule == "linear":
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    else:
        raise NotImplementedError(beta_schedule)
    return torch.cumprod(1.0 - betas, dim=0)


@dataclass
class StepPlan:
    """Scalars of one sampler step for the fused kernel:  x <- c_sample * x + c_eps * e_hat,
    e_hat = a[0] * e + a[1] * hist[slots[1]] + a[2] * hist[slots[2]] + a[3] * hist[slots[3]];
    store e into hist[slots[0]] when slots[0] >= 0."""
    timestep: int
    c_sample: float
    c_eps: float
    a: tuple = (1.0, 0.0, 0.0, 0.0)
    slots: tuple = (-1, 0, 0, 0)


class _Base:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 set_alpha_to_one=False, steps_offset=1, prediction_type="epsilon", timestep_spacing="leading",
                 **extra):
        if prediction_type != "epsilon" or timestep_spacing != "leading":
            raise NotImplementedError("epsilon prediction with leading timestep spacing only")
        self.config = _Cfg(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                           beta_schedule=beta_schedule, set_alpha_to_one=set_alpha_to_one,
                           steps_offset=steps_offset, prediction_type=prediction_type,
                           timestep_spacing=timestep_spacing, **extra)
        self.alphas_cumprod = _alphas_cumprod(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    @classmethod
    def from_config(cls, config, **kw):
        c = {k: v for k, v in dict(config).items() if not k.startswith("_")}
        c.update(kw)
        return cls(**c)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _ac(self, t: int) -> float:
        return float(self.alphas_cumprod[t]) if t >= 0 else float(self.final_alpha_cumprod)

    def _leading(self, n: int) -> np.ndarray:
        ratio = self.config.num_train_timesteps // n
        return (np.arange(0, n) * ratio).round().astype(np.int64) + self.config.steps_offset


class DDIMScheduler(_Base):
    def __init__(self, clip_sample=False, **kw):
        if clip_sample:
            raise NotImplementedError("clip_sample=False only (SD-1.5)")
        super().__init__(clip_sample=clip_sample, **kw)

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.from_numpy(self._leading(num_inference_steps)[::-1].copy()).to(device)

    def _coefs(self, t: int):
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t, a_p = self._ac(t), self._ac(prev)
        # x0 = (x - sqrt(1-a_t) e)/sqrt(a_t);  x' = sqrt(a_p) x0 + sqrt(1-a_p) e
        c_sample = (a_p / a_t) ** 0.5
        c_eps = (1.0 - a_p) ** 0.5 - (a_p * (1.0 - a_t) / a_t) ** 0.5
        return c_sample, c_eps

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output=False,
             generator=None, variance_noise=None, return_dict: bool = True):
        if eta != 0.0:
            raise NotImplementedError("eta = 0 only")
        cs, ce = self._coefs(int(timestep))
        prev = cs * sample + ce * model_output
        return SchedulerOutput(prev) if return_dict else (prev,)

    def step_plan(self) -> List[StepPlan]:
        return [StepPlan(int(t), *self._coefs(int(t))) for t in self.timesteps.tolist()]


class PNDMScheduler(_Base):
    def __init__(self, skip_prk_steps=True, **kw):
        if not skip_prk_steps:
            raise NotImplementedError("skip_prk_steps=True only (SD-1.5)")
        super().__init__(skip_prk_steps=skip_prk_steps, **kw)
        self.ets, self.counter, self.cur_sample = [], 0, None

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        ts = self._leading(num_inference_steps)
        plms = np.concatenate([ts[:-1], ts[-2:-1], ts[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64)).to(device)
        self.ets, self.counter, self.cur_sample = [], 0, None

    def _coefs(self, t: int, prev: int):
        a_t, a_p = self._ac(t), self._ac(prev)
        c_sample = (a_p / a_t) ** 0.5
        denom = a_t * (1.0 - a_p) ** 0.5 + (a_t * (1.0 - a_t) * a_p) ** 0.5
        return c_sample, -(a_p - a_t) / denom

    def step(self, model_output, timestep, sample, return_dict: bool = True):
        """step_plms.  Like diffusers, `cur_sample` keeps a REFERENCE to `sample` at the first call; a caller that
        updates that storage in place (the reference pipeline does, :364) makes the second call start from the
        updated latents."""
        timestep = int(timestep)
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        prev = timestep - ratio
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        else:
            prev, timestep = timestep, timestep + ratio
        if len(self.ets) == 1 and self.counter == 0:
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            model_output = (model_output + self.ets[-1]) / 2
            sample, self.cur_sample = self.cur_sample, None
        elif len(self.ets) == 2:
            model_output = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            model_output = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            model_output = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3] - 9 * self.ets[-4])
        cs, ce = self._coefs(timestep, prev)
        self.counter += 1
        prev_sample = cs * sample + ce * model_output
        return SchedulerOutput(prev_sample) if return_dict else (prev_sample,)

    def step_plan(self) -> List[StepPlan]:
        """PLMS as coefficient rows over a 4-slot ring of past CFG-combined predictions.  The second call re-steps
        981 -> 961 with the averaged prediction FROM THE CURRENT LATENTS, which is what the reference pipeline's
        in-place latent update makes diffusers do."""
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        plans, n_hist, head = [], 0, 0  # head = ring slot of the newest stored prediction
        for i, t in enumerate(self.timesteps.tolist()):
            t = int(t)
            if i == 1:
                cs, ce = self._coefs(t + ratio, t)
                plans.append(StepPlan(t, cs, ce, (0.5, 0.5, 0.0, 0.0), (-1, head, 0, 0)))
                continue
            cs, ce = self._coefs(t, t - ratio)
            new = (head + 1) % 4 if n_hist else 0
            o1, o2, o3 = (new - 1) % 4, (new - 2) % 4, (new - 3) % 4  # previous predictions, newest first
            n_hist = min(n_hist + 1, 4)
            if n_hist == 1:
                a = (1.0, 0.0, 0.0, 0.0)
            elif n_hist == 2:
                a = (1.5, -0.5, 0.0, 0.0)
            elif n_hist == 3:
                a = (23 / 12, -16 / 12, 5 / 12, 0.0)
            else:
                a = (55 / 24, -59 / 24, 37 / 24, -9 / 24)
            plans.append(StepPlan(t, cs, ce, a, (new, o1, o2, o3)))
            head = new
        return plans


def plan_for(scheduler) -> Optional[List[StepPlan]]:
    """Step plans for our schedulers and for diffusers' own DDIMScheduler / PNDMScheduler objects (recognised by
    class name + config, so scripts/animation_demo.py's `PNDMScheduler.from_pretrained(...)` takes the fused path).
    None -> the pipeline falls back to calling scheduler.step() per step."""
    if isinstance(scheduler, (DDIMScheduler, PNDMScheduler)):
        return scheduler.step_plan()
    name = type(scheduler).__name__
    cfg = getattr(scheduler, "config", None)
    if cfg is None or name not in ("DDIMScheduler", "PNDMScheduler"):
        return None
    get = (lambda k, d=None: cfg.get(k, d)) if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
    if get("prediction_type", "epsilon") != "epsilon" or get("timestep_spacing", "leading") != "leading":
        return None
    common = dict(num_train_timesteps=get("num_train_timesteps", 1000), beta_start=get("beta_start"),
                  beta_end=get("beta_end"), beta_schedule=get("beta_schedule"),
                  set_alpha_to_one=get("set_alpha_to_one", False), steps_offset=get("steps_offset", 0))
    if get("trained_betas") is not None:
        return None
    if name == "DDIMScheduler":
        if get("clip_sample", False) or get("thresholding", False):
            return None
        mine = DDIMScheduler(**common)
    else:
        if not get("skip_prk_steps", False):
            return None
        mine = PNDMScheduler(**common)
    mine.set_timesteps(scheduler.num_inference_steps)
    if mine.timesteps.tolist() != [int(t) for t in scheduler.timesteps.tolist()]:
        return None
    return mine.step_plan()

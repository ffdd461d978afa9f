"""oracle/sampler_ref.py — TEST INFRASTRUCTURE.  CPU restatement of the sampler side of the hot path:
  * DDIM (eta 0) and PNDM/PLMS `step`, following diffusers==0.29.2 (requirements.txt:2) operation by operation in
    fp32 torch (schedulers/scheduling_ddim.py `step`, scheduling_pndm.py `step_plms` / `_get_prev_sample`);
  * the denoising loop of AudioCondAnimationPipeline.__call__
    (/root/reference/avgen/pipelines/pipeline_audio_cond_animation.py:325-365): k-fold latent duplication, UNet,
    CFG combine (:349-361), scheduler step on frames 1.. written back IN PLACE (:364).
PARITY UNPINNED for the scheduler arithmetic: diffusers is not installable here and the reference holds no vectors
for it; the restatement is checked only for self-consistency (DDIM closed form, PLMS order conditions)."""
import numpy as np
import torch


def alphas_cumprod(T=1000, beta_start=0.00085, beta_end=0.012):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, T, dtype=torch.float32) ** 2  # scaled_linear
    return torch.cumprod(1.0 - betas, dim=0)


def leading_timesteps(n, T=1000, offset=1):
    return (np.arange(0, n) * (T // n)).round().astype(np.int64) + offset


class DDIMRef:
    def __init__(self, n, T=1000):
        self.ac, self.n, self.T = alphas_cumprod(T), n, T
        self.final = self.ac[0]  # set_alpha_to_one = False
        self.timesteps = leading_timesteps(n, T)[::-1].copy()

    def step(self, eps, t, x):
        prev = t - self.T // self.n
        a_t = self.ac[t]
        a_p = self.ac[prev] if prev >= 0 else self.final
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        direction = (1 - a_p) ** 0.5 * eps  # eta = 0 -> sigma = 0
        return a_p ** 0.5 * x0 + direction


class PNDMRef:
    def __init__(self, n, T=1000):
        self.ac, self.n, self.T = alphas_cumprod(T), n, T
        self.final = self.ac[0]
        ts = leading_timesteps(n, T)
        self.timesteps = np.concatenate([ts[:-1], ts[-2:-1], ts[-1:]])[::-1].copy()  # skip_prk_steps
        self.ets, self.counter, self.cur_sample = [], 0, None

    def _prev_sample(self, x, t, prev, eps):
        a_t = self.ac[t]
        a_p = self.ac[prev] if prev >= 0 else self.final
        b_t, b_p = 1 - a_t, 1 - a_p
        coeff = (a_p / a_t) ** 0.5
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return coeff * x - (a_p - a_t) * eps / denom

    def step(self, eps, t, x):
        ratio = self.T // self.n
        prev = t - ratio
        if self.counter != 1:
            self.ets = self.ets[-3:] + [eps]
        else:
            prev, t = t, t + ratio
        e = self.ets
        if len(e) == 1 and self.counter == 0:
            self.cur_sample = x  # reference, not a copy: aliasing with the caller's storage is intended (F5)
        elif len(e) == 1 and self.counter == 1:
            eps = (eps + e[-1]) / 2
            x, self.cur_sample = self.cur_sample, None
        elif len(e) == 2:
            eps = (3 * e[-1] - e[-2]) / 2
        elif len(e) == 3:
            eps = (23 * e[-1] - 16 * e[-2] + 5 * e[-3]) / 12
        else:
            eps = (1 / 24) * (55 * e[-1] - 59 * e[-2] + 37 * e[-3] - 9 * e[-4])
        self.counter += 1
        return self._prev_sample(x, t, prev, eps)


def denoise_loop(unet, sched, latents, text, audio, mask, audio_scale=4.0, text_scale=1.0, trace=None):
    """unet(sample (k,4,F,h,w), t, text, audio, mask) -> eps.  latents (1,4,F,h,w) fp32, updated in place on frames
    1.. like the reference.  text/audio/mask are already k-fold CFG-batched by the caller (encode_text/encode_audio
    ordering: dual = [uncond, text, text+audio]; text-only = [audio, text+audio]; audio-only = [text, text+audio])."""
    do_t, do_a = text_scale > 1.0, audio_scale > 1.0
    k = 1 + int(do_t) + int(do_a)
    for i, t in enumerate(sched.timesteps.tolist()):
        eps = unet(torch.cat([latents] * k), t, text, audio, mask)
        if do_t and do_a:
            e_u, e_t, e_ta = eps.chunk(3)
            eps = e_u + text_scale * (e_t - e_u) + audio_scale * (e_ta - e_t)
        elif do_t:
            e_a, e_ta = eps.chunk(2)
            eps = e_a + text_scale * (e_ta - e_a)
        elif do_a:
            e_t, e_ta = eps.chunk(2)
            eps = e_t + audio_scale * (e_ta - e_t)
        latents[:, :, 1:] = sched.step(eps[:, :, 1:], t, latents[:, :, 1:])
        if trace is not None:
            trace.append(latents.clone())
    return latents

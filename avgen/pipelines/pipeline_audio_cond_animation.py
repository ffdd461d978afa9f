"""avgen.pipelines.pipeline_audio_cond_animation - drop-in for the reference module of the same path
(/root/reference/avgen/pipelines/pipeline_audio_cond_animation.py): AudioCondAnimationPipeline with the same
constructor / __call__ signature and return value, generate_videos, generate_videos_for_dataset.

Conditioning prep (CLIP text, ImageBind audio, VAE encode/decode) calls the stock torch modules it is given, as in
the reference.  The denoising loop (:330-365) is the hot path and runs as ONE CUDA-graph replay per step:
UNet forward for the k CFG branches (asva_b200.engine) + the fused CFG-combine/sampler kernel
(asva_cfg_ddim_step / asva_cfg_plms_step) updating frames 1.. of the fp32 latents in place."""
import inspect
import json
import os
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from asva_b200 import engine as _engine
from asva_b200 import schedulers as _sched
from avgen.models.unets import AudioUNet3DConditionModel  # noqa: F401  (re-exported like the reference module)


def _data_utils():
    """The reference's own I/O helpers (avgen/data/utils.py: torchvision video reader, torchaudio, ImageBind mel).
    They are outside this repo's scope and come from the reference checkout via avgen.__path__."""
    try:
        from avgen.data import utils as du
    except Exception as e:  # noqa: BLE001
        raise ImportError("avgen.data.utils (reference checkout + ImageBind submodule) is needed for file / waveform "
                          "inputs; pass mel-spectrogram tensors or use denoise() with precomputed contexts") from e
    return du


class _ProgressMixin:
    _progress_bar_config = {}

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    def progress_bar(self, iterable=None, total=None):
        from tqdm.auto import tqdm
        if iterable is not None:
            return tqdm(iterable, **self._progress_bar_config)
        return tqdm(total=total, **self._progress_bar_config)


class AudioCondAnimationPipeline(_ProgressMixin):
    def __init__(self, text_encoder, tokenizer, unet, scheduler, vae, audio_encoder,
                 null_text_encodings_path: str = ""):
        self.text_encoder, self.tokenizer, self.unet = text_encoder, tokenizer, unet
        self.scheduler, self.vae, self.audio_encoder = scheduler, vae, audio_encoder
        if null_text_encodings_path:
            self.null_text_encoding = torch.load(null_text_encodings_path).view(1, 77, 768)
        self.melspectrogram_shape = (128, 204)
        self.vae_scale_factor = 2 ** (len(self.vae.config.block_out_channels) - 1) if vae is not None else 8
        self._audio_processor = None
        self._loop = None  # cached graph runner + static buffers of the fused denoising step
        self.last_launches = 0

    # ------------------------------------------------------------------------------------------ plumbing
    @property
    def components(self):
        return dict(text_encoder=self.text_encoder, tokenizer=self.tokenizer, unet=self.unet,
                    scheduler=self.scheduler, vae=self.vae, audio_encoder=self.audio_encoder)

    def to(self, torch_device=None, dtype=None, **kw):
        for m in (self.text_encoder, self.unet, self.vae, self.audio_encoder):
            if isinstance(m, torch.nn.Module):
                m.to(device=torch_device, dtype=dtype)
        return self

    @property
    def device(self):
        return self.unet.device

    @property
    def dtype(self):
        return self.unet.dtype

    @property
    def audio_processor(self):
        if self._audio_processor is None:
            self._audio_processor = _data_utils().AudioMelspectrogramExtractor()
        return self._audio_processor

    # ------------------------------------------------------------------------------------------ conditioning
    @torch.no_grad()
    def encode_text(self, texts, device, dtype, do_text_classifier_free_guidance,
                    do_audio_classifier_free_guidance, text_encodings=None):
        """-> ((k b), 77, 768); CFG branch order as the reference (:149-154): dual [uncond, text, text],
        text-only [uncond, text], audio-only [text, text]."""
        if text_encodings is None:
            ids = self.tokenizer(texts, padding="max_length", max_length=self.tokenizer.model_max_length,
                                 truncation=True, return_tensors="pt")
            use_mask = getattr(self.text_encoder.config, "use_attention_mask", False)
            am = ids.attention_mask.to(device) if use_mask else None
            text_encodings = self.text_encoder(ids.input_ids.to(device), attention_mask=am)[0]
        elif isinstance(text_encodings, (list, tuple)):
            text_encodings = torch.cat(list(text_encodings))
        text_encodings = text_encodings.to(dtype=dtype, device=device)
        b = len(text_encodings)
        if do_text_classifier_free_guidance:
            if hasattr(self, "null_text_encoding"):
                uncond = self.null_text_encoding
            else:
                ids = self.tokenizer("", padding="max_length", max_length=text_encodings.shape[1], truncation=True,
                                     return_tensors="pt")
                use_mask = getattr(self.text_encoder.config, "use_attention_mask", False)
                am = ids.attention_mask.to(device) if use_mask else None
                uncond = self.text_encoder(ids.input_ids.to(device), attention_mask=am)[0]
            uncond = uncond.expand(b, -1, -1).contiguous().to(dtype=dtype, device=device)
            if do_audio_classifier_free_guidance:
                return torch.cat([uncond, text_encodings, text_encodings])
            return torch.cat([uncond, text_encodings])
        if do_audio_classifier_free_guidance:
            return torch.cat([text_encodings, text_encodings])
        return text_encodings

    @torch.no_grad()
    def encode_audio(self, audios, video_length: int = 12, do_text_classifier_free_guidance: bool = False,
                     do_audio_classifier_free_guidance: bool = False, device=torch.device("cuda:0"),
                     dtype=torch.float32):
        """-> audio encodings ((k b), f, 229, 768) (frame axis is a stride-0 expand: the engine projects keys/values
        once per clip) and boolean segment masks ((k b), f, 229).  Branch order (:186-194): dual [null, null, audio],
        text-only [audio, audio], audio-only [null, audio].  Unlike the reference the null masks are
        repeated to the batch size, so b > 1 also works."""
        b = len(audios)
        if torch.is_tensor(audios[0]) and audios[0].dim() == 3:  # already (1,128,204) mel-spectrograms
            mel = torch.stack(list(audios)).to(device=device, dtype=dtype)
        else:
            mel = self.audio_processor(audios).to(device=device, dtype=dtype)
        _, enc, masks = self.audio_encoder(mel, normalize=False, return_dict=False)
        if do_audio_classifier_free_guidance:
            null_mel = torch.zeros(1, 1, *self.melspectrogram_shape, device=device, dtype=dtype)
            _, null_enc, null_masks = self.audio_encoder(null_mel, normalize=False, return_dict=False)
            null_enc, null_masks = null_enc.expand(b, -1, -1), null_masks.expand(b, -1, -1)
            if do_text_classifier_free_guidance:
                enc, masks = torch.cat([null_enc, null_enc, enc]), torch.cat([null_masks, null_masks, masks])
            else:
                enc, masks = torch.cat([null_enc, enc]), torch.cat([null_masks, masks])
        elif do_text_classifier_free_guidance:
            enc, masks = torch.cat([enc, enc]), torch.cat([masks, masks])
        # the frame axis stays a stride-0 view (the reference materialises the repeat, :177): the engine sees at a
        # glance that the context is frame-invariant and projects keys / values once per clip
        return enc.unsqueeze(1).expand(-1, video_length, -1, -1), masks

    def _preprocess_images(self, images) -> torch.Tensor:
        """PIL / array / tensor -> (b,3,H,W) in [-1,1], H and W rounded down to a multiple of the VAE factor
        (diffusers VaeImageProcessor.preprocess defaults)."""
        if torch.is_tensor(images):
            return images
        out = []
        for im in images:
            if torch.is_tensor(im):
                out.append(im.float())
                continue
            import PIL.Image
            w, h = im.size
            w, h = w - w % self.vae_scale_factor, h - h % self.vae_scale_factor
            if (w, h) != im.size:
                im = im.resize((w, h), resample=PIL.Image.LANCZOS)
            a = np.asarray(im.convert("RGB"), dtype=np.float32) / 255.0
            out.append(torch.from_numpy(a).permute(2, 0, 1) * 2.0 - 1.0)
        return torch.stack(out)

    def _maybe_fast_vae(self, on_cuda: bool) -> None:
        """A diffusers AutoencoderKL on a CUDA device is run by the B200 engine (asva_b200.vae: same state dict, tcgen05
        convs, both encode and decode); ASVA_STOCK_VAE=1 keeps the module's own code."""
        if on_cuda and os.environ.get("ASVA_STOCK_VAE", "0") != "1":
            from asva_b200 import vae as _vae
            self.vae = _vae.wrap_vae(self.vae)

    @torch.no_grad()
    def encode_latents(self, image: torch.Tensor):
        self._maybe_fast_vae(torch.device(self.device).type == "cuda")
        image = image.to(device=self.device, dtype=self.vae.dtype)
        return self.vae.encode(image).latent_dist.sample() * self.vae.config.scaling_factor

    @torch.no_grad()
    def decode_latents(self, latents):
        """(b f) c h w latents -> images in [0, 1] on the CPU (:205-213)."""
        self._maybe_fast_vae(latents.is_cuda)
        latents = latents.to(dtype=next(self.vae.parameters()).dtype) / self.vae.config.scaling_factor
        image = self.vae.decode(latents).sample
        return (image / 2 + 0.5).clamp(0, 1).cpu().float()

    def prepare_extra_step_kwargs(self, generator, eta):
        params = set(inspect.signature(self.scheduler.step).parameters.keys())
        kw = {}
        if "eta" in params:
            kw["eta"] = eta
        if "generator" in params:
            kw["generator"] = generator
        return kw

    def prepare_video_latents(self, image_latents, num_channels_latents, video_length=12, height=256, width=256,
                              device=torch.device("cuda"), dtype=torch.float32, generator=None):
        """frame 0 = image latent, frames 1.. = N(0,1) noise, times init_noise_sigma (:234-261)."""
        b = len(image_latents)
        shape = (b, num_channels_latents, video_length - 1, height // self.vae_scale_factor,
                 width // self.vae_scale_factor)
        noise = torch.randn(shape, generator=generator, device=device, dtype=dtype)
        return torch.cat([image_latents.unsqueeze(2), noise], dim=2) * self.scheduler.init_noise_sigma

    # ------------------------------------------------------------------------------------------ hot loop
    @torch.no_grad()
    def open_session(self, text_encodings, audio_encodings, audio_masks, video_length, latent_h, latent_w,
                     num_inference_steps: int, audio_guidance_scale: float = 4.0,
                     text_guidance_scale: float = 1.0) -> Optional["DenoiseSession"]:
        """Binds one clip's conditioning and sampler schedule to the fused denoising step.  Returns None when the
        scheduler is not one the fused CFG+sampler kernel implements (the caller then uses scheduler.step())."""
        self.scheduler.set_timesteps(num_inference_steps, device=self.device)
        plans = _sched.plan_for(self.scheduler)
        if plans is None:
            return None
        return DenoiseSession(self, plans, text_encodings, audio_encodings, audio_masks, video_length, latent_h,
                              latent_w, audio_guidance_scale, text_guidance_scale)

    @torch.no_grad()
    def denoise(self, video_latents, text_encodings, audio_encodings, audio_masks, num_inference_steps: int,
                audio_guidance_scale: float = 4.0, text_guidance_scale: float = 1.0, generator=None,
                callback=None):
        """The denoising loop of the reference's __call__ (:325-365) on tensors: video_latents (b,4,F,h,w) with the
        conditioning frame at f = 0; contexts already CFG-batched ((k b),F,n,768) / masks ((k b),F,229).
        Returns the final latents (b,4,F,h,w) fp32 on the device."""
        do_t, do_a = text_guidance_scale > 1.0, audio_guidance_scale > 1.0
        k = 1 + int(do_t) + int(do_a)
        b, C, F, h, w = video_latents.shape
        assert text_encodings.shape[0] == k * b and audio_encodings.shape[0] == k * b, \
            (text_encodings.shape, audio_encodings.shape, k, b)
        sess = self.open_session(text_encodings, audio_encodings, audio_masks, F, h, w, num_inference_steps,
                                 audio_guidance_scale, text_guidance_scale)
        if sess is None:
            self.scheduler.set_timesteps(num_inference_steps, device=video_latents.device)
            return self._denoise_generic(video_latents, text_encodings, audio_encodings, audio_masks, k, do_t, do_a,
                                         audio_guidance_scale, text_guidance_scale, generator)
        sess.load_latents(video_latents)
        n0 = sess.launches
        for i in self.progress_bar(range(sess.num_steps)):
            sess.step(i)
            if callback is not None:
                callback(i, sess.plans[i].timestep, sess.latents)
        self.last_launches = sess.launches - n0
        return sess.latents.clone()

    def _denoise_generic(self, video_latents, text, audio, masks, k, do_t, do_a, s_a, s_t, generator):
        """Any other scheduler: the reference's loop in structure - our UNet module per step, CFG combine and
        scheduler.step() as tensor ops."""
        extra = self.prepare_extra_step_kwargs(generator, eta=0.0)
        video_latents = video_latents.clone()
        for t in self.progress_bar(self.scheduler.timesteps):
            x = self.scheduler.scale_model_input(torch.cat([video_latents] * k), t)
            e = self.unet(x, t, encoder_hidden_states=text, audio_encoder_hidden_states=audio,
                          audio_attention_mask=masks).sample
            if do_t and do_a:
                e_u, e_t, e_ta = e.chunk(3)
                e = e_u + s_t * (e_t - e_u) + s_a * (e_ta - e_t)
            elif do_t:
                e_a, e_ta = e.chunk(2)
                e = e_a + s_t * (e_ta - e_a)
            elif do_a:
                e_t, e_ta = e.chunk(2)
                e = e_t + s_a * (e_ta - e_t)
            video_latents[:, :, 1:] = self.scheduler.step(e[:, :, 1:], t, video_latents[:, :, 1:], **extra).prev_sample
        return video_latents

    @torch.no_grad()
    def __call__(self, images, audios, texts, text_encodings=None, video_length: int = 12, height: int = 256,
                 width: int = 256, num_inference_steps: int = 20, audio_guidance_scale: float = 4.0,
                 text_guidance_scale: float = 1.0, generator: Optional[torch.Generator] = None,
                 return_dict: bool = True):
        device, dtype = self.device, self.dtype
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        do_t, do_a = text_guidance_scale > 1.0, audio_guidance_scale > 1.0
        text_enc = self.encode_text(texts=texts, text_encodings=text_encodings, device=device, dtype=dtype,
                                    do_text_classifier_free_guidance=do_t, do_audio_classifier_free_guidance=do_a)
        text_enc = text_enc.unsqueeze(1).expand(-1, video_length, -1, -1)
        audio_enc, audio_masks = self.encode_audio(audios, video_length, do_t, do_a, device, dtype)
        image_latents = self.encode_latents(self._preprocess_images(images)).to(device=device, dtype=dtype)
        video_latents = self.prepare_video_latents(image_latents, self.unet.config.in_channels, video_length, height,
                                                   width, device=device, dtype=dtype, generator=generator)
        video_latents = self.denoise(video_latents.float(), text_enc, audio_enc, audio_masks, num_inference_steps,
                                     audio_guidance_scale, text_guidance_scale, generator).to(dtype)
        b = video_latents.shape[0]
        flat = video_latents.permute(0, 2, 1, 3, 4).reshape(b * video_length, *video_latents.shape[1:2],
                                                             *video_latents.shape[3:])
        videos = self.decode_latents(flat).detach().cpu()
        videos = videos.view(b, video_length, *videos.shape[1:])  # (b f c h w) in [0, 1]
        if not return_dict:
            return videos
        return {"videos": videos}


class DenoiseSession:
    """b clips on one GPU (b = 1 in the reference's own use; more clips per GPU batch along the token
    axis and fill the deep, small-M levels of the UNet): static fp32 latents (b,C,F,h,w) on the device, the per-step
    parameter table, and a CUDA-graph runner whose body is  UNet forward (k CFG branches x b clips, branch-major like
    the reference's torch.cat([latents] * k)) -> fused CFG combine + DDIM/PLMS update of frames 1.. in place.
    `step(i)` is one graph replay plus ONE small device-to-device copy of that step's parameter row
    (timesteps | CFG weights and sampler coefficients | PLMS history slots)."""

    def __init__(self, pipe, plans, text, audio, masks, F, h, w, audio_scale, text_scale):
        do_t, do_a = text_scale > 1.0, audio_scale > 1.0
        k = 1 + int(do_t) + int(do_a)
        s_t, s_a = float(text_scale), float(audio_scale)
        if do_t and do_a:  # e_u + s_t (e_t - e_u) + s_a (e_ta - e_t)   (reference :349-353)
            wts = (1.0 - s_t, s_t - s_a, s_a)
        elif do_t:
            wts = (1.0 - s_t, s_t, 0.0)
        elif do_a:
            wts = (1.0 - s_a, s_a, 0.0)
        else:
            wts = (1.0, 0.0, 0.0)
        unet = pipe.unet
        eng = unet.engine()
        dev = unet.device
        C, Co = unet.config.in_channels, unet.config.out_channels
        if C != Co:
            raise ValueError(f"the fused CFG + sampler step needs in_channels == out_channels (got {C}, {Co})")
        kb = text.shape[0]
        assert kb % k == 0 and audio.shape[0] == kb, \
            f"contexts must hold k = {k} CFG branches per clip, branch-major (got {tuple(text.shape)}, {tuple(audio.shape)})"
        b = kb // k
        self.plans, self.num_steps, self.k, self.clips, self.dev = plans, len(plans), k, b, dev
        self.plms = any(p.slots[0] >= 0 or tuple(p.a) != (1.0, 0.0, 0.0, 0.0) for p in plans)
        with torch.cuda.device(dev):
            if eng.shape != (kb, F, h, w):
                eng.prepare(kb, F, h, w)
                unet._runner, unet._ctx_key, pipe._loop = None, None, None
            unet.bind_context(text, audio, masks)
            # eng.gen changes whenever the engine (re)allocates buffers (prepare(), new context geometry), also when
            # that happens behind this pipeline's back (unet.forward with another batch, a second pipeline on the
            # same unet): a cached graph is only replayed against the buffers it was captured on
            key = (k, b, C, F, h, w, self.plms, id(eng), eng.ctx_sig, eng.gen)
            if pipe._loop is None or pipe._loop["key"] != key:
                par = torch.zeros(kb + 13, dtype=torch.float32, device=dev)  # [timesteps kb | coef 9 | slots 4 (int32)]
                L = dict(key=key, par=par, ts=par[:kb], coef=par[kb:kb + 9], slots=par[kb + 9:].view(torch.int32),
                         lat=torch.empty(b, C, F, h, w, dtype=torch.float32, device=dev),
                         eps=torch.empty(kb, Co, F, h, w, dtype=torch.float32, device=dev),
                         hist=torch.zeros(4, b, C, F, h, w, dtype=torch.float32, device=dev) if self.plms else None)
                be, plms = eng.be, self.plms

                def step_fn():
                    eng.forward(L["lat"], L["ts"], L["eps"])
                    if plms:
                        be.cfg_plms_step(L["eps"], k, L["lat"], L["hist"], L["coef"], L["slots"], C, F, h * w, b)
                    else:
                        be.cfg_ddim_step(L["eps"], k, L["lat"], L["coef"], C, F, h * w, b)

                L["runner"] = _engine.GraphRunner(step_fn, be)
                pipe._loop = L
            self.L = pipe._loop
            tab = torch.zeros(len(plans), kb + 13, dtype=torch.float32)
            tab_i = tab.view(torch.int32)  # the slot indices are int32 bit patterns in the same rows
            for i, p in enumerate(plans):
                tab[i, :kb] = float(p.timestep)
                tab[i, kb:kb + 9] = torch.tensor([*wts, p.c_sample, p.c_eps, *p.a], dtype=torch.float32)
                tab_i[i, kb + 9:] = torch.tensor(list(p.slots), dtype=torch.int32)
            self.table = tab.to(dev)

    @property
    def latents(self) -> torch.Tensor:
        return self.L["lat"]

    @property
    def launches(self) -> int:
        return self.L["runner"].total_launches

    def load_latents(self, latents: torch.Tensor) -> None:
        """Host (pinned) or device tensor (b,C,F,h,w) -> the session's static fp32 latents."""
        self.L["lat"].copy_(latents, non_blocking=True)

    def read_latents(self, out: torch.Tensor) -> None:
        out.copy_(self.L["lat"], non_blocking=True)

    def step(self, i: int) -> None:
        L = self.L
        with torch.cuda.device(self.dev):
            L["par"].copy_(self.table[i])
            L["runner"]()


@torch.no_grad()
def generate_videos(pipeline, image_path: str = "", audio_path: str = "", video_path: str = "", category: str = "",
                    category_text_encoding: Optional[torch.Tensor] = None, image_size: Tuple[int, int] = (256, 256),
                    video_fps: int = 6, video_num_frame: int = 12, num_clips_per_video: int = 3,
                    audio_guidance_scale: float = 4.0, text_guidance_scale: float = 1.0, seed: int = 0,
                    save_template: str = "", device: torch.device = torch.device("cuda")):
    """Same contract as the reference function (:379-468): loads the conditioning image / audio clips (reference I/O
    helpers), runs one pipeline call per clip with the generator re-seeded to `seed`, writes
    `<save_template>_clip-XX.mp4` (uint8 frames + 16 kHz aac audio) or returns (videos, audios)."""
    assert not (image_path and audio_path and video_path), \
        "Can not specify image_path, audio_path, video_path all three"
    du = _data_utils()
    import PIL.Image
    import torchvision
    clip_duration = video_num_frame / video_fps
    images = audios = None
    if image_path:
        images = [du.load_image(image_path, image_size)] * num_clips_per_video
    if audio_path:
        audios = du.load_audio_clips_uniformly(audio_path, clip_duration, num_clips_per_video,
                                               load_audio_as_melspectrogram=False)
    if video_path:
        vids, auds = du.load_av_clips_uniformly(video_path, video_fps, video_num_frame, image_size,
                                                num_clips_per_video, load_audio_as_melspectrogram=False)
        images = [v[0] for v in vids] if images is None else images
        audios = auds if audios is None else audios
    images = [PIL.Image.fromarray((255 * im).byte().permute(1, 2, 0).contiguous().numpy()) for im in images]
    out_v, out_a = [], []
    generator = torch.Generator(device=device)
    for k, (image, audio) in enumerate(zip(images, audios)):
        generator.manual_seed(seed)
        video = pipeline(images=[image], audios=[audio], texts=[category],
                         text_encodings=[category_text_encoding] if category_text_encoding is not None else None,
                         video_length=video_num_frame, height=image_size[0], width=image_size[1],
                         num_inference_steps=50, audio_guidance_scale=audio_guidance_scale,
                         text_guidance_scale=text_guidance_scale, generator=generator, return_dict=False)[0]
        video = (video.permute(0, 2, 3, 1).contiguous() * 255).byte()
        if save_template:
            path = f"{save_template}_clip-{k:02d}.mp4"
            os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
            if not hasattr(torchvision.io, "write_video"):
                raise ImportError("torchvision.io.write_video is gone from this torchvision (removed with the video "
                                  "IO deprecation); install the version the reference pins, or call generate_videos "
                                  "without save_template and write the returned uint8 frames yourself")
            torchvision.io.write_video(filename=path, video_array=video, fps=video_fps, audio_array=audio,
                                       audio_fps=16000, audio_codec="aac")
        else:
            out_v.append(video)
            out_a.append(audio)
    if save_template:
        return None
    return out_v, out_a


@torch.no_grad()
def generate_videos_for_dataset(exp_root: str, checkpoint: int, dataset: str = "AVSync15",
                                image_size: Tuple[int, int] = (256, 256), video_fps: int = 6,
                                video_num_frame: int = 12, num_clips_per_video: int = 3,
                                audio_guidance_scale: float = 4.0, text_guidance_scale: float = 1.0,
                                random_seed: int = 0, device: torch.device = torch.device("cuda"),
                                dtype: torch.dtype = torch.float16, rank: Optional[int] = None,
                                world_size: Optional[int] = None):
    """Same contract as the reference (:472-551) plus clip-level sharding: with RANK/WORLD_SIZE set (torchrun) or
    rank/world_size given, process r generates files r, r+W, r+2W, ... on its own GPU - clips are independent, no
    collective is needed on the data path."""
    from diffusers.models import AutoencoderKL
    from diffusers.schedulers import PNDMScheduler
    from tqdm import tqdm
    from transformers import CLIPTextModel, CLIPTokenizer

    from avgen.models.audio_encoders import ImageBindSegmaskAudioEncoder
    from avgen.utils import freeze_and_make_eval
    du = _data_utils()
    rank = int(os.environ.get("RANK", 0)) if rank is None else rank
    world_size = int(os.environ.get("WORLD_SIZE", 1)) if world_size is None else world_size
    if world_size > 1 and torch.device(device).index is None:
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    ckpt = f"{exp_root}/ckpts/checkpoint-{checkpoint}/modules"
    save_root = (f"{exp_root}/evaluations/checkpoint-{checkpoint}/AG-{audio_guidance_scale}_TG-{text_guidance_scale}"
                 f"/seed-{random_seed}/videos")
    video_root, filenames, categories, _ = du.get_evaluation_data(dataset)
    if dataset == "TheGreatestHits":
        enc = torch.load("./datasets/TheGreatestHits/class_clip_text_encodings_stable-diffusion-v1-5.pt",
                         map_location="cpu")
        cat_map, enc_map = {"hitting with a stick": "hitting with a stick"}, {"hitting with a stick": enc}
    elif dataset in ("Landscapes", "AVSync15"):
        cat_map = json.load(open(f"./datasets/{dataset}/class_mapping.json"))
        enc_map = torch.load(f"./datasets/{dataset}/class_clip_text_encodings_stable-diffusion-v1-5.pt",
                             map_location="cpu")
    else:
        raise ValueError(dataset)
    sd_path = "./pretrained/stable-diffusion-v1-5"
    pipeline = AudioCondAnimationPipeline(
        text_encoder=CLIPTextModel.from_pretrained(sd_path, subfolder="text_encoder").to(device=device, dtype=dtype),
        tokenizer=CLIPTokenizer.from_pretrained(sd_path, subfolder="tokenizer"),
        unet=AudioUNet3DConditionModel.from_pretrained(ckpt, subfolder="unet").to(device=device, dtype=dtype),
        scheduler=PNDMScheduler.from_pretrained(sd_path, subfolder="scheduler"),
        vae=AutoencoderKL.from_pretrained(sd_path, subfolder="vae").to(device=device, dtype=dtype),
        audio_encoder=freeze_and_make_eval(ImageBindSegmaskAudioEncoder(n_segment=video_num_frame).to(
            device=device, dtype=dtype)),
        null_text_encodings_path="./pretrained/openai-clip-l_null_text_encoding.pt")
    pipeline.to(torch_device=device, dtype=dtype)
    pipeline.set_progress_bar_config(disable=True)
    todo = list(zip(filenames, categories))[rank::world_size]
    for filename, category in tqdm(todo, total=len(todo), disable=rank != 0):
        generate_videos(pipeline, video_path=os.path.join(video_root, filename),
                        category_text_encoding=enc_map[cat_map[category]].view(1, 77, 768), image_size=image_size,
                        video_fps=video_fps, video_num_frame=video_num_frame,
                        num_clips_per_video=num_clips_per_video, text_guidance_scale=text_guidance_scale,
                        audio_guidance_scale=audio_guidance_scale, seed=random_seed,
                        save_template=os.path.join(save_root, filename.replace(".mp4", "")), device=device)

"""Stand-ins for the stock modules the reference pipeline is constructed with (scripts/animation_demo.py:74-92):
CLIP tokenizer / text encoder, SD-1.5 AutoencoderKL, ImageBindSegmaskAudioEncoder, diffusers' PNDMScheduler and the
reference's avgen.data.utils I/O helpers.  None of the real ones is installable here (no diffusers, no ImageBind, no
checkpoints), so the drop-in tests drive AudioCondAnimationPipeline.__call__ / generate_videos with objects that expose
exactly the attributes and call signatures the pipeline code touches
(/root/reference/avgen/pipelines/pipeline_audio_cond_animation.py:84-213, 264-375, 379-468) and compute cheap,
deterministic functions, so every stage's output can be predicted independently in the test.  TEST INFRASTRUCTURE."""
import types

import numpy as np
import torch

from asva_b200 import synth


class _Cfg(dict):
    __getattr__ = dict.get


class StubTokenizer:
    """CLIPTokenizer protocol: tokenizer(texts, padding=, max_length=, truncation=, return_tensors=) ->
    .input_ids / .attention_mask (b, 77) int64; .model_max_length."""
    model_max_length = 77

    def __call__(self, texts, padding="max_length", max_length=77, truncation=True, return_tensors="pt"):
        texts = [texts] if isinstance(texts, str) else list(texts)
        ids = torch.zeros(len(texts), max_length, dtype=torch.int64)
        for i, t in enumerate(texts):
            codes = [ord(c) % 997 + 1 for c in t][:max_length]
            ids[i, :len(codes)] = torch.tensor(codes, dtype=torch.int64)
        return types.SimpleNamespace(input_ids=ids, attention_mask=(ids != 0).long())


class StubTextEncoder(torch.nn.Module):
    """CLIPTextModel protocol: encoder(input_ids, attention_mask=None)[0] -> (b, 77, 768); .config.use_attention_mask."""

    def __init__(self, dim=768):
        super().__init__()
        self.config = _Cfg(use_attention_mask=False)
        g = torch.Generator().manual_seed(5)
        self.table = torch.nn.Parameter(torch.randn(998, dim, generator=g), requires_grad=False)

    def forward(self, input_ids, attention_mask=None):
        return (self.table[input_ids],)


class StubVAE(torch.nn.Module):
    """AutoencoderKL protocol: .config.block_out_channels / .scaling_factor, .dtype, encode(x).latent_dist.sample(),
    decode(z).sample.  encode = 8x8 average pooling of the RGB image into 4 channels (R, G, B, mean); decode =
    nearest 8x upsampling of the first three channels times `gain` (keeps random-weight latents inside the [0,1]
    clamp of decode_latents), so decode(encode(x)) is predictable."""

    gain = 0.02

    def __init__(self):
        super().__init__()
        self.config = _Cfg(block_out_channels=(128, 256, 512, 512), scaling_factor=0.18215, latent_channels=4)
        self.dummy = torch.nn.Parameter(torch.zeros(1), requires_grad=False)

    @property
    def dtype(self):
        return self.dummy.dtype

    def encode(self, x):
        p = torch.nn.functional.avg_pool2d(x.float(), 8)
        z = torch.cat([p, p.mean(dim=1, keepdim=True)], dim=1).to(x.dtype)
        return types.SimpleNamespace(latent_dist=types.SimpleNamespace(sample=lambda generator=None: z))

    def decode(self, z):
        img = (torch.nn.functional.interpolate(z[:, :3].float(), scale_factor=8, mode="nearest") * self.gain).to(z.dtype)
        return types.SimpleNamespace(sample=img)


class StubAudioEncoder(torch.nn.Module):
    """ImageBindSegmaskAudioEncoder protocol (segmask_imagebind.py:80-123): encoder(mel (b,1,128,204),
    normalize=False, return_dict=False) -> (cls (b,768), tokens (b,229,768), masks (b,S,229) bool) with the real
    segment-mask rule (synth.audio_segment_mask restates :62-78,104-114); tokens are a fixed random projection of the
    mel-spectrogram, so a zero mel gives the 'null audio' tokens."""

    def __init__(self, n_segment=12, dim=768):
        super().__init__()
        self.n_segment = n_segment
        g = torch.Generator().manual_seed(6)
        self.proj = torch.nn.Parameter(torch.randn(128 * 204 // 64, dim, generator=g) * 0.05, requires_grad=False)
        self.pos = torch.nn.Parameter(torch.randn(229, dim, generator=g), requires_grad=False)

    def forward(self, mel, normalize=False, return_dict=False):
        b = mel.shape[0]
        feat = mel.reshape(b, -1, 64).mean(-1) @ self.proj.to(mel.dtype)          # (b, 768)
        tokens = self.pos.to(mel.dtype)[None] + feat[:, None]                      # (b, 229, 768)
        tokens = torch.nn.functional.layer_norm(tokens, (tokens.shape[-1],))
        masks = synth.audio_segment_mask(self.n_segment).to(mel.device)[None].expand(b, -1, -1)
        return tokens[:, 0], tokens, masks


class StubMelExtractor:
    """avgen.data.utils.AudioMelspectrogramExtractor protocol: list of (c, t) waveforms @16 kHz -> (b,1,128,204)."""

    def __call__(self, audios):
        out = []
        for a in audios:
            a = torch.as_tensor(a).float().reshape(-1)
            n = 128 * 204
            a = torch.nn.functional.pad(a, (0, max(0, n - a.numel())))[:n]
            out.append(a.view(1, 128, 204))
        return torch.stack(out)


class PNDMScheduler:
    """Shaped like diffusers.PNDMScheduler (class name, .config keys of SD-1.5's scheduler_config.json,
    init_noise_sigma, set_timesteps, timesteps, scale_model_input, step(...).prev_sample) and NOT derived from the
    product's own scheduler classes: the arithmetic is the oracle's restatement (oracle/sampler_ref.py).  The pipeline
    must recognise it by name + config (schedulers.plan_for) and take the fused path."""
    init_noise_sigma = 1.0
    order = 1

    def __init__(self):
        self.config = _Cfg(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                           beta_schedule="scaled_linear", set_alpha_to_one=False, steps_offset=1,
                           skip_prk_steps=True, prediction_type="epsilon", timestep_spacing="leading",
                           trained_betas=None)
        self.num_inference_steps, self.timesteps, self._ref = None, None, None
        self.step_calls = 0

    def set_timesteps(self, num_inference_steps, device=None):
        from oracle import sampler_ref
        self._ref = sampler_ref.PNDMRef(num_inference_steps)
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.from_numpy(np.asarray(self._ref.timesteps)).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, return_dict=True):
        self.step_calls += 1
        prev = self._ref.step(model_output, int(timestep), sample)
        return types.SimpleNamespace(prev_sample=prev)


def data_utils_stub(n_samples=32000):
    """avgen.data.utils protocol used by generate_videos (:405-421): load_image -> (3,H,W) in [0,1];
    load_audio_clips_uniformly -> list of (1, t) waveforms; AudioMelspectrogramExtractor."""

    def load_image(path, image_size):
        g = torch.Generator().manual_seed(len(path))
        return torch.rand(3, image_size[0], image_size[1], generator=g)

    def load_audio_clips_uniformly(path, clip_duration, num_clips, load_audio_as_melspectrogram=False):
        g = torch.Generator().manual_seed(len(path) + 1)
        return [torch.randn(1, n_samples, generator=g) * 0.1 for _ in range(num_clips)]

    return types.SimpleNamespace(load_image=load_image, load_audio_clips_uniformly=load_audio_clips_uniformly,
                                 AudioMelspectrogramExtractor=StubMelExtractor)


class StubAutoencoderKL(torch.nn.Module):
    """Shaped like diffusers.AutoencoderKL for the calls the pipeline makes: a parameter tree whose state_dict() has the
    library's keys (`encoder.*`, `quant_conv.*`, `post_quant_conv.*`, `decoder.*`), `.config`, `.dtype`,
    `encode(x).latent_dist.sample()` and `decode(z).sample` computed by the restated VAE (oracle/vae_ref.py) on whatever device
    the parameters live on - it plays the stock module that asva_b200.vae.FastDecodeVAE wraps."""

    def __init__(self, block_out_channels=(128, 256, 512, 512), seed=7):
        super().__init__()
        from oracle import vae_ref
        self.cfg = dict(block_out_channels=tuple(block_out_channels))
        self.config = _Cfg(block_out_channels=tuple(block_out_channels), scaling_factor=0.18215, latent_channels=4,
                           layers_per_block=2, out_channels=3, norm_num_groups=32)
        sd = synth.synth_state_dict(vae_ref.state_dict_shapes(self.cfg) + vae_ref.encoder_state_dict_shapes(self.cfg),
                                    seed=seed)
        for key, val in sd.items():
            mod = self
            parts = key.split(".")
            for name in parts[:-1]:
                if not hasattr(mod, name):
                    mod.add_module(name, torch.nn.Module())
                mod = getattr(mod, name)
            mod.register_parameter(parts[-1], torch.nn.Parameter(val, requires_grad=False))

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def encode(self, x):
        from asva_b200.vae import DiagonalGaussian
        from oracle import vae_ref
        return types.SimpleNamespace(latent_dist=DiagonalGaussian(vae_ref.encode_moments(dict(self.state_dict()), x, self.cfg)))

    def decode(self, z, return_dict=True):
        from oracle import vae_ref
        return types.SimpleNamespace(sample=vae_ref.decode(dict(self.state_dict()), z, self.cfg))


def dataset_data_utils_stub(filenames, categories, n_samples=8000):
    """avgen.data.utils protocol used by generate_videos_for_dataset (:494-551): get_evaluation_data(dataset) ->
    (video_root, filenames, categories, _); load_av_clips_uniformly(...) -> (videos [(F,3,H,W) in [0,1]], audios)."""

    def get_evaluation_data(dataset):
        return "/videos/" + dataset, list(filenames), list(categories), None

    def load_av_clips_uniformly(video_path, video_fps, video_num_frame, image_size, num_clips,
                                load_audio_as_melspectrogram=False):
        g = torch.Generator().manual_seed(sum(map(ord, video_path)))
        vids = [torch.rand(video_num_frame, 3, image_size[0], image_size[1], generator=g) for _ in range(num_clips)]
        auds = [torch.randn(1, n_samples, generator=g) * 0.1 for _ in range(num_clips)]
        return vids, auds

    return types.SimpleNamespace(get_evaluation_data=get_evaluation_data, load_av_clips_uniformly=load_av_clips_uniformly,
                                 AudioMelspectrogramExtractor=StubMelExtractor)

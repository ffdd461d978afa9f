"""Deterministic synthetic weights and inputs for the denoising hot path (no checkpoints or datasets are
reachable from this environment; BASELINE.json asks for random-init weights + synthetic audio context).

Weights are a pure function of (key name, shape, seed) so the reference model, the CPU oracle and the CUDA engine
can all be given bit-identical parameters on any machine.  The zero-initialised temporal parameters of the
reference constructor (`*.conv_temp.*`, `*.attn_temp.to_out.0.weight`; utils.py:31-32,
ff_spatio_audio_temp_transformer_3d.py:267) are randomised like everything else, otherwise the temporal paths
would not be exercised."""
import math
import zlib
from typing import Dict, Iterable, Tuple

import numpy as np
import torch


def _gen(key: str, seed: int) -> torch.Generator:
    return torch.Generator(device="cpu").manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def synth_tensor(key: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    g = _gen(key, seed)
    leaf = key.rsplit(".", 1)[-1]
    is_norm = ".norm" in key or "conv_norm_out" in key or "group_norm" in key
    if leaf == "weight" and is_norm and len(shape) == 1:
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if leaf == "bias":
        return 0.02 * torch.randn(shape, generator=g)
    fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
    return torch.randn(shape, generator=g) / math.sqrt(fan_in)


def synth_state_dict(shapes: Iterable[Tuple[str, Tuple[int, ...]]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, tuple(s), seed) for k, s in shapes}


def audio_segment_mask(n_segment: int, n_freq: int = 12, n_time: int = 19) -> torch.Tensor:
    """Per-frame key mask over the 1 + 12*19 = 229 ImageBind audio tokens: frame s attends the CLS token and all
    12 frequency patches of ceil(19/S) consecutive time patches starting at round(linspace(0, 19-chunk, S))
    (the rule of ImageBindSegmaskAudioEncoder._auto_split / forward, segmask_imagebind.py:62-78,104-114).
    Returns bool (S, 229), True = attend."""
    chunk = int(math.ceil(n_time / n_segment))
    starts = np.round(np.linspace(0, n_time - chunk, n_segment, endpoint=True)).astype(np.int64)
    m = torch.zeros(n_segment, n_freq, n_time, dtype=torch.bool)
    for s, st in enumerate(starts):
        m[s, :, st:st + chunk] = True
    return torch.cat([torch.ones(n_segment, 1, dtype=torch.bool), m.reshape(n_segment, n_freq * n_time)], dim=1)


def synth_inputs(F: int = 12, h: int = 32, w: int = 32, seed: int = 123, k: int = 2, ctx_dim: int = 768,
                 n_text: int = 77):
    """One clip's synthetic conditioning in the pipeline's wire format (pipeline_audio_cond_animation.py:291-322):
    latents (1,4,F,h,w) fp32 (frame 0 stands in for the VAE latent of the conditioning image), and the CFG-batched
    (k = 2: [text only, text+audio]; k = 3: [uncond, text, text+audio] as encode_text / encode_audio order them,
    :149-154,186-194) text (k,F,77,768), audio (k,F,229,768) contexts and audio masks (k,F,229)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    lat = torch.randn(1, 4, F, h, w, generator=g)
    text = torch.randn(1, n_text, ctx_dim, generator=g)
    audio = torch.randn(1, 229, ctx_dim, generator=g)
    null_audio = torch.randn(1, 229, ctx_dim, generator=g)
    mask = audio_segment_mask(F)  # (F,229)
    if k == 1:
        return lat, text.unsqueeze(1).expand(1, F, -1, -1), audio.unsqueeze(1).expand(1, F, -1, -1), mask[None]
    if k == 3:  # dual CFG: branch 0 carries the unconditional (null-prompt) text encoding
        uncond = torch.randn(1, n_text, ctx_dim, generator=torch.Generator(device="cpu").manual_seed(seed + 1000))
        text_k = torch.cat([uncond, text, text]).unsqueeze(1).expand(k, F, -1, -1)
    else:
        text_k = text.expand(k, -1, -1).unsqueeze(1).expand(k, F, -1, -1)
    audio_k = torch.cat([null_audio] * (k - 1) + [audio]).unsqueeze(1).expand(k, F, -1, -1)
    mask_k = mask[None].expand(k, -1, -1).contiguous()
    return lat, text_k, audio_k, mask_k

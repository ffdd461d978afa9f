"""The drop-in claim, executed: AudioCondAnimationPipeline.__call__ and generate_videos (the calls
scripts/animation_demo.py:82-110 makes) run end to end on the B200 with stand-ins for the stock modules that cannot be
installed here (tests/stubs.py: tokenizer, text encoder, VAE, audio encoder, a diffusers-shaped PNDMScheduler, the
reference's I/O helpers) and the REAL hot path (CUDA UNet engine + fused CFG/sampler session).  The expected videos are
computed independently on the CPU: stub encoders -> reference-ordered CFG batches -> oracle UNet (oracle/unet_ref.py)
+ restated diffusers sampler (oracle/sampler_ref.py) -> stub VAE decode.
Reference: /root/reference/avgen/pipelines/pipeline_audio_cond_animation.py:264-375 (__call__), :379-468."""
import os

import numpy as np
import PIL.Image
import pytest
import torch

import stubs
from asva_b200 import schedulers, synth

pytestmark = pytest.mark.gpu
CHANS = (64, 128, 256, 256)
F, H, W = 4, 64, 64


def _unet():
    from avgen.models.unets import AudioUNet3DConditionModel
    m = AudioUNet3DConditionModel(sample_size=64, cross_attention_dim=768, attention_head_dim=8, block_out_channels=CHANS)
    sd = synth.synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], seed=0)
    m.load_state_dict(sd)
    return m, sd


def _pipeline(scheduler, unet):
    from avgen.pipelines.pipeline_audio_cond_animation import AudioCondAnimationPipeline
    pipe = AudioCondAnimationPipeline(text_encoder=stubs.StubTextEncoder(), tokenizer=stubs.StubTokenizer(), unet=unet,
                                      scheduler=scheduler, vae=stubs.StubVAE(), audio_encoder=stubs.StubAudioEncoder(F))
    pipe.to(torch_device=torch.device("cuda"), dtype=torch.float32)   # animation_demo.py:91
    pipe.set_progress_bar_config(disable=True)                        # :92
    pipe._audio_processor = stubs.StubMelExtractor()                  # the real one needs ImageBind's mel code
    return pipe


def _image(seed):
    g = np.random.default_rng(seed)
    return PIL.Image.fromarray(g.integers(0, 256, size=(H, W, 3), dtype=np.uint8))


def _expected_videos(sd, image, audio, text, steps, s_a, s_t, seed, sampler):
    """The reference's __call__ restated on the CPU with the stubs + oracle (one clip)."""
    from oracle import sampler_ref, unet_ref
    do_t, do_a = s_t > 1.0, s_a > 1.0
    tok, te, vae, ae = stubs.StubTokenizer(), stubs.StubTextEncoder(), stubs.StubVAE(), stubs.StubAudioEncoder(F)
    txt = te(tok([text]).input_ids)[0]
    unc = te(tok("").input_ids)[0]
    texts = ([unc, txt, txt] if do_t and do_a else [unc, txt] if do_t else [txt, txt] if do_a else [txt])
    mel = stubs.StubMelExtractor()([audio])
    _, a_enc, a_mask = ae(mel)
    _, n_enc, n_mask = ae(torch.zeros(1, 1, 128, 204))
    if do_t and do_a:
        auds, masks = [n_enc, n_enc, a_enc], [n_mask, n_mask, a_mask]
    elif do_t:
        auds, masks = [a_enc, a_enc], [a_mask, a_mask]
    elif do_a:
        auds, masks = [n_enc, a_enc], [n_mask, a_mask]
    else:
        auds, masks = [a_enc], [a_mask]
    text_k = torch.cat(texts).unsqueeze(1).expand(-1, F, -1, -1)
    audio_k = torch.cat(auds).unsqueeze(1).expand(-1, F, -1, -1)
    mask_k = torch.cat(masks)
    img = torch.from_numpy(np.asarray(image, dtype=np.float32) / 255.0).permute(2, 0, 1)[None] * 2.0 - 1.0
    z0 = vae.encode(img).latent_dist.sample() * vae.config.scaling_factor
    gen = torch.Generator(device="cuda").manual_seed(seed)
    noise = torch.randn((1, 4, F - 1, H // 8, W // 8), generator=gen, device="cuda", dtype=torch.float32).cpu()
    lat = torch.cat([z0.unsqueeze(2), noise], dim=2)
    cls = sampler_ref.DDIMRef if sampler == "ddim" else sampler_ref.PNDMRef
    with torch.no_grad():
        sampler_ref.denoise_loop(lambda x, t, a, b, c: unet_ref.unet_forward(sd, dict(block_out_channels=CHANS), x, t, a, b, c),
                                 cls(steps), lat, text_k, audio_k, mask_k, audio_scale=s_a, text_scale=s_t)
    flat = lat.permute(0, 2, 1, 3, 4).reshape(F, 4, H // 8, W // 8) / vae.config.scaling_factor
    vid = (vae.decode(flat).sample / 2 + 0.5).clamp(0, 1)
    return vid.view(1, F, 3, H, W), z0


@pytest.mark.parametrize("sampler,s_a,s_t", [("ddim", 4.0, 1.0), ("pndm", 4.0, 2.5), ("pndm", 1.0, 1.0)])
def test_pipeline_call_matches_cpu_restatement(cuda_backend, sampler, s_a, s_t):
    unet, sd = _unet()
    sched = schedulers.DDIMScheduler() if sampler == "ddim" else stubs.PNDMScheduler()
    pipe = _pipeline(sched, unet)
    image, audio, steps, seed = _image(1), torch.randn(1, 20000, generator=torch.Generator().manual_seed(2)) * 0.1, 4, 77
    gen = torch.Generator(device="cuda").manual_seed(seed)
    out = pipe(images=[image], audios=[audio], texts=["a dog barking"], video_length=F, height=H, width=W,
               num_inference_steps=steps, audio_guidance_scale=s_a, text_guidance_scale=s_t, generator=gen)
    vid = out["videos"]
    assert tuple(vid.shape) == (1, F, 3, H, W) and vid.device.type == "cpu" and vid.dtype == torch.float32
    assert float(vid.min()) >= 0.0 and float(vid.max()) <= 1.0
    exp, z0 = _expected_videos(sd, image, audio, "a dog barking", steps, s_a, s_t, seed, sampler)
    # frame 0 is the conditioning image's latent, never touched by the loop (:363-364): decode(encode(image)) exactly
    assert torch.allclose(vid[:, 0], exp[:, 0], atol=1e-6)
    d = (vid[:, 1:] - 0.5) - (exp[:, 1:] - 0.5)
    rel = float(d.norm() / (exp[:, 1:] - 0.5).norm())
    print(f"[pipeline] {sampler} s_a {s_a} s_t {s_t}: generated frames rel-L2 {rel:.3e}, "
          f"saturated {float(((exp == 0) | (exp == 1)).float().mean()):.3f}")
    assert rel <= 6e-2, rel   # toy-net bf16 error after `steps` sampler steps (anchor: tests/test_unet_gpu.py)
    if isinstance(sched, stubs.PNDMScheduler):
        assert sched.step_calls == 0, "a diffusers-shaped PNDMScheduler must take the fused CFG + sampler path"
        assert pipe.last_launches > 0
    # return_dict=False returns the bare tensor (generate_videos indexes [0] of it, :447); same seed -> same video
    gen.manual_seed(seed)
    again = pipe(images=[image], audios=[audio], texts=["a dog barking"], video_length=F, height=H, width=W,
                 num_inference_steps=steps, audio_guidance_scale=s_a, text_guidance_scale=s_t, generator=gen,
                 return_dict=False)
    assert torch.is_tensor(again) and torch.equal(again, vid)


def test_pipeline_call_precomputed_text_encodings_and_inference_mode(cuda_backend):
    """text_encodings= bypasses the tokenizer / text encoder (generate_videos passes category encodings, :436-441);
    the whole call also works under torch.inference_mode() (inference tensors have no version counter)."""
    unet, sd = _unet()
    pipe = _pipeline(schedulers.PNDMScheduler(), unet)
    pipe.tokenizer = pipe.text_encoder = None
    enc = stubs.StubTextEncoder()(stubs.StubTokenizer()(["x"]).input_ids)[0]
    image, audio = _image(3), torch.zeros(1, 100)
    with torch.inference_mode():
        v = pipe(images=[image], audios=[audio], texts=[""], text_encodings=[enc], video_length=F, height=H, width=W,
                 num_inference_steps=3, generator=torch.Generator(device="cuda").manual_seed(0), return_dict=False)
    assert tuple(v.shape) == (1, F, 3, H, W) and torch.isfinite(v).all()


def test_generate_videos_contract(cuda_backend, monkeypatch):
    """generate_videos (:379-468) with the reference's I/O helpers stubbed: one pipeline call per clip, generator
    re-seeded per clip, uint8 (F,H,W,3) videos + the audio clips returned when no save_template is given."""
    from avgen.pipelines import pipeline_audio_cond_animation as P
    unet, _ = _unet()
    pipe = _pipeline(stubs.PNDMScheduler(), unet)
    monkeypatch.setattr(P, "_data_utils", lambda: stubs.data_utils_stub())
    vids, auds = P.generate_videos(pipe, image_path="img.png", audio_path="clip.wav", category="dog",
                                   image_size=(H, W), video_fps=6, video_num_frame=F, num_clips_per_video=2,
                                   audio_guidance_scale=4.0, text_guidance_scale=1.0, seed=3, save_template="",
                                   device=torch.device("cuda"))
    assert len(vids) == 2 and len(auds) == 2
    for v in vids:
        assert v.dtype == torch.uint8 and tuple(v.shape) == (F, H, W, 3)
    assert torch.equal(vids[0][0], vids[1][0]), "same conditioning image -> same first frame"
    assert not torch.equal(vids[0][1:], vids[1][1:]), "different audio clips must give different videos"
    assert pipe.scheduler.step_calls == 0


def test_generate_videos_for_dataset_sharded(cuda_backend, monkeypatch, tmp_path):
    """generate_videos_for_dataset (:472-551, scripts/animation_gen.py:32-45) end to end with every external piece
    stubbed at its import site (diffusers AutoencoderKL / PNDMScheduler, transformers CLIP, ImageBind encoder, the
    reference's data helpers, torchvision.io.write_video) and the REAL UNet checkpoint path
    (AudioUNet3DConditionModel.from_pretrained on a directory written by save_pretrained).  Two ranks of a
    world_size-2 job must together write every <file>_clip-XX.mp4 exactly once (clip-level sharding)."""
    import json
    import sys
    import types

    import torchvision
    import transformers

    from avgen.models.unets import AudioUNet3DConditionModel
    from avgen.pipelines import pipeline_audio_cond_animation as P
    unet, _ = _unet()
    exp = tmp_path / "exp"
    unet.save_pretrained(str(exp / "ckpts" / "checkpoint-7" / "modules" / "unet"))
    files = [f"vid{i}.mp4" for i in range(5)]
    cats = ["dog", "rain", "dog", "drum", "rain"]
    monkeypatch.chdir(tmp_path)
    (tmp_path / "datasets" / "AVSync15").mkdir(parents=True)
    (tmp_path / "pretrained").mkdir()
    json.dump({c: c for c in set(cats)}, open(tmp_path / "datasets" / "AVSync15" / "class_mapping.json", "w"))
    enc = stubs.StubTextEncoder()
    tok = stubs.StubTokenizer()
    torch.save({c: enc(tok([c]).input_ids)[0] for c in set(cats)},
               tmp_path / "datasets" / "AVSync15" / "class_clip_text_encodings_stable-diffusion-v1-5.pt")
    torch.save(enc(tok("").input_ids)[0], tmp_path / "pretrained" / "openai-clip-l_null_text_encoding.pt")

    def from_pretrained_of(factory):
        return types.SimpleNamespace(from_pretrained=staticmethod(lambda *a, **k: factory()))

    fake_diffusers = types.ModuleType("diffusers")
    fake_models = types.ModuleType("diffusers.models")
    fake_models.AutoencoderKL = from_pretrained_of(stubs.StubVAE)
    fake_sched = types.ModuleType("diffusers.schedulers")
    fake_sched.PNDMScheduler = from_pretrained_of(stubs.PNDMScheduler)
    fake_audio = types.ModuleType("avgen.models.audio_encoders")
    fake_audio.ImageBindSegmaskAudioEncoder = lambda n_segment=12: stubs.StubAudioEncoder(n_segment)
    for name, mod in (("diffusers", fake_diffusers), ("diffusers.models", fake_models),
                      ("diffusers.schedulers", fake_sched), ("avgen.models.audio_encoders", fake_audio)):
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.setattr(transformers.CLIPTextModel, "from_pretrained", staticmethod(lambda *a, **k: stubs.StubTextEncoder()))
    monkeypatch.setattr(transformers.CLIPTokenizer, "from_pretrained", staticmethod(lambda *a, **k: stubs.StubTokenizer()))
    monkeypatch.setattr(P, "_data_utils", lambda: stubs.dataset_data_utils_stub(files, cats))
    written = []
    # (torchvision >= 0.24 no longer ships write_video; the reference pins an older one - the attribute is injected)
    monkeypatch.setattr(torchvision.io, "write_video",
                        lambda filename, video_array, fps, **kw: written.append((filename, tuple(video_array.shape))),
                        raising=False)
    for rank in (0, 1):
        P.generate_videos_for_dataset(str(exp), 7, dataset="AVSync15", image_size=(H, W), video_fps=6,
                                      video_num_frame=F, num_clips_per_video=2, audio_guidance_scale=4.0,
                                      text_guidance_scale=1.0, random_seed=0, device=torch.device("cuda", 0),
                                      dtype=torch.float32, rank=rank, world_size=2)
    names = sorted(os.path.basename(f) for f, _ in written)
    assert names == sorted(f"vid{i}_clip-{k:02d}.mp4" for i in range(5) for k in range(2)), names
    assert all(shape == (F, H, W, 3) for _, shape in written)
    root = os.path.dirname(written[0][0])
    assert root.endswith(os.path.join("evaluations", "checkpoint-7", "AG-4.0_TG-1.0", "seed-0", "videos")), root

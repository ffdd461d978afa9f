"""Whole-path parity on the B200 through the PUBLIC API (avgen.models.unets / avgen.pipelines): the CUDA engine vs
(1) tests/golden/ fixtures produced by executing the reference's own UNet files on CPU in fp32
(oracle/make_goldens.py) and (2) the clean-room CPU oracle on fresh seeds.
Stated tolerance (reference fp32 vs bf16-storage / fp32-accumulate kernels; SURVEY.md section 8(c) calibrates torch's
own bf16 eager run of the reference at rel-L2 1.45e-2): one UNet forward at the SD-1.5 geometry rel-L2 <= 2e-2 and
cosine >= 0.9995 (measured 1.3e-2 .. 1.4e-2); the 64..256-channel toy geometries (fewer channels to average the
rounding over) rel-L2 <= 4e-2 and N-step sampler latents on the toy geometry rel-L2 <= 4e-2.  The toy numbers are
noisy by construction: which bf16 roundings flip depends on the fp32 summation order inside a GEMM, and that order
changes with the tile plan the tuner picks (split-K or not) - measured 2.3e-2 .. 3.0e-2 across boxes for the same
seed, hence the margin.  The cosine bound (>= 0.9995) is the same everywhere."""
import glob
import os

import pytest
import torch

from asva_b200 import schedulers, synth

pytestmark = pytest.mark.gpu
TOL_TOY = 4e-2
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_MODELS = {}


def _model(chans):
    from avgen.models.unets import AudioUNet3DConditionModel
    chans = tuple(chans)
    if chans not in _MODELS:
        m = AudioUNet3DConditionModel(sample_size=64, cross_attention_dim=768, attention_head_dim=8,
                                      block_out_channels=chans)
        sd = synth.synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], seed=0)
        m.load_state_dict(sd)
        _MODELS.clear()  # one model resident at a time
        _MODELS[chans] = (m.to("cuda"), sd)
    return _MODELS[chans]


def _check(name, got, ref, tol):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert torch.isfinite(got).all(), f"{name}: non-finite"
    rel = float((got - ref).norm() / ref.norm())
    cos = float(torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0))
    print(f"[parity] {name}: rel-L2 {rel:.3e} cos {cos:.6f} max|d| {float((got - ref).abs().max()):.3e}")
    assert rel <= tol and cos >= 0.9995, (name, rel, cos)
    return rel


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "unet_tiny_*.pt"))), ids=os.path.basename)
def test_unet_vs_reference_golden_tiny(cuda_backend, path):
    g = torch.load(path)
    m, _ = _model(g["chans"])
    lat, text, audio, mask = synth.synth_inputs(F=g["F"], h=g["h"], w=g["w"], k=g["k"], seed=g["input_seed"])
    x = lat.expand(g["k"], -1, -1, -1, -1).contiguous().cuda()
    for rep in range(3):  # eager, graph capture, graph replay must all agree
        y = m(x, g["t"], encoder_hidden_states=text.cuda(), audio_encoder_hidden_states=audio.cuda(),
              audio_attention_mask=mask.cuda()).sample
        _check(f"{os.path.basename(path)} call {rep}", y, g["out"], TOL_TOY if "tiny" in path else 2e-2)


def test_unet_vs_reference_golden_sd15(cuda_backend):
    g = torch.load(os.path.join(GOLD, "unet_sd15_cfg2.pt"))
    m, _ = _model(g["chans"])
    lat, text, audio, mask = synth.synth_inputs(F=g["F"], h=g["h"], w=g["w"], k=g["k"], seed=g["input_seed"])
    x = lat.expand(g["k"], -1, -1, -1, -1).contiguous().cuda()
    for rep in range(3):
        y = m(x, torch.tensor(g["t"]), encoder_hidden_states=text.cuda(), audio_encoder_hidden_states=audio.cuda(),
              audio_attention_mask=mask.cuda()).sample
        _check(f"sd15 12x32x32 call {rep}", y, g["out"], 2e-2)


def test_unet_vs_oracle_fresh_seed_and_frame_varying_context(cuda_backend):
    """Seeds and shapes the goldens do not hold; contexts that differ per frame (the API allows (B,F,n,768))."""
    from oracle import unet_ref
    chans = (64, 128, 256, 256)
    m, sd = _model(chans)
    g = torch.Generator().manual_seed(7)
    B, F, h, w = 2, 6, 16, 8
    x = torch.randn(B, 4, F, h, w, generator=g)
    text = torch.randn(B, F, 77, 768, generator=g)
    audio = torch.randn(B, F, 229, 768, generator=g)
    mask = synth.audio_segment_mask(F)[None].expand(B, -1, -1).contiguous()
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, dict(block_out_channels=chans), x, 37, text, audio, mask)
    y = m(x.cuda(), 37, encoder_hidden_states=text.cuda(), audio_encoder_hidden_states=audio.cuda(),
          audio_attention_mask=mask.cuda(), return_dict=False)[0]
    _check("fresh seed, per-frame contexts", y, ref, TOL_TOY)
    y2 = m(x.cuda(), 37, encoder_hidden_states=text.cuda(), audio_encoder_hidden_states=audio.cuda(),
           audio_attention_mask=None).sample
    with torch.no_grad():
        ref2 = unet_ref.unet_forward(sd, dict(block_out_channels=chans), x, 37, text, audio, None)
    _check("no audio mask", y2, ref2, TOL_TOY)


@pytest.mark.parametrize("B,F,h,w,n_text,masked", [(1, 8, 24, 40, 77, True), (3, 4, 8, 16, 77, True),
                                                     (2, 16, 8, 8, 5, True), (2, 24, 16, 16, 77, False),
                                                     (2, 1, 8, 8, 77, True), (2, 12, 8, 24, 1, True)])
def test_unet_geometry_sweep(cuda_backend, B, F, h, w, n_text, masked):
    """Shapes off the headline path: no CFG (B = 1) and 3-way CFG, non-square latents whose rows are not a multiple of
    the 128-row tile, 1 / 16 / 24 frames (config 4's frame count), 1 .. 77 text keys, with and without the audio
    segment mask - each against the CPU oracle on a fresh seed."""
    from oracle import unet_ref
    chans = (64, 128, 256, 256)
    m, sd = _model(chans)
    g = torch.Generator().manual_seed(1000 + 31 * B + 7 * F + h + w + n_text)
    x = torch.randn(B, 4, F, h, w, generator=g)
    text = torch.randn(B, 1, n_text, 768, generator=g).expand(B, F, n_text, 768).contiguous()
    audio = torch.randn(B, 1, 229, 768, generator=g).expand(B, F, 229, 768).contiguous()
    mask = synth.audio_segment_mask(F)[None].expand(B, -1, -1).contiguous() if masked else None
    t = 11 + 40 * F
    with torch.no_grad():
        ref = unet_ref.unet_forward(sd, dict(block_out_channels=chans), x, t, text, audio, mask)
    for rep in range(2):  # first call (tuning pass + capture) and the replay
        y = m(x.cuda(), t, encoder_hidden_states=text.cuda(), audio_encoder_hidden_states=audio.cuda(),
              audio_attention_mask=mask.cuda() if masked else None).sample
        _check(f"B{B} F{F} {h}x{w} text{n_text} mask{int(masked)} call {rep}", y, ref, TOL_TOY)


@pytest.mark.parametrize("name", ["ddim", "pndm"])
def test_sampler_trace_vs_golden(cuda_backend, name):
    from avgen.pipelines.pipeline_audio_cond_animation import AudioCondAnimationPipeline
    g = torch.load(os.path.join(GOLD, f"sampler_{name}.pt"))
    m, _ = _model(g["chans"])
    sched = schedulers.DDIMScheduler() if name == "ddim" else schedulers.PNDMScheduler()
    pipe = AudioCondAnimationPipeline(None, None, m, sched, None, None)
    pipe.set_progress_bar_config(disable=True)
    lat, text, audio, mask = synth.synth_inputs(F=g["F"], h=g["h"], w=g["w"], k=2)
    trace = []
    for run in range(2):  # second run replays the captured graph and must reproduce the first
        trace.clear()
        out = pipe.denoise(lat.cuda(), text.cuda(), audio.cuda(), mask.cuda(), g["steps"],
                           audio_guidance_scale=g["audio_scale"],
                           callback=lambda i, t, l: trace.append(l.clone().cpu()))
        assert len(trace) == g["trace"].shape[0]
        assert torch.equal(out.cpu()[:, :, 0], lat[:, :, 0]), "conditioning frame must never change"
        for i in (0, 1, 2, len(trace) - 1):
            _check(f"{name} run {run} after step {i + 1}", trace[i], g["trace"][i], TOL_TOY)
    assert pipe.last_launches > 0


def test_generic_scheduler_path_matches_fused(cuda_backend):
    """A scheduler the fused kernel does not recognise goes through scheduler.step(); both routes must agree."""
    from avgen.pipelines.pipeline_audio_cond_animation import AudioCondAnimationPipeline
    m, _ = _model((64, 128, 256, 256))

    class Foreign(schedulers.DDIMScheduler):  # not isinstance-matched by plan_for via a wrapper
        pass

    lat, text, audio, mask = synth.synth_inputs(F=4, h=8, w=8, k=2)
    args = (lat.cuda(), text.cuda(), audio.cuda(), mask.cuda(), 4)
    fused = AudioCondAnimationPipeline(None, None, m, schedulers.DDIMScheduler(), None, None)
    fused.set_progress_bar_config(disable=True)
    a = fused.denoise(*args)
    gen = AudioCondAnimationPipeline(None, None, m, schedulers.DDIMScheduler(), None, None)
    gen.set_progress_bar_config(disable=True)
    gen.scheduler.set_timesteps(4)
    b = gen._denoise_generic(args[0], args[1], args[2], args[3], 2, False, True, 4.0, 1.0, None)
    _check("generic vs fused loop", b, a, 2e-2)


def test_context_swap_after_graph_capture(cuda_backend):
    """A captured step graph must see the NEXT clip's conditioning (persistent context buffers), including a
    change of context geometry (ragged mask -> rule mask) that forces a re-capture."""
    from oracle import unet_ref
    chans = (64, 128, 256, 256)
    m, sd = _model(chans)
    B, F, h, w = 2, 4, 8, 8
    g = torch.Generator().manual_seed(99)
    x = torch.randn(B, 4, F, h, w, generator=g)

    def run(text, audio, mask, name, reps):
        with torch.no_grad():
            ref = unet_ref.unet_forward(sd, dict(block_out_channels=chans), x, 500, text, audio, mask)
        for r in range(reps):
            y = m(x.cuda(), 500, encoder_hidden_states=text.cuda(), audio_encoder_hidden_states=audio.cuda(),
                  audio_attention_mask=mask.cuda()).sample
            _check(f"{name} rep {r}", y, ref, TOL_TOY)

    rule = synth.audio_segment_mask(F)[None].expand(B, -1, -1).contiguous()
    t1 = torch.randn(B, 1, 77, 768, generator=g).expand(B, F, 77, 768)
    a1 = torch.randn(B, 1, 229, 768, generator=g).expand(B, F, 229, 768)
    run(t1, a1, rule, "clip 1", 3)                       # eager, capture, replay
    t2 = torch.randn(B, 1, 77, 768, generator=g).expand(B, F, 77, 768)
    a2 = torch.randn(B, 1, 229, 768, generator=g).expand(B, F, 229, 768)
    run(t2, a2, rule, "clip 2 (same geometry, replayed graph)", 2)
    ragged = torch.rand(B, F, 229, generator=g) < 0.2
    ragged[:, :, 0] = True
    run(t2, a2, ragged, "clip 3 (ragged mask)", 3)

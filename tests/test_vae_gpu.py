"""SD-1.5 VAE decoder on the B200 (asva_b200/vae.py) against the restated AutoencoderKL
decoder (oracle/vae_ref.py, fp32 CPU).  Tolerance: rel-L2 <= 1.5 x the error of the same restatement run in bfloat16
(the anchor of tests/test_unet_gpu.py), cosine >= 0.9995."""
import math
import os

import pytest
import torch

import stubs
from asva_b200 import synth, vae
from sim_backend import SimBackend

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_softmax_rows(cuda_backend):
    g = torch.Generator().manual_seed(5)
    for rows, cols, ld in ((64, 64, 64), (300, 1024, 1024), (17, 4096, 4100)):
        s = (torch.randn(rows, ld, generator=g) * 4).to(DEV)
        p_cu = torch.zeros(rows, ld, dtype=torch.bfloat16, device=DEV)
        p_ref = torch.zeros_like(p_cu)
        cuda_backend.softmax_rows(s, p_cu, rows, cols, 0.37)
        SimBackend().softmax_rows(s, p_ref, rows, cols, 0.37)
        torch.cuda.synchronize()
        assert (p_cu[:, cols:] == 0).all()
        err = float((p_cu.float() - p_ref.float()).norm() / p_ref.float().norm())
        assert err < 4e-3, (rows, cols, err)
        assert float((p_cu[:, :cols].float().sum(-1) - 1).abs().max()) < 2e-2


@pytest.mark.parametrize("n,h,w", [(2, 16, 16), (3, 8, 24), (1, 32, 32)])
def test_vae_decode_vs_oracle(cuda_backend, n, h, w):
    from oracle import vae_ref
    sd = synth.synth_state_dict(vae_ref.state_dict_shapes(), seed=3)
    z = torch.randn(n, 4, h, w, generator=torch.Generator().manual_seed(100 + n + h))
    with torch.no_grad():
        want = vae_ref.decode(sd, z)
        low = vae_ref.decode(sd, z, dtype=torch.bfloat16).float()
    anchor = float((low - want).norm() / want.norm())
    eng = vae.VAEDecoderEngine(sd, device=DEV)
    for rep in range(2):  # the first call measures tile plans; the second reuses them
        got = eng.decode(z.to(DEV)).cpu()
        assert got.shape == (n, 3, 8 * h, 8 * w) and torch.isfinite(got).all()
        rel = float((got - want).norm() / want.norm())
        cos = float(torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0))
        print(f"[parity] vae decode n{n} {h}x{w} call {rep}: rel-L2 {rel:.3e} (bf16-eager anchor {anchor:.3e}) cos {cos:.6f}")
        assert rel <= 1.5 * anchor and cos >= 0.9995, (rel, anchor, cos)


@pytest.mark.parametrize("n,H,W", [(1, 256, 256), (2, 128, 192)])
def test_vae_encode_vs_oracle(cuda_backend, n, H, W):
    from oracle import vae_ref
    sd = synth.synth_state_dict(vae_ref.encoder_state_dict_shapes(), seed=4)
    x = torch.rand(n, 3, H, W, generator=torch.Generator().manual_seed(200 + n)) * 2 - 1
    with torch.no_grad():
        want = vae_ref.encode_moments(sd, x)
        low = vae_ref.encode_moments(sd, x, dtype=torch.bfloat16).float()
    anchor = float((low - want).norm() / want.norm())
    eng = vae.VAEEncoderEngine(sd, device=DEV)
    for rep in range(2):
        got = eng.encode_moments(x.to(DEV)).cpu()
        assert got.shape == (n, 8, H // 8, W // 8) and torch.isfinite(got).all()
        rel = float((got - want).norm() / want.norm())
        cos = float(torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0))
        print(f"[parity] vae encode n{n} {H}x{W} call {rep}: rel-L2 {rel:.3e} (bf16-eager anchor {anchor:.3e}) cos {cos:.6f}")
        assert rel <= 1.5 * anchor and cos >= 0.9995, (rel, anchor, cos)


def test_pipeline_decodes_with_the_engine(cuda_backend, monkeypatch):
    """decode_latents wraps an AutoencoderKL-shaped module (diffusers state-dict keys) in FastDecodeVAE; the videos
    match the module's own decode, and ASVA_STOCK_VAE=1 leaves the module alone."""
    from avgen.pipelines.pipeline_audio_cond_animation import AudioCondAnimationPipeline
    v = stubs.StubAutoencoderKL().to(DEV)
    pipe = AudioCondAnimationPipeline(None, None, None, None, v, None)
    lat = torch.randn(4, 4, 16, 16, generator=torch.Generator().manual_seed(9)).to(DEV) * 0.18215
    monkeypatch.setenv("ASVA_STOCK_VAE", "1")
    stock = pipe.decode_latents(lat)
    assert pipe.vae is v
    monkeypatch.setenv("ASVA_STOCK_VAE", "0")
    fast = pipe.decode_latents(lat)
    assert isinstance(pipe.vae, vae.FastDecodeVAE) and pipe.vae.inner is v
    assert fast.shape == stock.shape == (4, 3, 128, 128) and fast.device.type == "cpu"
    rel = float((fast - stock).norm() / stock.norm())
    print(f"[parity] pipeline.decode_latents engine vs stock module: rel-L2 {rel:.3e}")
    assert rel < 2e-2
    # config / dtype still come from the wrapped module; encode runs on the engine too (diffusers encoder keys present)
    assert pipe.vae.config.scaling_factor == 0.18215 and pipe.vae.dtype == torch.float32
    img = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(3)).to(DEV) * 2 - 1
    dist = pipe.vae.encode(img).latent_dist
    ref = v.encode(img).latent_dist
    assert dist.sample().shape == (1, 4, 8, 8)
    rel_e = float((dist.mean - ref.mean).norm() / ref.mean.norm())
    print(f"[parity] FastDecodeVAE.encode mean vs stock module: rel-L2 {rel_e:.3e}")
    assert rel_e < 3e-2 and isinstance(pipe.vae._enc, vae.VAEEncoderEngine)

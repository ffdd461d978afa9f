"""Per-block parity at SD-1.5 channel widths (C1 / R1 / T1 / A1-A4 / FF1 at every level).

The whole-UNet goldens say THAT the network is right; these say WHERE it is not when one of them turns red: every kind
of block of the engine (ResBlock with and without shortcut / skip concat, the transformer with its five sub-blocks,
the down- and upsampler convs) is run on its own, at the real width and head dimension of each level (320 / 640 / 1280
channels, d = 40 / 80 / 160) on a small 4-frame 8x8 geometry, against the oracle's restatement of the same reference
block (oracle/unet_ref.py: resblock -> resnets/ff_spatio_temp_resnet_3d.py:161-191, transformer ->
transformers/ff_spatio_audio_temp_transformer_3d.py:94-158,278-373, ff_conv -> utils.py:22-57) with the same weights.
Tolerance: rel-L2 <= 1.2e-2 per block (bf16 storage, fp32 accumulation; measured: ResBlocks 3.5-4.0e-3, transformers
5.8-5.9e-3, down- and upsamplers 3.1e-3)."""
import pytest
import torch

from asva_b200 import synth
from test_unet_gpu import _check, _model

pytestmark = pytest.mark.gpu
TOL = 1.2e-2
CHANS = (320, 640, 1280, 1280)
B, F, H, W = 2, 4, 8, 8
T = 400.0


def _cl(x):  # (B,C,F,h,w) fp32 -> channels-last bf16 rows on the device, and the same values back as fp32 NCFHW
    xb = x.to(torch.bfloat16)
    rows = xb.permute(0, 2, 3, 4, 1).reshape(-1, x.shape[1]).contiguous().cuda()
    return rows, xb.float()


def _ncfhw(rows, C, h, w):
    return rows.float().cpu().view(B, F, h, w, C).permute(0, 4, 1, 2, 3)


@pytest.fixture(scope="module")
def rig(cuda_backend):
    from oracle import unet_ref
    m, sd = _model(CHANS)
    eng = m.engine()
    lat, text, audio, mask = synth.synth_inputs(F=F, h=H, w=W, k=B, seed=77)
    eng.prepare(B, F, H, W)
    eng.set_context(text.cuda(), audio.cuda(), mask.cuda())
    out = torch.empty(B, 4, F, H, W, device="cuda")
    eng.forward(lat.expand(B, -1, -1, -1, -1).contiguous().cuda(), torch.full((B,), T, device="cuda"), out)  # fills tproj / pos
    torch.cuda.synchronize()
    eng._frozen = False  # the block calls below use their own buffer shapes
    with torch.no_grad():
        temb = unet_ref.time_mlp(sd, "time_embedding", unet_ref.sinusoid(torch.full((B,), T), CHANS[0]))
    yield dict(eng=eng, sd=sd, temb=temb, text=text, audio=audio, mask=mask, ref=unet_ref)
    m._eng = None  # the next test that uses the model prepares its own geometry
    m._runner = None
    import gc
    gc.collect()  # free the 2.3 GB engine now, not at some later collection
    torch.cuda.empty_cache()


def _x(C, seed, h=H, w=W):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, C, F, h, w, generator=g)


RES = [  # (engine path, state-dict prefix, C of x0, C of the skip x1)
    ("down.0.res.0", "down_blocks.0.resnets.0", 320, 0), ("down.1.res.0", "down_blocks.1.resnets.0", 320, 0),
    ("down.2.res.1", "down_blocks.2.resnets.1", 1280, 0), ("down.3.res.0", "down_blocks.3.resnets.0", 1280, 0),
    ("mid.res.0", "mid_block.resnets.0", 1280, 0), ("up.0.res.0", "up_blocks.0.resnets.0", 1280, 1280),
    ("up.1.res.2", "up_blocks.1.resnets.2", 1280, 640), ("up.2.res.2", "up_blocks.2.resnets.2", 640, 320),
    ("up.3.res.1", "up_blocks.3.resnets.1", 320, 320),
]


def _get(eng, path):
    parts = path.split(".")
    o = getattr(eng, parts[0])
    for p in parts[1:]:
        o = o[int(p)] if p.isdigit() else o[p]
    return o


@pytest.mark.parametrize("path,prefix,c0,c1", RES, ids=[r[1] for r in RES])
def test_resblock_vs_oracle(rig, path, prefix, c0, c1):
    eng, sd, ref = rig["eng"], rig["sd"], rig["ref"]
    r = _get(eng, path)
    x0r, x0 = _cl(_x(c0, 500 + c0 + c1))
    x1r, x1 = _cl(_x(c1, 600 + c0 + c1)) if c1 else (None, None)
    y = eng._resblock(r, x0r, x1r, B, F, H, W, "blk_test")
    torch.cuda.synchronize()
    with torch.no_grad():
        want = ref.resblock(sd, prefix, torch.cat([x0, x1], 1) if c1 else x0, rig["temb"], 32, 1e-5)
    _check(f"resblock {prefix}", _ncfhw(y, want.shape[1], H, W), want, TOL)


ATT = [("down.0.attn.0", "down_blocks.0.attentions.0", 320), ("down.1.attn.1", "down_blocks.1.attentions.1", 640),
       ("down.2.attn.0", "down_blocks.2.attentions.0", 1280), ("mid.attn", "mid_block.attentions.0", 1280),
       ("up.1.attn.2", "up_blocks.1.attentions.2", 1280), ("up.2.attn.0", "up_blocks.2.attentions.0", 640),
       ("up.3.attn.2", "up_blocks.3.attentions.2", 320)]


@pytest.mark.parametrize("path,prefix,C", ATT, ids=[a[1] for a in ATT])
def test_transformer_vs_oracle(rig, path, prefix, C):
    """GroupNorm -> proj_in -> first-frame / audio / text / temporal attention -> GEGLU FF -> proj_out + residual."""
    eng, sd, ref = rig["eng"], rig["sd"], rig["ref"]
    a = _get(eng, path)
    idx = [i for i, t in enumerate(eng.attns) if t is a][0]
    xr, x = _cl(_x(C, 700 + idx))
    y = eng._transformer(a, xr, B, F, H, W, idx, "blk_test")
    torch.cuda.synchronize()
    with torch.no_grad():
        want = ref.transformer(sd, prefix, x, rig["text"], rig["audio"], rig["mask"], 8, 32)
    _check(f"transformer {prefix} (block {idx})", _ncfhw(y, C, H, W), want, TOL)


@pytest.mark.parametrize("level", [0, 1, 2])
def test_downsampler_vs_oracle(rig, level):
    eng, sd, ref = rig["eng"], rig["sd"], rig["ref"]
    cv = eng.down[level]["down"]
    xr, x = _cl(_x(cv.cin, 800 + level))
    y, ho, wo = eng._conv3(cv, xr, B, F, H, W, stride=2)
    out = eng.buf("blk_ds", (B * F * ho * wo, cv.cout))
    eng._ffconv_tail(cv, y, out, B, F, ho * wo)
    torch.cuda.synchronize()
    with torch.no_grad():
        want = ref.ff_conv(sd, f"down_blocks.{level}.downsamplers.0.conv", x, stride=2)
    _check(f"downsampler {level}", _ncfhw(out, cv.cout, ho, wo), want, TOL)


@pytest.mark.parametrize("level", [0, 1, 2])
def test_upsampler_vs_oracle(rig, level):
    """Nearest 2x (fused into the GroupNorm apply kernel's copy mode) + FFInflatedConv3d."""
    eng, sd, ref = rig["eng"], rig["sd"], rig["ref"]
    cv = eng.up[level]["up"]
    xr, x = _cl(_x(cv.cin, 900 + level, 4, 4))
    u = eng.buf("blk_up_in", (B * F * 64, cv.cin))
    eng.be.groupnorm_apply(xr, cv.cin, None, 0, None, B, B * F, 4, 4, False, True, u)
    y, _, _ = eng._conv3(cv, u, B, F, 8, 8)
    out = eng.buf("blk_up", (B * F * 64, cv.cout))
    eng._ffconv_tail(cv, y, out, B, F, 64)
    torch.cuda.synchronize()
    with torch.no_grad():
        up = torch.nn.functional.interpolate(x.permute(0, 2, 1, 3, 4).reshape(B * F, cv.cin, 4, 4), scale_factor=2.0,
                                             mode="nearest").view(B, F, cv.cin, 8, 8).permute(0, 2, 1, 3, 4)
        want = ref.ff_conv(sd, f"up_blocks.{level}.upsamplers.0.conv", up)
    _check(f"upsampler {level}", _ncfhw(out, cv.cout, 8, 8), want, TOL)

"""Generates tests/golden/*.pt by EXECUTING THE REFERENCE's own UNet files (oracle/ref_loader.py) on CPU in fp32 with
the deterministic synthetic weights/inputs of asva_b200/synth.py.  Run here (where /root/reference exists):
    python -m oracle.make_goldens [--full] [--only tiny,samplers,sd15,cfg4,traces]
The fixtures pin both the clean-room oracle (CPU tests) and the CUDA engine (GPU tests) to the reference.

  tiny      four toy geometries, one forward each                                   (seconds)
  samplers  6-step DDIM / PNDM traces on a toy geometry, audio-only and dual CFG    (seconds)
  sd15      the full SD-1.5 geometry at the config-2 shape  (k=2, 12f x 32x32)      (~1 min, 1.17 B parameters)
  cfg4      the full SD-1.5 geometry at the config-4 shape  (k=2, 24f x 64x64)      (~5 min)
  traces    50-step DDIM and 51-call PNDM at the SD-1.5 config-2 geometry; latents after steps 1, 2, 3, N (~40 min)
`--full` = everything; no flag = tiny + samplers.

Every unet_*.pt also records `bf16_eager_rel`: the rel-L2 error of torch's own eager run of the SAME reference model
in pure bfloat16 (weights and activations) against its fp32 run - the anchor SURVEY.md section 8(c) asks for
(err(ours) <= 1.5 x err(torch-bf16 eager) on the same inputs), so the GPU tolerances are tied to what bf16 can do."""
import argparse
import copy
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from asva_b200 import synth  # noqa: E402
from oracle import ref_loader, sampler_ref  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SD15 = (320, 640, 1280, 1280)

TINY_CASES = [
    # name, block_out_channels, k, F, h, w, timestep
    ("tiny_a", (64, 128, 256, 256), 2, 4, 8, 8, 981),
    ("tiny_b", (64, 128, 256, 256), 1, 5, 8, 16, 501),   # non-square, odd F, single branch
    ("tiny_c", (128, 128, 192, 256), 3, 3, 16, 16, 1),   # dual CFG batch, head dims 16/16/24/32
    ("tiny_d", (64, 64, 128, 128), 2, 12, 8, 8, 261),    # F = 12 -> 25 valid audio keys per frame
]


def ref_config(chans):
    return dict(sample_size=64, cross_attention_dim=768, attention_head_dim=8, norm_eps=1e-5,
                block_out_channels=tuple(chans))


def build(chans, seed=0):
    m = ref_loader.build_reference_unet(ref_config(chans))
    sd = synth.synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], seed=seed)
    m.load_state_dict(sd)
    return m


def run_ref(m, lat, t, text, audio, mask):
    with torch.no_grad():
        return m(lat, t, encoder_hidden_states=text, audio_encoder_hidden_states=audio,
                 audio_attention_mask=mask).sample


def rel_l2(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def bf16_eager_rel(m, y32, lat, t, text, audio, mask, in_place=False):
    """torch's own bf16 eager run of the reference model vs its fp32 output (the SURVEY 8(c) anchor)."""
    mb = m if in_place else copy.deepcopy(m)
    mb = mb.to(torch.bfloat16)
    yb = run_ref(mb, lat.to(torch.bfloat16), t, text.to(torch.bfloat16), audio.to(torch.bfloat16), mask)
    return rel_l2(yb, y32)


def unet_fixture(name, chans, k, F, h, w, t, anchor=True):
    t0 = time.time()
    m = build(chans)
    lat, text, audio, mask = synth.synth_inputs(F=F, h=h, w=w, k=k)
    x = lat.expand(k, -1, -1, -1, -1).contiguous()
    y = run_ref(m, x, t, text, audio, mask)
    rec = dict(chans=tuple(chans), k=k, F=F, h=h, w=w, t=t, seed=0, input_seed=123, out=y.clone())
    if anchor:
        rec["bf16_eager_rel"] = bf16_eager_rel(m, y, x, t, text, audio, mask, in_place=True)
    torch.save(rec, os.path.join(GOLD, f"unet_{name}.pt"))
    print(name, tuple(y.shape), "std %.4f" % float(y.std()), "bf16-eager rel %.3e" % rec.get("bf16_eager_rel", -1),
          "%.1fs" % (time.time() - t0), flush=True)


def sampler_traces(tag, chans, F, h, w, n, keep, scales=((4.0, 1.0),), anchor=False):
    """Reference UNet + restated diffusers step + restated pipeline loop -> sampler_<name><tag>.pt
    anchor: also run the loop with the reference model in pure bf16 (latents / sampler arithmetic stay fp32, as in
    the product) and record its rel-L2 error per kept step as `bf16_eager_rel`."""
    m = build(chans)
    mb = copy.deepcopy(m).to(torch.bfloat16) if anchor else None
    for a_s, t_s in scales:
        k = 1 + int(a_s > 1.0) + int(t_s > 1.0)
        for sname, cls in (("ddim", sampler_ref.DDIMRef), ("pndm", sampler_ref.PNDMRef)):
            t0 = time.time()
            lat, text, audio, mask = synth.synth_inputs(F=F, h=h, w=w, k=k)
            trace = []
            sampler_ref.denoise_loop(lambda x, t, a, b, c: run_ref(m, x, t, a, b, c), cls(n), lat.clone(), text,
                                     audio, mask, audio_scale=a_s, text_scale=t_s, trace=trace)
            idx = list(range(len(trace))) if keep is None else [i for i in keep if i < len(trace) - 1] + [len(trace) - 1]
            suffix = tag + ("_dual" if k == 3 else "")
            rec = dict(chans=tuple(chans), F=F, h=h, w=w, steps=n, audio_scale=a_s, text_scale=t_s, k=k,
                       calls=len(trace), kept=idx, trace=torch.stack([trace[i] for i in idx]))
            if anchor:
                tb = []
                sampler_ref.denoise_loop(
                    lambda x, t, a, b, c: run_ref(mb, x.to(torch.bfloat16), t, a.to(torch.bfloat16),
                                                  b.to(torch.bfloat16), c).float(),
                    cls(n), lat.clone(), text, audio, mask, audio_scale=a_s, text_scale=t_s, trace=tb)
                rec["bf16_eager_rel"] = [rel_l2(tb[i], trace[i]) for i in idx]
                print("   bf16-eager rel per kept step:", ["%.2e" % v for v in rec["bf16_eager_rel"]], flush=True)
            torch.save(rec, os.path.join(GOLD, f"sampler_{sname}{suffix}.pt"))
            print(sname + suffix, len(trace), "calls, kept", idx, "final std %.4f" % float(trace[-1].std()),
                  "%.1fs" % (time.time() - t0), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    what = set(args.only.split(",")) if args.only else ({"tiny", "samplers", "sd15", "cfg4", "traces"} if args.full
                                                        else {"tiny", "samplers"})
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if "tiny" in what:
        for name, chans, k, F, h, w, t in TINY_CASES:
            unet_fixture(name, chans, k, F, h, w, t)
    if "samplers" in what:
        sampler_traces("", (64, 128, 256, 256), 4, 8, 8, 6, None, scales=((4.0, 1.0), (4.0, 2.5)), anchor=True)
    if "sd15" in what:
        unet_fixture("sd15_cfg2", SD15, 2, 12, 32, 32, 981)
    if "cfg4" in what:
        unet_fixture("sd15_cfg4", SD15, 2, 24, 64, 64, 481, anchor=False)
    if "traces" in what:
        sampler_traces("_sd15", SD15, 12, 32, 32, 50, (0, 1, 2))


if __name__ == "__main__":
    main()

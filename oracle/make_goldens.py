"""Generates tests/golden/*.pt by EXECUTING THE REFERENCE's own UNet files (oracle/ref_loader.py) on CPU in fp32 with
the deterministic synthetic weights/inputs of asva_b200/synth.py.  Run here (where /root/reference exists):
    python -m oracle.make_goldens [--full]
The fixtures pin both the clean-room oracle (CPU tests) and the CUDA engine (GPU tests) to the reference.
--full also writes the SD-1.5-geometry fixture (1.17 B parameters, ~1 min)."""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for name, chans, k, F, h, w, t in TINY_CASES:
        m = build(chans)
        lat, text, audio, mask = synth.synth_inputs(F=F, h=h, w=w, k=k)
        y = run_ref(m, lat.expand(k, -1, -1, -1, -1).contiguous(), t, text, audio, mask)
        torch.save(dict(chans=chans, k=k, F=F, h=h, w=w, t=t, seed=0, input_seed=123, out=y.clone()),
                   os.path.join(GOLD, f"unet_{name}.pt"))
        print(name, tuple(y.shape), float(y.std()))
    # sampler traces: reference UNet + restated diffusers step + restated pipeline loop
    chans = (64, 128, 256, 256)
    m = build(chans)
    for sname, cls, n in (("ddim", sampler_ref.DDIMRef, 6), ("pndm", sampler_ref.PNDMRef, 6)):
        lat, text, audio, mask = synth.synth_inputs(F=4, h=8, w=8, k=2)
        trace = []
        sampler_ref.denoise_loop(lambda x, t, a, b, c: run_ref(m, x, t, a, b, c), cls(n), lat.clone(), text, audio,
                                 mask, audio_scale=4.0, trace=trace)
        torch.save(dict(chans=chans, F=4, h=8, w=8, steps=n, audio_scale=4.0, trace=torch.stack(trace)),
                   os.path.join(GOLD, f"sampler_{sname}.pt"))
        print(sname, len(trace), float(trace[-1].std()))
    if args.full:
        t0 = time.time()
        m = build((320, 640, 1280, 1280))
        lat, text, audio, mask = synth.synth_inputs(F=12, h=32, w=32, k=2)
        y = run_ref(m, lat.expand(2, -1, -1, -1, -1).contiguous(), 981, text, audio, mask)
        torch.save(dict(chans=(320, 640, 1280, 1280), k=2, F=12, h=32, w=32, t=981, seed=0, input_seed=123,
                        out=y.clone()), os.path.join(GOLD, "unet_sd15_cfg2.pt"))
        print("sd15", tuple(y.shape), float(y.std()), "%.1fs" % (time.time() - t0))


if __name__ == "__main__":
    main()

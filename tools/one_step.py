"""Runs a few eager (no CUDA graph) denoising steps of the headline workload - the command ncu wraps.
    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv --log-file out.csv \
        python tools/one_step.py --steps 3
Prints the number of launches per step and the launch index where the last step starts."""
import argparse
import os
import sys

os.environ["ASVA_NO_GRAPH"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from asva_b200 import schedulers, synth  # noqa: E402
from avgen.models.unets import AudioUNet3DConditionModel  # noqa: E402
from avgen.pipelines.pipeline_audio_cond_animation import AudioCondAnimationPipeline  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
F, h, w, _ = bench.WORKLOADS[args.workload]
chans = bench.CHANS[args.workload]
sd = bench._build_weights(chans)
with torch.device("meta"):
    model = AudioUNet3DConditionModel(sample_size=64, cross_attention_dim=768, attention_head_dim=8,
                                      block_out_channels=chans)
model.load_state_dict(sd, assign=True)
model.to("cuda")
pipe = AudioCondAnimationPipeline(None, None, model, schedulers.DDIMScheduler(), None, None)
pipe.set_progress_bar_config(disable=True)
lat, text, audio, mask = synth.synth_inputs(F=F, h=h, w=w, k=2)
sess = pipe.open_session(text.cuda(), audio.cuda(), mask.cuda(), F, h, w, 50, audio_guidance_scale=4.0)
be = model.engine().be
sess.load_latents(lat.cuda())
marks = []
for i in range(args.steps):
    marks.append(be.launches)
    sess.step(i)
torch.cuda.synchronize()
print(f"launches before each step (C-ABI kernels only): {marks}; per step: {be.launches - marks[-1]}")

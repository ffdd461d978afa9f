#!/usr/bin/env python
"""bench.py - UNet denoise steps/s of ASVA's hot path on B200 (BASELINE.json metric).

One step = one CFG-batched UNet forward (k = 2: [text only, text+audio]) + CFG combine + DDIM update of frames 1..
for one 12-frame 256x256 clip (latents 1x4x12x32x32).  One clip per GPU, N independent clips at --gpus N (weak
scaling, no collective on the data path; a single NCCL all-gather of the final latents after the timed region).

  value        whole-job steps/s with the clip resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e          same steps through the pipeline's public session API with HOST latents: every step copies the latents
               host->device from pinned memory, runs the step and reads the updated latents back device->host
  roofline     dominant kernel family (tcgen05 GEMM / implicit-GEMM conv, asva_gemm): algorithmic GEMM FLOPs per step
               / the duration of that family's launches of one step (re-captured as one CUDA graph and timed with
               CUDA events) vs the measured sustained bf16 peak in MEASURED_PEAKS.json
  cpu_baseline the reference's own UNet code (oracle/_ref staged copy; else the oracle port) timed on this box's host
               cores on a bounded sample (rank 0, N = 1 only)
`--impl reference` times that CPU implementation as its own arm."""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (F, latent h, latent w, description)
    "cfg2": (12, 32, 32, "50-step-DDIM clip 12fx256x256, audio_guidance 4.0 (CFG batch 2), random-init SD-1.5-geometry "
                         "AVSyncD UNet, synthetic audio/text context"),
    "cfg4": (24, 64, 64, "24fx512x512 long clip, audio_guidance 4.0 (CFG batch 2)"),
    "tiny": (4, 8, 8, "plumbing config (64,128,256,256) channels"),
}
CHANS = {"cfg2": (320, 640, 1280, 1280), "cfg4": (320, 640, 1280, 1280), "tiny": (64, 128, 256, 256)}
METRIC = "UNet denoise steps/sec (12f x 256^2, CFG x2)"
UNIT = "steps/s"


def _config(args, desc):
    """Identical in both arms (the driver compares them key by key)."""
    return {"workload": f"{args.workload}: {desc}", "clips_per_gpu": args.clips_per_gpu, "sampler": "DDIM eta=0",
            "l2": "no flush: the step streams 2.34 GB of bf16 weights (4.5 GB fp32 on the CPU arm) + activations, "
                  "far more than the 126 MB L2"}


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:  # noqa: BLE001
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def _build_weights(chans):
    from asva_b200 import synth
    from avgen.models.unets import AudioUNet3DConditionModel
    with torch.device("meta"):
        skel = AudioUNet3DConditionModel(sample_size=64, cross_attention_dim=768, attention_head_dim=8,
                                         block_out_channels=chans)
    return synth.synth_state_dict([(k, tuple(v.shape)) for k, v in skel.state_dict().items()], seed=0)


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_steps(sd, chans, F, h, w, warmup, steps, budget_s):
    """Times the reference's own UNet code (fp32, all host threads) + restated CFG/DDIM step on full-size steps.
    Returns (seconds per step, steps timed, kind, cores)."""
    from asva_b200 import synth
    from oracle import ref_loader, sampler_ref, unet_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lat, text, audio, mask = synth.synth_inputs(F=F, h=h, w=w, k=2)
    if ref_loader.available():
        kind = "reference"
        m = ref_loader.build_reference_unet(dict(sample_size=64, cross_attention_dim=768, attention_head_dim=8,
                                                 norm_eps=1e-5, block_out_channels=chans))
        m.load_state_dict(sd)

        def unet(x, t, a, b, c):
            return m(x, t, encoder_hidden_states=a, audio_encoder_hidden_states=b, audio_attention_mask=c).sample
    else:
        kind = "port"

        def unet(x, t, a, b, c):
            return unet_ref.unet_forward(sd, dict(block_out_channels=chans), x, t, a, b, c)
    sched = sampler_ref.DDIMRef(50)
    ts = sched.timesteps.tolist()
    lat = lat.clone()

    def one(i):
        t = ts[i % len(ts)]
        with torch.no_grad():
            e = unet(torch.cat([lat] * 2), t, text, audio, mask)
        e_t, e_ta = e.chunk(2)
        e = e_t + 4.0 * (e_ta - e_t)
        lat[:, :, 1:] = sched.step(e[:, :, 1:], t, lat[:, :, 1:])

    t0 = time.perf_counter()
    one(0)
    first = time.perf_counter() - t0
    n_w = max(0, min(warmup - 1, int(0.25 * budget_s / max(first, 1e-3))))
    for i in range(n_w):
        one(1 + i)
    n_t = max(1, min(steps, int((budget_s - (time.perf_counter() - t0)) / max(first, 1e-3))))
    t1 = time.perf_counter()
    for i in range(n_t):
        one(1 + n_w + i)
    return (time.perf_counter() - t1) / n_t, n_t, kind, cores, 1 + n_w


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    F, h, w, desc = WORKLOADS[args.workload]
    chans = CHANS[args.workload]
    sd = _build_weights(chans)
    sec, n_t, kind, cores, n_w = cpu_reference_steps(sd, chans, F, h, w, args.warmup, args.steps, args.cpu_budget)
    v = 1.0 / sec  # clip-steps/s: the CPU arm steps one clip at a time whatever clips_per_gpu the GPU arm batches
    sample = (f"{n_t} full-size CFG denoise steps (UNet fwd B=2 fp32 + CFG + DDIM) after {n_w} warm-up, "
              f"{cores} host threads; capped by a {args.cpu_budget:.0f}s budget (asked for {args.steps})")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": n_t,
        "warmup": n_w, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": _config(args, desc),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ product arm
def run_product_arm(args):
    import torch.distributed as dist
    from asva_b200 import flops, schedulers, synth
    from avgen.models.unets import AudioUNet3DConditionModel
    from avgen.pipelines.pipeline_audio_cond_animation import AudioCondAnimationPipeline

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (product arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    F, h, w, desc = WORKLOADS[args.workload]
    chans = CHANS[args.workload]
    K, W = args.steps, max(args.warmup, 3)

    sd = _build_weights(chans)
    with torch.device("meta"):  # skip the 1.17 B-parameter default init; the synthetic weights are assigned
        model = AudioUNet3DConditionModel(sample_size=64, cross_attention_dim=768, attention_head_dim=8,
                                          block_out_channels=chans)
    model.load_state_dict(sd, assign=True)
    model.to(dev)
    pipe = AudioCondAnimationPipeline(None, None, model, schedulers.DDIMScheduler(), None, None)
    pipe.set_progress_bar_config(disable=True)
    # `clips` independent clips per rank (1 = BASELINE's configuration; more batch along the token axis in one
    # fused session), CFG-batched branch-major like the pipeline: [text-only of every clip, text+audio of every clip]
    clips = args.clips_per_gpu
    per_clip = [synth.synth_inputs(F=F, h=h, w=w, k=2, seed=123 + rank * clips + c) for c in range(clips)]
    lat = torch.cat([c[0] for c in per_clip])
    text, audio, mask = (torch.cat([torch.stack([c[i][j] for c in per_clip]) for j in range(2)]) for i in (1, 2, 3))
    n_sched = max(50, K + W + 2)
    text_d, audio_d, mask_d = text.to(dev), audio.to(dev), mask.to(dev)
    sess = pipe.open_session(text_d, audio_d, mask_d, F, h, w, n_sched, audio_guidance_scale=4.0)
    be = model.engine().be

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- value: clip resident in HBM
    sess.load_latents(lat.to(dev))
    for i in range(W):
        sess.step(i)
    clocks = ClockSampler(local) if rank == 0 else None
    n0 = sess.launches
    ms = timed(lambda i: sess.step((W + i) % sess.num_steps), K)
    launches = sess.launches - n0
    clk = clocks.stop() if clocks else None
    final = sess.latents.clone()
    finite = bool(torch.isfinite(final).all())

    # ---- e2e: host latents in, host latents out, every step
    host_in = lat.clone().pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    ctx_host = [t.contiguous().pin_memory() for t in (text[:, :1], audio[:, :1], mask)]
    ctx_bytes = sum(t.numel() * t.element_size() for t in ctx_host)

    def e2e_clip(n, first_step):
        tx = ctx_host[0].to(dev, non_blocking=True).expand(-1, F, -1, -1)
        au = ctx_host[1].to(dev, non_blocking=True).expand(-1, F, -1, -1)
        mk = ctx_host[2].to(dev, non_blocking=True)
        s = pipe.open_session(tx, au, mk, F, h, w, n_sched, audio_guidance_scale=4.0)
        for i in range(n):
            s.load_latents(host_in)               # H2D, pinned
            s.step((first_step + i) % s.num_steps)
            s.read_latents(host_out)              # D2H, pinned
            torch.cuda.current_stream().synchronize()
            host_in.copy_(host_out)               # the next step's input is this step's result

    e2e_clip(W, 0)
    host_in.copy_(lat)
    ms_e2e = timed(lambda i: e2e_clip(K, W) if i == 0 else None, 1)
    lat_bytes = host_in.numel() * 4

    # ---- roofline of the dominant kernel family, live: the launches of one step are recorded per kernel family,
    #      each family is re-captured as its own CUDA graph (same launches, same order, same buffers) and its replay
    #      is timed with CUDA events - per-launch durations without host launch overhead in the measurement
    fl = flops.step_flops(2 * clips, F, h, w, chans=chans)
    peaks, peak_src = _peaks()
    os.environ["ASVA_NO_GRAPH"] = "1"
    eager = pipe.__class__(None, None, model, schedulers.DDIMScheduler(), None, None)
    eager.set_progress_bar_config(disable=True)
    pipe._loop = None
    es = eager.open_session(text_d, audio_d, mask_d, F, h, w, n_sched, audio_guidance_scale=4.0)
    es.load_latents(lat.to(dev))
    es.step(0)
    FAMILIES = {"gemm": "gemm", "attention": "attention", "temporal_attention": "temporal_attention",
                "layernorm": "layernorm", "groupnorm": "groupnorm", "groupnorm_stats": "groupnorm",
                "groupnorm_apply": "groupnorm",
                "small_linear": "small_linear", "conv_in_im2col": "misc", "conv_out_finish": "misc", "tconv_gather": "misc",
                "timestep_features": "misc", "cfg_ddim_step": "cfg_step", "cfg_plms_step": "cfg_step"}
    recorded, originals = [], {}
    for name, famname in FAMILIES.items():
        orig = getattr(be, name)
        originals[name] = orig

        def wrapped(*a, _orig=orig, _fam=famname, **k):
            recorded.append((_fam, _orig, a, k))
            return _orig(*a, **k)

        setattr(be, name, wrapped)
    es.step(1)
    for name, orig in originals.items():
        setattr(be, name, orig)
    torch.cuda.synchronize()
    os.environ.pop("ASVA_NO_GRAPH", None)
    fam = {}
    for famname in sorted(set(FAMILIES.values())):
        calls = [(fn, a, k) for f_, fn, a, k in recorded if f_ == famname]
        if not calls:
            continue
        n0 = be.launches
        g = torch.cuda.CUDAGraph()
        gc.collect()  # nothing may be garbage-collected inside the capture window (see engine.GraphRunner)
        gc.disable()
        try:
            with torch.cuda.graph(g):
                for fn, a, k in calls:
                    fn(*a, **k)
        finally:
            gc.enable()
        n_launch = be.launches - n0
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        fam[famname] = (n_launch, e0.elapsed_time(e1) / reps)
        del g
    gemm_calls, gemm_ms = fam.get("gemm", (0, 0.0))
    kernel_ms_total = sum(t for _, t in fam.values())
    step_ms = ms / K
    achieved = fl["gemm"] / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
        except Exception:  # noqa: BLE001
            traffic = None

    # ---- gather the finished latents (the only collective of the path; outside the timed region)
    if world > 1:
        gathered = [torch.empty_like(final) for _ in range(world)]
        dist.all_gather(gathered, final)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, n_t, kind, cores, n_w = cpu_reference_steps(sd, chans, F, h, w, 1, 2, args.cpu_budget_inline)
        cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{n_t} full-size CFG denoise step(s) (reference UNet fwd B=2 fp32 + CFG + DDIM) after "
                         f"{n_w} warm-up on {cores} host threads"}
    if rank == 0:
        value = world * clips * K / (ms * 1e-3)  # clip-steps per second over all clips of all ranks
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": _config(args, desc),
            "notes": {"cuda_graph": os.environ.get("ASVA_NO_GRAPH", "0") != "1",
                      "algorithmic_gflop_per_graph_replay": fl["total"] / 1e9,
                      "value_counts": "clip-steps/s = ranks x clips_per_gpu x graph replays / s"},
            "e2e": {"value": world * clips * K / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": lat_bytes + ctx_bytes // K, "d2h_bytes_per_step": lat_bytes},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "kernel": "gemm_tc_kernel (asva_gemm: linear / implicit-GEMM conv / temporal conv)",
                         "peak_source": f"{peak_src} bf16_tflops_sustained",
                         "gemm_launches_per_step": gemm_calls, "gemm_ms_per_step": gemm_ms,
                         "gemm_share_of_kernel_time": gemm_ms / kernel_ms_total if kernel_ms_total else None,
                         "whole_step_frac": fl["total"] / (step_ms * 1e-3) / 1e12 / peak if peak else None,
                         "family_ms": {k: round(t, 4) for k, (_, t) in sorted(fam.items())}},
            "cpu_baseline": cpu,
            "finite": finite,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--clips-per-gpu", dest="clips_per_gpu", type=int, default=1,
                    help="independent clips batched in one fused session per GPU (BASELINE's configuration is 1)")
    ap.add_argument("--cpu-budget", dest="cpu_budget", type=float, default=200.0,
                    help="seconds of host time the reference arm may spend on timed steps")
    ap.add_argument("--cpu-budget-inline", dest="cpu_budget_inline", type=float, default=45.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_product_arm(args)


if __name__ == "__main__":
    main()

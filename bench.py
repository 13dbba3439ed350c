#!/usr/bin/env python
"""bench.py — motion frames/sec of the DiffuseStyleGesture sampling hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this framework (libdsg, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host cores

One "step" = one pass of the hot path over one batch of synthetic input: B clips per GPU, each a 320-frame
ZEGGS clip = 4 sequential 88-frame segments x 1000 DDPM steps (denoiser + posterior), segment hand-off included
(BASELINE.json configs[1]; reference main/mydiffusion_zeggs/sample.py:236-296).  Weak scaling: every rank runs
B clips; clips are independent, so there is no data-path collective — only the final gather of motions.

`value`  : frames/s with conditioning features already resident in HBM, result left on the device.
`e2e`    : frames/s through the public API (sample.inference_batch) with pinned HOST feature buffers (H2D of
           each segment's features inside the timed region), at N > 1 the ONE gather of all finished motions to
           rank 0 (distributed.gather_motions, from the device buffers), and the D2H read of the result.
`sweep`  : (N = 1) the same step at B in {1, 8, 64, 256, 512} clips (BASELINE config 5);  `config3`: B = 64, DDIM-100,
           six styles (BASELINE config 3);  `strong` (N > 1): 512 clips sharded over the N ranks (strong scaling).
`parity` : clip 0 of the timed run against the committed 1000-step reference golden (tests/golden).
`roofline`: dominant kernel class, algorithmic FLOPs / CUDA-event time measured in a separate profiled pass.
`cpu_baseline`: the oracle port of the reference (torch fp32 CPU, all host threads), bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "motion frames/sec (1000-step DDPM, 320-frame ZEGGS clips)"
UNIT = "frames/s"
N_FRAMES = 320
# Algorithmic FLOPs per denoiser call and clip, as written in the reference
FLOP_PER_CLIP_STEP = 1_315_520_512
FLOP_CLASS = {  # per clip-step, by kernel class (2*M*N*K as written in the reference)
    "gemm_in": 51_408_896 + 25_952_256, "local_attention": 1_982_464,
    "gemm_qkv": 8 * 2 * 89 * 256 * 768, "self_attention": 64_888_832, "gemm_outproj": 8 * 2 * 89 * 256 * 256,
    "gemm_ff1": 8 * 2 * 89 * 256 * 1024, "gemm_ff2": 8 * 2 * 89 * 1024 * 256, "gemm_head_posterior": 51_408_896,
}
POSTERIOR_BYTES_PER_CLIP_STEP = 1_204_896       # read x_t, read x0, write x_{t-1} (fp32), noise generated in-kernel


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '', 1).isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '', 1).isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """CPU threads this process may really use: scheduler affinity capped by the cgroup CPU quota (os.cpu_count()
    reports the host's cores and oversubscribes a quota-limited container)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


def log(msg):
    sys.stderr.write("[bench %.1fs] %s\n" % (time.perf_counter() - T0, msg))
    sys.stderr.flush()


T0 = time.perf_counter()


def cpu_baseline(sample_steps, batch, threads):
    """The reference sampler on the host cores: `sample_steps` DDPM steps of one 88-frame segment at batch `batch`,
    extrapolated linearly to 4 segments x 1000 steps (every step costs the same: same shapes, same kernels).
    kind "reference": the UNMODIFIED reference (`MDM` + `SpacedDiffusion.p_sample_loop`, staged under oracle/_ref by
    oracle/build_ref.py) through its own public API; kind "port": oracle/dsg_oracle.py when no staged copy exists."""
    from diffusestylegesture_b200.config import ZEGGS
    from diffusestylegesture_b200.synthetic import synthetic_state_dict, synthetic_conditioning
    from oracle import build_ref
    torch.set_num_threads(threads)
    g = ZEGGS
    sd = synthetic_state_dict(g, seed=0)
    y = synthetic_conditioning(g, batch, segment=0)
    shape = (batch, g.njoints, 1, g.n_poses)
    if build_ref.available():
        kind = "reference"
        MDM, gd, SpacedDiffusion, space_timesteps = build_ref.load_reference("zeggs")
        with open(os.devnull, "w") as devnull:             # the reference constructor prints its configuration
            old = sys.stdout
            sys.stdout = devnull
            try:
                model = build_ref.make_reference_model(MDM, g, sd)
            finally:
                sys.stdout = old
        diffusion = build_ref.make_reference_diffusion(gd, SpacedDiffusion, space_timesteps)
        yy = {'style': y['style'], 'seed': y['seed'], 'audio': y['audio'], 'mask_local': torch.ones(1, g.n_poses).bool(),
              'mask': (torch.zeros([1, 1, 1, g.n_poses]) < 1)}                      # sample.py:226-234

        def run(nsteps):                                    # the call of sample.py:253-264
            return diffusion.p_sample_loop(model, shape, clip_denoised=False, model_kwargs={'y': yy}, skip_timesteps=1000 - nsteps,
                                           init_image=None, progress=False, dump_steps=None, noise=None, const_noise=False)
        what = "the unmodified reference (oracle/_ref: MDM + SpacedDiffusion.p_sample_loop, torch fp32 CPU"
    else:
        kind = "port"
        from oracle import dsg_oracle as O
        sched = O.Schedule(1000)
        # the reference draws its noise with torch.randn (gaussian_diffusion.py:542): time THAT, not the numpy Philox
        # restatement the parity tests use
        O.noise_tensor = lambda seed, clip_ids, segment, draw, shp: torch.randn((len(clip_ids),) + tuple(shp))

        def run(nsteps):
            return O.p_sample_loop(sd, g, sched, y, batch, skip_timesteps=1000 - nsteps)
        what = "oracle/dsg_oracle.py (port of the reference, torch fp32 CPU"
    with torch.no_grad():
        run(3)                                              # warm-up
        t0 = time.perf_counter()
        run(sample_steps)
        dt = time.perf_counter() - t0
    per_step = dt / sample_steps
    clip_seconds = per_step * 1000 * 4
    return {"value": batch * N_FRAMES / clip_seconds, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{sample_steps} DDPM steps of one 88-frame segment at batch {batch}: {what}, "
                      f"{threads} threads; {dt:.1f} s), extrapolated linearly to 4 segments x 1000 steps",
            "ms_per_denoise_step": per_step * 1e3, "batch": batch}


WAVLM_FLOP_PER_SEGMENT = 162.5e9      # 70,400 samples -> 219 frames: convs 21.6 + pos-conv 3.7 + 24 layers 137 GFLOP


def bench_wavlm(dev, batch, peaks, frames_per_s_per_gpu, pipeline=None):
    """Side measurements (rank 0): the WavLM-Large forward alone, and — `pipeline` = (run_from_features, B, nseg, n_frames) —
    the whole path from raw 16 kHz waveforms in pinned host memory to motions in host memory (WavLM + diffusion)."""
    from diffusestylegesture_b200.wavlm import WavLM
    from diffusestylegesture_b200.wavlm_config import WAVLM_LARGE, synthetic_wavlm_state_dict, synthetic_wav
    log("wavlm: building synthetic WavLM-Large (315 M parameters)")
    m = WavLM(max_batch=batch)
    m.load_state_dict(synthetic_wavlm_state_dict(WAVLM_LARGE, seed=0))
    m.to(dev).eval()
    wav = synthetic_wav(batch, 70400).pin_memory()
    for _ in range(3):
        m.wav2wavlm(wav, 88)
    torch.cuda.synchronize(dev)
    l0 = m.launches
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    a.record()
    for _ in range(K):
        out = m.wav2wavlm(wav, 88)
    b.record()
    torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b) / K
    assert bool(torch.isfinite(out).all())
    ach = WAVLM_FLOP_PER_SEGMENT * batch / (ms * 1e-3) / 1e12
    seg_per_s = batch / (ms * 1e-3)
    res = {"segments_per_call": batch, "ms_per_call": ms, "segments_per_s": seg_per_s, "frames_per_s": seg_per_s * 80,
           "tflops": ach, "frac_of_sustained_bf16": ach / peaks["bf16_tflops_sustained"], "launches_per_call": (m.launches - l0) // K,
           "h2d_bytes_per_call": int(wav.numel() * 4),
           "share_of_clip_time": (frames_per_s_per_gpu / (seg_per_s * 80)) if frames_per_s_per_gpu else None,
           "note": "one 70,400-sample window per 80 new frames; share_of_clip_time = WavLM time / diffusion time for the same frames"}
    log("wavlm: %.2f ms per %d segments, %.0f TFLOP/s" % (ms, batch, ach))
    if pipeline is not None:
        run_from_features, B, nseg, n_frames = pipeline
        # every clip gets its own waveform windows: [nseg][B, 70400] (8 seed frames of the previous window + 80 new frames)
        wavs = [torch.cat([wav] * ((B + batch - 1) // batch))[:B].contiguous().pin_memory() for _ in range(nseg)]

        def once():
            feats = [m.wav2wavlm(w, 88) for w in wavs]          # host -> device inside dsg_wavlm_forward
            return run_from_features(feats)                       # motions come back in host memory

        once()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = once()
        b.record()
        torch.cuda.synchronize(dev)
        ms2 = a.elapsed_time(b)
        res["from_waveform"] = {"value": B * n_frames / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2, "clips": B,
                                "h2d_bytes_per_step": int(sum(w.numel() * 4 for w in wavs)), "d2h_bytes_per_step": int(out.numel() * 4),
                                "note": "one step = raw waveform windows in pinned host memory -> WavLM-Large -> 4 x 1000-step DDPM -> motions in host memory"}
        log("waveform -> motion: %.1f ms per step, %.0f frames/s" % (ms2, res["from_waveform"]["value"]))
    m.close()
    return res


def bench_plus(args):
    """BASELINE config 4: DiffuseStyleGesture+ long-form clips (900 frames = 8 chained segments of 150 frames, stride 120,
    1000-step DDPM) through the BEAT-TWH driver mirror (sample_beat_twh.inference_batch_beat) on the tcgen05 multi-kernel
    path (D = 384 / 512), `--batch` clips per GPU (default 8: batch 32 over 4 GPUs), one gather to rank 0 at N > 1."""
    from diffusestylegesture_b200.distributed import init_from_env, barrier_max_ms, gather_motions
    from diffusestylegesture_b200.config import BEAT_PLUS, TWH_PLUS
    from diffusestylegesture_b200.mdm import MDM
    from diffusestylegesture_b200.model_util import create_gaussian_diffusion, load_model_wo_clip
    from diffusestylegesture_b200.synthetic import synthetic_state_dict
    from diffusestylegesture_b200 import sample_beat_twh as SB
    import torch.distributed as dist
    rank, world, local = init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    g = BEAT_PLUS if args.workload == "beat+" else TWH_PLUS
    dataset = "BEAT" if args.workload == "beat+" else "TWH"
    flop_clip_step = 4_184.4e6 if args.workload == "beat+" else 6_312.1e6
    B = args.batch or 8
    n_frames, steps = 900, args.ddpm_steps
    model = MDM(njoints=g.njoints, cond_mode='cross_local_attention4_style1_sample', audio_feat='wavlm', n_seed=g.n_seed,
                latent_dim=g.latent_dim, style_dim=g.style_in, source_audio_dim=g.audio_dim, audio_feat_dim_latent=g.audio_latent,
                precision="bf16", max_batch=B)
    load_model_wo_clip(model, synthetic_state_dict(g, seed=0))
    model.to(dev).eval()
    diffusion = create_gaussian_diffusion('' if steps == 1000 else [steps])
    cfg = SB.Config(dict(n_poses=g.n_poses, n_seed=g.n_seed, version="v0", name="DiffuseStyleGesture+"))
    gen = torch.Generator().manual_seed(1234 + rank)
    ta_pin = torch.randn(B, n_frames, g.audio_dim, generator=gen).pin_memory()
    ta_dev = ta_pin.to(dev)
    styles = np.zeros((B, g.style_in), dtype=np.float32)
    styles[np.arange(B), (rank * B + np.arange(B)) % g.style_in] = 1
    mean, std = SB.load_stats(dataset)
    seeds = mean + std * np.cumsum(0.05 * np.random.default_rng(3).standard_normal((g.n_seed + 2, g.njoints // 3)), axis=0)
    ids = list(range(rank * B, rank * B + B))
    eng = model.get_engine(B)

    def step(ta, e2e):
        out = SB.inference_batch_beat(cfg, ta, diffusion.p_sample_loop, model, styles, seeds, seed=123456, dataset=dataset,
                                      clip_ids=ids, out_device=dev)
        if e2e:
            allm = gather_motions(out.contiguous(), world * B)
            return allm.cpu() if rank == 0 else None
        return out

    def timed(ta, e2e, K, W):
        for _ in range(W):
            step(ta, e2e)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        l0 = eng.launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(K):
            out = step(ta, e2e)
        b.record()
        torch.cuda.synchronize(dev)
        return barrier_max_ms(a.elapsed_time(b), dev) / K, eng.launches - l0, out

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, launches, out = timed(ta_dev, False, args.steps, args.warmup)
    clk = clocks.stop() if rank == 0 else None
    e_ms, _, out_h = timed(ta_pin, True, max(1, args.steps // 2), 1)
    if rank == 0:
        peaks = load_peaks()
        nseg = 8
        ach = flop_clip_step * B * steps * nseg / (ms * 1e-3) / 1e12
        frames = world * B * n_frames
        line = {"metric": "motion frames/sec (1000-step DDPM, 900-frame DiffuseStyleGesture+ clips)", "value": frames / (ms * 1e-3),
                "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{args.workload} (D={g.latent_dim}, J={g.njoints}, T=150): 900-frame clips = 8 chained segments, "
                                       f"{steps}-step DDPM, {B} clips per GPU, synthetic features [B,900,{g.audio_dim}] and weights",
                           "clips_per_gpu": B, "global_clips": world * B, "segments": nseg, "ddpm_steps": steps, "precision": "bf16",
                           "parallelism": f"clip-dp{world}", "l2": "working set (weights 27-41 MB bf16 + activations) re-streamed every step; "
                                                                    "8000 kernel-graph replays per timed step"},
                "e2e": {"value": frames / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(ta_pin.numel() * 4),
                        "d2h_bytes_per_step": int(out_h.numel() * 4),
                        "collective": None if world == 1 else "one gather of the finished motions to rank 0 inside the timed region"},
                "gpu_launches": int(launches),
                "roofline": {"kernel": "multi-kernel tcgen05 path (44 kernels per DDPM step, CUDA-graph replay)", "bound": "tensor",
                             "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                             "frac": ach / peaks["bf16_tflops_sustained"], "traffic": None,
                             "peak_source": peaks["source"] + ", sustained bf16", "us_per_ddpm_step": ms * 1e3 / (steps * nseg)},
                "cpu_baseline": None, "clocks": clk}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else that lands on fd 1 (NCCL's version banner, library
    prints) was redirected to stderr in main()."""
    os.write(REAL_STDOUT, (json.dumps(line) + "\n").encode())


REAL_STDOUT = 1


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    vals = []
    for _ in range(args.warmup):
        cpu_baseline(max(4, args.ref_sample_steps // 8), args.ref_batch, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_baseline(args.ref_sample_steps, args.ref_batch, threads))
    wall = time.perf_counter() - t0
    v = float(np.mean([c["value"] for c in vals]))
    cb = dict(vals[-1], value=v)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ZEGGS 320-frame clips (4 segments x 88 frames), 1000-step DDPM, the reference sampler on the "
                                   "host cores (" + cb["kind"] + "); each bench step = bounded sample extrapolated",
                       "batch": args.ref_batch, "host_threads": threads},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def main():
    global REAL_STDOUT
    sys.stdout.flush()
    REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dsg", choices=["dsg", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU (0 = default for the precision)")
    ap.add_argument("--precision", default=os.environ.get("DSG_PRECISION", "auto"), choices=["auto", "bf16", "fp32"])
    ap.add_argument("--ddpm-steps", type=int, default=1000, help="diffusion steps per segment (1000 = the metric's config)")
    ap.add_argument("--segments", type=int, default=4)
    ap.add_argument("--ref-sample-steps", type=int, default=240)
    ap.add_argument("--ref-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=40)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank runs --batch clips (default); strong: --global-batch clips are sharded over the ranks")
    ap.add_argument("--global-batch", type=int, default=512, help="total clips of a --scaling strong run")
    ap.add_argument("--workload", default="zeggs", choices=["zeggs", "beat+", "twh+"],
                    help="zeggs = the metric's configuration (default); beat+ / twh+ = BASELINE config 4 (900-frame long-form clips)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the batch sweep / config 3 / strong-scaling side measurements")
    ap.add_argument("--no-wavlm", action="store_true", help="skip the side measurement of the WavLM-Large conditioning forward")
    ap.add_argument("--wavlm-batch", type=int, default=32)
    args = ap.parse_args()

    from diffusestylegesture_b200.distributed import init_from_env, barrier_max_ms, gather_motions
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return

    if args.workload != "zeggs":
        bench_plus(args)
        return
    rank, world, local = init_from_env()
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU fallback (use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    import torch.distributed as dist
    from diffusestylegesture_b200.config import ZEGGS
    from diffusestylegesture_b200.mdm import MDM
    from diffusestylegesture_b200.model_util import create_gaussian_diffusion, load_model_wo_clip
    from diffusestylegesture_b200.synthetic import synthetic_state_dict, synthetic_conditioning
    from diffusestylegesture_b200 import sample as S

    g = ZEGGS
    precision = args.precision
    sd = synthetic_state_dict(g, seed=0)

    def make_model(prec, mb):
        m = MDM(njoints=g.njoints, cond_mode='cross_local_attention3_style1', audio_feat='wavlm', n_seed=g.n_seed,
                precision=prec, max_batch=mb)
        load_model_wo_clip(m, sd)
        m.to(dev).eval()
        m.get_engine(mb)
        return m

    if precision == "auto":
        try:
            make_model("bf16", 1)
            precision = "bf16"
        except NotImplementedError:
            precision = "fp32"
    # bf16: one persistent CTA per clip -> a multiple of the SM count (148) keeps every SM busy for the whole segment
    B = args.batch or (296 if precision == "bf16" else 8)
    clip0 = rank * B
    if args.scaling == "strong":            # contiguous shards of the global batch (distributed.shard_bounds); the slowest rank is timed
        from diffusestylegesture_b200.distributed import shard_bounds
        clip0, hi = shard_bounds(args.global_batch, rank, world)
        B = hi - clip0
        if B <= 0:
            raise SystemExit(f"--global-batch {args.global_batch} leaves rank {rank} without clips")
    total_clips = args.global_batch if args.scaling == "strong" else world * B
    model = make_model(precision, B)
    eng = model.get_engine(B)
    resp = '' if args.ddpm_steps == 1000 else [args.ddpm_steps]
    diffusion = create_gaussian_diffusion(resp)
    nseg = args.segments
    n_frames = nseg * (g.n_poses - g.n_seed)
    clip_ids = list(range(clip0, clip0 + B))
    conds = [synthetic_conditioning(g, B, segment=s, clip_offset=clip0) for s in range(nseg)]
    styles = conds[0]["style"]
    # global clip 0 is the clip of the committed reference golden (tests/golden/inference_zeggs_1000.npz: clip id 0, its
    # synthetic features, seed 123456): it carries that clip's style instead of 0 mod 6, so that the TIMED output itself
    # can be checked against the reference (`parity` in the JSON line)
    gold_path = os.path.join(ROOT, "tests", "golden", "inference_zeggs_1000.npz")
    if rank == 0 and os.path.exists(gold_path):
        styles[0] = torch.from_numpy(np.asarray(np.load(gold_path)["style"], dtype=np.float32))
    feats_dev = [c["audio"].to(dev) for c in conds]
    feats_pin = [c["audio"].pin_memory() for c in conds]
    styles_pin = styles.pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # > 126 MB L2

    out_pin = torch.empty(B, n_frames - g.n_seed, g.njoints, dtype=torch.float32).pin_memory()     # the caller's result buffer

    def one_step(feats, out_device):
        return S.inference_batch(model, diffusion, feats, styles_pin if out_device == "cpu" else styles, seed=123456,
                                 clip_ids=clip_ids, out_device=out_device, out=out_pin if out_device == "cpu" else None)

    out_all_pin = torch.empty(total_clips, n_frames - g.n_seed, g.njoints, dtype=torch.float32).pin_memory() \
        if (world > 1 and rank == 0 and not args.no_e2e) else None

    def e2e_step(feats, _unused):
        """host features -> motions in HOST memory on rank 0.  N = 1: straight into the caller's pinned buffer.  N > 1: every
        rank samples into device memory, ONE gather to rank 0 (NCCL, from the device buffers), rank 0 reads the result."""
        if world == 1:
            return one_step(feats, "cpu")
        loc = S.inference_batch(model, diffusion, feats, styles_pin, seed=123456, clip_ids=clip_ids, out_device=dev)
        allm = gather_motions(loc, total_clips)
        if rank == 0:
            out_all_pin.copy_(allm, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return out_all_pin
        return loc

    def timed(feats, out_device, K, W, step_fn=None):
        step_fn = step_fn or one_step
        for _ in range(W):
            step_fn(feats, out_device)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        total = 0.0
        l0 = eng.launches
        for _ in range(K):
            flush.add_(1.0)                                    # L2 flush between timed iterations (untimed)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = step_fn(feats, out_device)
            b.record()
            torch.cuda.synchronize(dev)
            total += a.elapsed_time(b)
        launches = eng.launches - l0
        if world > 1:
            dist.barrier()
        return barrier_max_ms(total, dev), launches, out

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    log(f"timed region: precision {precision}, {B} clips/GPU, {nseg} segments x {diffusion.num_timesteps} steps, W={args.warmup} K={args.steps}")
    total_ms, launches, out = timed(feats_dev, dev, args.steps, args.warmup)
    log("timed region done: %.1f ms/step" % (total_ms / args.steps))
    clk = clocks.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    frames_all = total_clips * n_frames
    value = frames_all / (ms_per_step * 1e-3)

    e2e = None
    if not args.no_e2e:
        e_ms, _, out_h = timed(feats_pin, "cpu", args.steps, 1, step_fn=e2e_step)
        log("e2e done: %.1f ms/step" % (e_ms / args.steps))
        e2e = {"value": frames_all / (e_ms / args.steps * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(sum(f.numel() * 4 for f in feats_pin) + styles_pin.numel() * 4),
               "d2h_bytes_per_step": int(out_h.numel() * 4) if rank == 0 else 0,
               "collective": None if world == 1 else f"one gather of [{B}, {n_frames - g.n_seed}, {g.njoints}] fp32 per rank to rank 0 "
                                                     f"(NCCL, {world} ranks), inside the timed region; the D2H of all "
                                                     f"{total_clips} motions is rank 0's"}
        assert rank != 0 or out_h.shape[0] == total_clips

    # ---- parity tie: clip 0 of the run just timed (global clip id 0, seed 123456, style 0, synthetic features) IS the clip of
    # tests/golden/inference_zeggs_1000.npz (the reference's own 4-segment x 1000-step run): its noise is keyed by clip id, so
    # it must come out the same whatever the batch around it was.  Compared in normalised units.
    parity = None
    if rank == 0 and nseg == 4 and diffusion.num_timesteps == 1000:
        gp = os.path.join(ROOT, "tests", "golden", "inference_zeggs_1000.npz")
        sp = os.path.join(ROOT, "tests", "golden", "zeggs_mean_std.npz")
        if os.path.exists(gp) and os.path.exists(sp):
            gold, st = np.load(gp), np.load(sp)
            if np.array_equal(np.asarray(gold["style"], dtype=np.float32), styles[0].numpy()):
                std = np.clip(np.array(st["std"]).squeeze(), 0.01, None)
                want = (gold["poses"] - np.array(st["mean"]).squeeze()) / std
                got = out[0].float().cpu().numpy()
                err = np.abs(got - want)
                tol = (0.15, 0.02) if precision == "bf16" else (5e-3, 5e-4)
                parity = {"clip0_vs_reference_golden": {"max": float(err.max()), "rms": float(np.sqrt((err ** 2).mean())),
                                                        "tolerance_max_rms": list(tol), "units": "normalised motion (|x| <~ 3)",
                                                        "golden": "tests/golden/inference_zeggs_1000.npz (reference sample.inference)"}}
                parity["ok"] = bool(err.max() < tol[0] and np.sqrt((err ** 2).mean()) < tol[1])
                log("parity: clip 0 vs reference golden: max %.4g rms %.4g (%s)" % (err.max(), np.sqrt((err ** 2).mean()), parity["ok"]))
                if not np.isfinite(err).all() or err.max() > 1.0:
                    raise SystemExit("bench: clip 0 of the timed run does not match the reference golden: %r" % (parity,))

    # ---- BASELINE config 5 (batch sweep, N = 1), config 3 (B = 64, DDIM-100, six styles), strong scaling (N > 1)
    def quick(Bq, diff, sampler="ddpm", offset=0):
        """one warm-up + one timed step of Bq clips on this rank (device-resident features)."""
        cq = [synthetic_conditioning(g, Bq, segment=s_, clip_offset=offset) for s_ in range(nseg)]
        fq = [c["audio"].to(dev) for c in cq]
        ids = list(range(offset, offset + Bq))

        def run():
            return S.inference_batch(model, diff, fq, cq[0]["style"], seed=123456, clip_ids=ids, out_device=dev, sampler=sampler)
        run()
        torch.cuda.synchronize(dev)
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        o = run()
        b.record()
        torch.cuda.synchronize(dev)
        assert bool(torch.isfinite(o).all())
        return a.elapsed_time(b)

    sweep, config3, strong = None, None, None
    if precision == "bf16" and not args.no_sweep:
        if world == 1:
            sweep = []
            for Bq in (1, 8, 64, 256, 512):
                ms = quick(Bq, diffusion)                     # (the engine is rebuilt for a larger workspace when Bq > B)
                sweep.append({"clips": Bq, "ms_per_step": ms, "value": Bq * n_frames / (ms * 1e-3), "unit": UNIT})
                log("sweep: B=%d  %.1f ms  %.0f frames/s" % (Bq, ms, sweep[-1]["value"]))
            d3 = create_gaussian_diffusion("ddim100")
            ms = quick(64, d3, sampler="ddim")
            config3 = {"workload": "batch 64 ZEGGS 320-frame clips, styles i mod 6, DDIM-100 (eta 0), 4 segments", "clips": 64,
                       "ms_per_step": ms, "value": 64 * n_frames / (ms * 1e-3), "unit": UNIT,
                       "parity": "tests/test_gpu_r2.py::test_config3_b64_ddim100_six_styles_vs_reference_golden"}
            log("config 3: %.1f ms, %.0f frames/s" % (ms, config3["value"]))
            eng = model.get_engine(B)                        # the live engine (rebuilt by the 512-clip point)
        else:
            G_ = 512
            from diffusestylegesture_b200.distributed import shard_bounds
            lo, hi = shard_bounds(G_, rank, world)
            dist.barrier()
            ms = barrier_max_ms(quick(hi - lo, diffusion, offset=lo), dev)
            strong = {"global_clips": G_, "clips_per_gpu": hi - lo, "ms_per_step": ms, "value": G_ * n_frames / (ms * 1e-3), "unit": UNIT,
                      "scaling": "strong", "note": "512 clips sharded contiguously over the ranks (max over ranks), device-resident"}

    # ---- roofline.  The bf16 engine runs a segment as ONE persistent kernel (csrc/dsg_clip_kernel.cuh), so "the
    # dominant kernel" is that kernel: its launch is timed live with CUDA events on the launching stream, and
    # achieved = algorithmic denoiser FLOPs of the launch (1.3155 GFLOP x clips x steps, as written in the
    # reference) / duration.  DSG_TC_MODE=kernels / fp32: per-kernel-class event timing of the multi-kernel path.
    roofline, kernels = None, None
    peaks = load_peaks() if rank == 0 else None
    if rank == 0:
        y = dict(conds[0], audio=feats_dev[0], noise_seed=123456, segment=0, clip_ids=clip_ids)
        shape = (B, g.njoints, 1, g.n_poses)
        clip_mode = precision == "bf16" and os.environ.get("DSG_TC_MODE", "clip") != "kernels"
        if clip_mode:
            l0 = eng.launches
            diffusion.p_sample_loop(model, shape, clip_denoised=False, model_kwargs={'y': y})      # warm
            nl = eng.launches - l0
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            a.record()
            diffusion.p_sample_loop(model, shape, clip_denoised=False, model_kwargs={'y': y})
            b.record()
            torch.cuda.synchronize(dev)
            ms = a.elapsed_time(b)
            flop = FLOP_PER_CLIP_STEP * B * diffusion.num_timesteps
            ach = flop / (ms * 1e-3) / 1e12
            peak = peaks["bf16_tflops_sustained"]
            roofline = {"kernel": "clip_kernel (persistent: input GEMM + local attention + 8 layers + head + posterior, all steps of a segment)",
                        "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                        "peak_source": peaks["source"] + ", sustained bf16", "flop_per_launch": flop, "avg_launch_us": ms * 1e3,
                        "launches_per_segment": nl, "us_per_ddpm_step": ms * 1e3 / diffusion.num_timesteps,
                        "note": "segment call timed with CUDA events (conditioning GEMMs + x_T draw + the loop kernel); "
                                "FLOPs = reference's 2*M*N*K count, excludes padding (89 -> 128 token rows) and hoisted terms"}
            # HBM traffic of the launch: x_t read and written once (fp32), its bf16 k-block image written and read once, the
            # pre-drawn noise written and read once per clip-step (everything else is on chip or L2-resident weights): ncu
            # measured 1.95 MB per clip-step (profiles/r01_clip_kernel_v6_ncu_full_summary.csv: dram read + write of a
            # 148-clip x 12-step launch = 3.458 GB)
            # (an ncu --set full capture of this kernel, summarised by profiles/summarise_ncu.py into profiles/clip_kernel_traffic.json)
            tp = os.path.join(ROOT, "profiles", "clip_kernel_traffic.json")
            if os.path.exists(tp):
                tj = json.load(open(tp))
                roofline["traffic"] = tj["dram_bytes_per_clip_step"] * B * diffusion.num_timesteps
                roofline["traffic_note"] = ("bytes per launch = ncu dram__bytes_read.sum + dram__bytes_write.sum per clip-step (%.3g MB, %s) "
                                            "x clips x steps of this launch" % (tj["dram_bytes_per_clip_step"] / 1e6, tj["source"]))
            log("clip-kernel segment: %.1f ms (%.1f us per DDPM step), %.1f TFLOP/s" % (ms, ms * 1e3 / diffusion.num_timesteps, ach))
            try:      # where the persistent kernel spends its cycles (instrumented build of the same kernel, 50 steps)
                os.environ["DSG_CLIP_PROF"] = "1"
                d50 = create_gaussian_diffusion([50])
                d50.p_sample_loop(model, shape, clip_denoised=False, model_kwargs={'y': y})
                torch.cuda.synchronize(dev)
                ph = eng.clip_profile()
                mhz = (clk or {}).get("sm_mhz") or 1965.0
                kernels = {"clip_kernel_phases_us_per_step": {k: v / 50 / mhz for k, v in ph.items()},
                           "note": "clock64() deltas of CTA 0 (MMA thread: total / waiting for weights / waiting for the workers; "
                                   "worker thread 0: time per epilogue phase), instrumented build, 50 steps"}
            except (RuntimeError, NotImplementedError) as ex:
                sys.stderr.write(f"clip profile unavailable: {ex}\n")
            finally:
                os.environ.pop("DSG_CLIP_PROF", None)
        else:
            psteps = min(args.profile_steps, diffusion.num_timesteps - 1)
            try:
                eng.profile(True)
                diffusion.p_sample_loop(model, shape, clip_denoised=False, model_kwargs={'y': y},
                                        skip_timesteps=diffusion.num_timesteps - psteps)
                prof = eng.profile_read()
                eng.profile(False)
                log("profile pass done")
            except (RuntimeError, NotImplementedError) as ex:
                prof = {}
                sys.stderr.write(f"profile pass unavailable: {ex}\n")
            if prof:
                tot = sum(ms for _, ms in prof.values())
                kernels = {k: {"launches": n, "avg_us": ms / n * 1e3, "share": ms / tot} for k, (n, ms) in prof.items()}
                gemm = {k: v for k, v in prof.items() if k in FLOP_CLASS}
                dom = max(gemm, key=lambda k: gemm[k][1]) if gemm else None
                if dom:
                    n, ms = prof[dom]
                    per_launch_flop = FLOP_CLASS[dom] * B / (8 if dom in ("gemm_qkv", "gemm_outproj", "gemm_ff1", "gemm_ff2", "self_attention") else 1)
                    ach = per_launch_flop / (ms / n * 1e-3) / 1e12
                    peak = peaks["bf16_tflops_sustained"]
                    roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                                "traffic": None, "peak_source": peaks["source"] + ", sustained bf16",
                                "flop_per_launch": per_launch_flop, "avg_launch_us": ms / n * 1e3}
                if "posterior" in prof:
                    n, ms = prof["posterior"]
                    gbs = POSTERIOR_BYTES_PER_CLIP_STEP * B / (ms / n * 1e-3) / 1e9
                    kernels["posterior"]["hbm_gbs"] = gbs
                    kernels["posterior"]["hbm_frac_of_measured_peak"] = gbs / peaks["hbm_gbs"]

    # ---- side measurement (not part of `value`): the WavLM-Large conditioning forward of the same clips, raw 16 kHz
    # waveform in pinned host memory -> [segments, 88, 1024] features on the device, through dsg_wavlm_forward
    wavlm = None
    if rank == 0 and not args.no_wavlm and precision == "bf16":
        try:
            pipe = None if args.no_e2e else ((lambda feats: one_step(feats, "cpu")), B, nseg, n_frames)
            wavlm = bench_wavlm(dev, args.wavlm_batch, peaks, value / world, pipe)
        except (RuntimeError, NotImplementedError) as ex:
            sys.stderr.write(f"wavlm measurement unavailable: {ex}\n")

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        log("cpu baseline (oracle port, %d threads)" % host_threads())
        cb = cpu_baseline(args.ref_sample_steps, args.ref_batch, host_threads())
        log("cpu baseline done")

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": f"ZEGGS {n_frames}-frame clips = {nseg} sequential segments x 88 frames, "
                                       f"{diffusion.num_timesteps}-step DDPM, {B} clips per GPU, WavLM-shaped synthetic features "
                                       "[B,88,1024] per segment, synthetic weights (9.0 M params)",
                           "clips_per_gpu": B, "global_clips": total_clips, "segments": nseg, "ddpm_steps": diffusion.num_timesteps,
                           "precision": precision, "parallelism": f"clip-dp{world}", "l2": "flushed between timed iterations (256 MB write)"},
                "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels, "wavlm": wavlm,
                "cpu_baseline": cb, "clocks": clk, "parity": parity, "sweep": sweep, "config3": config3, "strong": strong}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

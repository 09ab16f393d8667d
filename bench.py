#!/usr/bin/env python
"""Benchmark of the PPG -> Mel -> WaveGlow inference path (contract: see DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision ...]

One "step" = one WaveGlow.infer() over the resident batch of BASELINE.json configs[1]
(mel (8, 80, 1379) -> audio (8, 220640): 8 x 10 s @ 22.05 kHz, fp32) per GPU.
N > 1 (torchrun): weak scaling -- every rank holds its own 8 x 10 s shard, weights are
packed on rank 0 and shipped with ONE NCCL broadcast, no collective in the data path.
`value` = audio samples/s with the mel already in HBM; `e2e` = the same through host
buffers (pinned H2D of the mel, D2H of the waveform inside the timed region).
`--impl reference` times the CPU restatement of the reference (oracle/, torch CPU fp32 on
all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from fac_via_ppg_b200 import synth  # noqa: E402

RATE = 22050
WG_MAC_PER_COLUMN = 81_358_288          # SURVEY.md section 8d / BASELINE.md section 4
WG_FLOP_PER_SAMPLE = 2 * WG_MAC_PER_COLUMN / 8
WG_HBM_BYTES_PER_SAMPLE = 10            # 2 mel + 4 noise + 4 audio (fp32), BASELINE.md section 4


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"],
                "tflops_sustained": p["bf16_tflops_sustained"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median over the busier half of the samples (the loop is always under load, the
        # first/last samples may straddle its edges)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def build_waveglow(device, precision):
    from fac_via_ppg_b200.waveglow.glow import WaveGlow
    cfg = synth.WAVEGLOW_CONFIG
    model = WaveGlow.remove_weightnorm(WaveGlow(**cfg))
    model.load_state_dict(synth.waveglow_state(cfg=cfg))
    model = model.to(device).eval()
    if hasattr(model, "set_precision"):
        model.set_precision(precision)
    return model


def measure_ppg2mel(dev, batch=8, frames=690):
    """Side metric of BASELINE.json ('mel frames/sec'): Tacotron2.inference on synthetic PPGs
    (batch x 5816 x frames, forced decode length), fp32-grade CUDA path (split-fp16 tensor-core GEMMs with
    fp32 accumulation, fp32 recurrences), CUDA-event phase times."""
    from fac_via_ppg_b200.common.hparams import create_hparams_stage
    from fac_via_ppg_b200.common.model import Tacotron2
    model = Tacotron2(create_hparams_stage())
    model.load_state_dict(synth.tacotron_state())
    model = model.to(dev).eval()
    model.decoder.gate_threshold, model.decoder.max_decoder_steps = 2.0, frames
    model.collect_timing, model.return_alignments = True, False
    ppg = synth.synthetic_ppg(batch, frames).to(dev)
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.inference(ppg)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    tm = model.last_timing
    return {"metric": "mel frames/sec (Tacotron2.inference PPG->Mel)", "value": batch * frames / best,
            "unit": "frames/s", "batch": batch, "frames": frames,
            "dtype": "f32 results; GEMM operands split into fp16 hi/lo pairs (3 tensor-core products each)",
            "decoder_us_per_step": tm["decoder_ms"] * 1e3 / frames, "encoder_ms": tm["encoder_ms"],
            "decoder_ms": tm["decoder_ms"], "postnet_ms": tm["postnet_ms"],
            "hbm_compulsory_gbs": batch * frames * 5816 * 4 / best / 1e9}


def measure_pipeline(dev, wg, batch=32, seconds=5.0):
    """Side metric: BASELINE.json configs[2], the full PPG -> Mel -> WaveGlow pipeline of generate_synthesis.py on
    32 x 5 s with the bf16 vocoder.  Resident = PPG already in HBM; e2e = pinned host PPG in, host waveform out."""
    from fac_via_ppg_b200.common.hparams import create_hparams_stage
    from fac_via_ppg_b200.common.model import Tacotron2
    taco = Tacotron2(create_hparams_stage())
    taco.load_state_dict(synth.tacotron_state())
    taco = taco.to(dev).eval()
    frames = synth.frames_for_seconds(seconds)
    taco.decoder.gate_threshold, taco.decoder.max_decoder_steps, taco.return_alignments = 2.0, frames, False
    old = wg.precision
    wg.set_precision("bf16")
    ppg_host = synth.synthetic_ppg(batch, frames).pin_memory()
    ppg = ppg_host.to(dev)
    out_host = torch.empty(batch, frames * 160).pin_memory()

    def run(host):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x = ppg_host.to(dev, non_blocking=True) if host else ppg
        mel = taco.inference(x)[1]
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        wav = wg.infer(mel.clamp(-11.5, 2.0).contiguous(), 0.6)
        if host:
            out_host.copy_(wav, non_blocking=True)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, t1 - t0

    try:
        import contextlib
        import io
        with contextlib.redirect_stderr(io.StringIO()):      # "Reached max decoder steps" (forced length)
            run(False)
            res = min(run(False) for _ in range(3))
            e2e = min(run(True) for _ in range(3))
    finally:
        wg.set_precision(old)
    n = batch * frames * 160
    return {"workload": "PPG->Mel->WaveGlow, batch=%dx%.0f s, bf16 vocoder, fp32-grade acoustic model (BASELINE configs[2])"
                        % (batch, seconds),
            "value": n / res[0], "unit": "samples/s", "rtf": n / res[0] / RATE, "ms": res[0] * 1e3,
            "ppg2mel_ms": res[1] * 1e3, "mel2wav_ms": (res[0] - res[1]) * 1e3,
            "e2e": {"value": n / e2e[0], "unit": "samples/s", "ms": e2e[0] * 1e3,
                    "h2d_bytes_per_step": ppg_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4}}


def cpu_port_samples_per_s(frames, repeats, threads):
    """The reference's CPU implementation of the step (oracle restatement, torch CPU fp32)."""
    from oracle import waveglow_oracle   # the one place bench.py executes oracle/: the CPU baseline
    torch.set_num_threads(threads)
    cfg = synth.WAVEGLOW_CONFIG
    sd = synth.waveglow_state(cfg=cfg)
    mel = synth.synthetic_mel(1, frames)
    best = None
    with torch.no_grad():
        for _ in range(repeats + 1):            # first pass is the warm-up
            t0 = time.perf_counter()
            out = waveglow_oracle.waveglow_infer(sd, cfg, mel, 0.6)
            dt = time.perf_counter() - t0
            if _ > 0:
                best = dt if best is None else min(best, dt)
    return out.numel() / best, best


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    frames = synth.frames_for_seconds(args.seconds)   # one full-length utterance of the config per step
    cfg = synth.WAVEGLOW_CONFIG
    from oracle import waveglow_oracle
    torch.set_num_threads(threads)
    sd = synth.waveglow_state(cfg=cfg)
    mel = synth.synthetic_mel(1, frames)
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 1))):
            waveglow_oracle.waveglow_infer(sd, cfg, mel, 0.6)
        t0 = time.perf_counter()
        n = 0
        for _ in range(args.steps):
            n += waveglow_oracle.waveglow_infer(sd, cfg, mel, 0.6).numel()
        dt = time.perf_counter() - t0
    value = n / dt
    sample = "1 x %d frames (%.2f s of audio) per step, oracle port of glow.py on CPU fp32" % (frames, frames * 160 / RATE)
    line = {
        "impl": "reference", "metric": "audio samples/sec @22.05 kHz (WaveGlow.infer mel->wav)", "value": value,
        "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "rtf": value / RATE,
        "config": dict(workload_config(args, args.batch, synth.frames_for_seconds(args.seconds)), precision="fp32 (torch CPU)"),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, batch, frames):
    return {"workload": "WaveGlow.infer mel->wav, batch=%dx10 s @22.05 kHz per GPU (BASELINE configs[1])" % batch,
            "batch_per_gpu": batch, "frames": frames, "samples_per_utterance": frames * 160,
            "sigma": 0.6, "precision": args.precision,
            "l2": "working set per step (>1 GB of activations) exceeds the 126 MB L2; no explicit flush",
            "parallelism": "utterance sharding, dp%d, one NCCL weight broadcast" % args.gpus}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ppg2mel", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: this package has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from fac_via_ppg_b200 import _ext
    lib = _ext.load()
    args.warmup = max(args.warmup, 3)

    # ---- weights: packed on rank 0, one NCCL broadcast of the flat buffer -------------
    model = build_waveglow(dev, args.precision)
    if distributed:
        from fac_via_ppg_b200 import dist as fdist
        fdist.broadcast_packed(model, src=0)
    precision = getattr(model, "precision", "fp32")
    args.precision = precision

    B, F = args.batch, synth.frames_for_seconds(args.seconds)
    samples_per_step = B * F * 160
    mel_host = synth.synthetic_mel(B, F, seed=synth.SEED + rank).pin_memory()
    mel = mel_host.to(dev)
    out_host = torch.empty(B, F * 160).pin_memory()
    torch.manual_seed(synth.SEED + rank)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if distributed:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    def step_resident():
        model.infer(mel, sigma=0.6)

    def step_e2e():
        m = mel_host.to(dev, non_blocking=True)
        audio = model.infer(m, sigma=0.6)
        out_host.copy_(audio, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.fac_reset_launch_count()
    torch.cuda.nvtx.range_push("fac_timed")      # ncu --nvtx --nvtx-include "fac_timed/" isolates this region
    ms = timed(step_resident, args.steps)
    torch.cuda.nvtx.range_pop()
    launches = lib.fac_launch_count()
    clocks = sampler.stop() if rank == 0 else None

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    value = world * samples_per_step * args.steps / (ms / 1e3)
    e2e_value = world * samples_per_step * args.steps / (ms_e2e / 1e3)

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the WN layer GEMMs), timed live with CUDA events
    peaks = measured_peaks()
    roof = model.profile_dominant_kernel(mel, peaks) if hasattr(model, "profile_dominant_kernel") else None
    traffic_path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if roof is not None and precision == "bf16x3" and os.path.isfile(traffic_path):
        with open(traffic_path) as fh:
            t = json.load(fh)
        roof["traffic"] = t["wn_gemm_tc_kernel"]["launch_weighted_mean_bytes"]     # DRAM bytes per launch (ncu)
        roof["traffic_source"] = t["source"]

    # ---- PPG -> Mel side metric (mel frames/s), short and outside the timed region ----------
    ppg2mel = pipeline = None
    if world == 1 and not args.no_ppg2mel:
        ppg2mel = measure_ppg2mel(dev)
        pipeline = measure_pipeline(dev, model)

    # ---- CPU baseline (bounded sample of the same workload, rank 0 only) ----------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        frames = synth.frames_for_seconds(args.seconds)   # one full-length utterance of the config
        v, secs = cpu_port_samples_per_s(frames, repeats=2, threads=threads)
        cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": "WaveGlow.infer on 1 x %d frames (%d samples), oracle port of reference glow.py, torch CPU "
                         "fp32, best of 2 after 1 warm-up (%.1f s each)" % (frames, frames * 160, secs)}

    line = {
        "metric": "audio samples/sec @22.05 kHz (WaveGlow.infer mel->wav)",
        "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (split-bf16 tensor cores, fp32 accumulate)",
                  "bf16": "bf16"}.get(precision, precision),
        "data": "synthetic (seeded random-init weights, synth.py)",
        "rtf": value / RATE,
        "config": workload_config(args, B, F),
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": mel_host.numel() * 4,
                "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "ppg2mel": ppg2mel,
        "pipeline": pipeline,
        "tflops_algorithmic": value * WG_FLOP_PER_SAMPLE / 1e12,
        "hbm": {"compulsory_bytes_per_sample": WG_HBM_BYTES_PER_SAMPLE,
                "achieved_gbs": value * WG_HBM_BYTES_PER_SAMPLE / 1e9 / world,
                "frac_of_measured_peak": value * WG_HBM_BYTES_PER_SAMPLE / 1e9 / world / peaks["hbm_gbs"],
                "note": "dense contraction: HBM is not the binding resource (BASELINE.md section 4)"},
    }
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

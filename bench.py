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

Side objects on the same JSON line (each timed with CUDA events after its own warm-up, each with its own
nvidia-smi clock sample; none of them is inside the headline's timed region):
  N = 1 : `fp32` (the exact-fp32 FFMA mode on the headline workload), `library_gpu` (the reference's arithmetic
          as stock PyTorch/cuDNN calls on the same B200, TF32 off and on -- SURVEY.md section 8d),
          `ppg2mel` (Tacotron2.inference, mel frames/s), `pipeline` (BASELINE configs[2]), `denoiser`, `cpu_baseline`.
  every N: `cfg4` (BASELINE configs[3]: WaveGlow.infer on 32 x 10 s per GPU = 256 utterances on 8 GPUs) and
          `cfg5` (BASELINE configs[4]: 8 x 60 s per GPU = 64 utterances on 8 GPUs, PPG -> Mel -> WaveGlow ->
          Denoiser as generate_synthesis.py:86-98 runs them, through pinned host buffers, whole-job RTF).
"""
from __future__ import annotations

import argparse
import contextlib
import io
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
FFMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # 74.4: the exact-fp32 mode's own pipe


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
        return self

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


class Env:
    """Rank / device / collective helpers shared by every leg."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.distributed = self.world > 1
        self.dev = None

    def init_cuda(self):
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.distributed:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.distributed:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """Device time of `steps` calls of fn (CUDA events on the launching stream, barrier + synchronize on both
        sides), max over ranks."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.distributed:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    def timed_with_clocks(self, fn, steps, warmup=1):
        for _ in range(warmup):
            fn()
        sampler = ClockSampler(self.local_rank).start() if self.rank == 0 else None
        ms = self.timed(fn, steps)
        return ms, (sampler.stop() if sampler else None)


def build_waveglow(device, precision):
    from fac_via_ppg_b200.waveglow.glow import WaveGlow
    cfg = synth.WAVEGLOW_CONFIG
    model = WaveGlow.remove_weightnorm(WaveGlow(**cfg))
    model.load_state_dict(synth.waveglow_state(cfg=cfg))
    model = model.to(device).eval()
    if hasattr(model, "set_precision"):
        model.set_precision(precision)
    return model


def build_tacotron(device, frames):
    from fac_via_ppg_b200.common.hparams import create_hparams_stage
    from fac_via_ppg_b200.common.model import Tacotron2
    model = Tacotron2(create_hparams_stage())
    model.load_state_dict(synth.tacotron_state())
    model = model.to(device).eval()
    model.decoder.gate_threshold, model.decoder.max_decoder_steps = 2.0, frames     # forced decode length
    model.return_alignments = False
    return model


@contextlib.contextmanager
def quiet_stderr():
    with contextlib.redirect_stderr(io.StringIO()):      # "Reached max decoder steps" (forced length)
        yield


def measure_ppg2mel(env, batch=8, frames=690):
    """Side metric of BASELINE.json ('mel frames/sec'): Tacotron2.inference on synthetic PPGs
    (batch x 5816 x frames, forced decode length), fp32-grade CUDA path (split-fp16 tensor-core GEMMs with
    fp32 accumulation, fp32 recurrences), CUDA-event phase times."""
    model = build_tacotron(env.dev, frames)
    model.collect_timing = True
    ppg = synth.synthetic_ppg(batch, frames).to(env.dev)
    with quiet_stderr():
        ms, clocks = env.timed_with_clocks(lambda: model.inference(ppg), 3, warmup=1)
    ms /= 3
    tm = model.last_timing
    return {"metric": "mel frames/sec (Tacotron2.inference PPG->Mel)", "value": batch * frames / (ms / 1e3),
            "unit": "frames/s", "batch": batch, "frames": frames, "ms": ms, "timing": "CUDA events, mean of 3 after 1 warm-up",
            "dtype": "f32 results; GEMM operands split into fp16 hi/lo pairs (3 tensor-core products each)",
            "decoder_us_per_step": tm["decoder_ms"] * 1e3 / frames, "encoder_ms": tm["encoder_ms"],
            "decoder_ms": tm["decoder_ms"], "postnet_ms": tm["postnet_ms"],
            "hbm_compulsory_gbs": batch * frames * 5816 * 4 / (ms / 1e3) / 1e9, "clocks": clocks}


def measure_pipeline(env, wg, batch=32, seconds=5.0):
    """Side metric: BASELINE.json configs[2], the full PPG -> Mel -> WaveGlow pipeline of generate_synthesis.py on
    32 x 5 s with the bf16 vocoder.  Resident = PPG already in HBM; e2e = pinned host PPG in, host waveform out."""
    dev = env.dev
    frames = synth.frames_for_seconds(seconds)
    taco = build_tacotron(dev, frames)
    old = wg.precision
    wg.set_precision("bf16")
    ppg_host = synth.synthetic_ppg(batch, frames).pin_memory()
    ppg = ppg_host.to(dev)
    out_host = torch.empty(batch, frames * 160).pin_memory()
    mid = torch.cuda.Event(enable_timing=True)
    stamps = []
    # e2e: like measure_cfg5, the PPG upload of call k + 1 runs on a copy stream behind the vocoder of call k
    copy_stream = torch.cuda.Stream(device=dev)
    ppg_dev = [ppg, torch.empty_like(ppg)]
    uploaded = [torch.cuda.Event() for _ in range(2)]
    state = {"k": 0}

    def upload(b):
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            ppg_dev[b].copy_(ppg_host, non_blocking=True)
            uploaded[b].record(copy_stream)

    def run(host):
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        x = ppg
        if host:
            b = state["k"] & 1
            state["k"] += 1
            torch.cuda.current_stream().wait_event(uploaded[b])
            x = ppg_dev[b]
        mel = taco.inference(x)[1]
        mid.record()
        if host:
            upload(b ^ 1)
        wav = wg.infer(mel.clamp(-11.5, 2.0).contiguous(), 0.6)
        if host:
            out_host.copy_(wav, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        stamps.append((e0, mid))

    try:
        with quiet_stderr():
            ms, clocks = env.timed_with_clocks(lambda: run(False), 3, warmup=1)
            torch.cuda.synchronize()
            ppg2mel_ms = stamps[-1][0].elapsed_time(stamps[-1][1])
            upload(0)
            ms_e2e, _ = env.timed_with_clocks(lambda: run(True), 3, warmup=1)
            torch.cuda.synchronize()
    finally:
        wg.set_precision(old)
    ms, ms_e2e = ms / 3, ms_e2e / 3
    n = batch * frames * 160
    return {"workload": "PPG->Mel->WaveGlow, batch=%dx%.0f s, bf16 vocoder, fp32-grade acoustic model (BASELINE configs[2])"
                        % (batch, seconds),
            "value": n / (ms / 1e3), "unit": "samples/s", "rtf": n / (ms / 1e3) / RATE, "ms": ms,
            "timing": "CUDA events, mean of 3 after 1 warm-up",
            "ppg2mel_ms": ppg2mel_ms, "mel2wav_ms": ms - ppg2mel_ms, "clocks": clocks,
            "e2e": {"value": n / (ms_e2e / 1e3), "unit": "samples/s", "ms": ms_e2e,
                    "h2d_bytes_per_step": ppg_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
                    "note": "the PPG upload of call k+1 runs on a copy stream behind the vocoder of call k"}}


def measure_fp32_mode(env, model, mel, peaks, samples_per_step):
    """The exact-fp32 (FFMA) mode on the headline workload: the precision BASELINE configs[1] literally names."""
    old = model.precision
    model.set_precision("fp32")
    try:
        ms, clocks = env.timed_with_clocks(lambda: model.infer(mel, sigma=0.6), 2, warmup=1)
        roof = model.profile_dominant_kernel(mel, peaks, passes=1)
    finally:
        model.set_precision(old)
    ms /= 2
    tfl = roof["achieved"]
    return {"workload": "WaveGlow.infer 8 x 10 s, exact fp32 on the FFMA pipe (no tensor cores)", "dtype": "f32",
            "value": samples_per_step / (ms / 1e3), "unit": "samples/s", "rtf": samples_per_step / (ms / 1e3) / RATE,
            "ms_per_step": ms, "timing": "CUDA events, mean of 2 after 1 warm-up", "clocks": clocks,
            "roofline": {"bound": "ffma", "kernel": "conv_gemm_f32_kernel", "achieved": tfl, "peak": FFMA_PEAK_TFLOPS,
                         "unit": "TFLOP/s", "frac_of_ffma_peak": tfl / FFMA_PEAK_TFLOPS,
                         "peak_source": "148 SM x 128 lanes x 2 x 1.965 GHz (nominal; MEASURED_PEAKS.json has no fp32 entry)",
                         "launches_timed": roof["launches_timed"], "avg_launch_ms": roof["avg_launch_ms"]}}


def measure_library_gpu(env, mel, samples_per_step):
    """SURVEY.md section 8d 'library' row: the reference's WaveGlow.infer arithmetic as stock PyTorch / cuDNN /
    cuBLAS calls on the same B200 (the oracle restatement of glow.py run on `cuda`), TF32 off (parity-grade fp32)
    and TF32 on (what stock PyTorch does by default for convolutions).  A comparator, never the product path."""
    from oracle import waveglow_oracle      # comparator only (same role as the cpu_baseline leg)
    cfg = synth.WAVEGLOW_CONFIG
    sd = {k: v.to(env.dev) for k, v in synth.waveglow_state(cfg=cfg).items()}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    out = {"what": "oracle/waveglow_oracle.waveglow_infer (torch conv1d / conv_transpose1d, fp32 tensors) on cuda, "
                   "same 8 x 10 s workload, CUDA events, mean of 2 after 1 warm-up", "unit": "samples/s"}
    try:
        with torch.no_grad():
            for name, flag in (("tf32_off", False), ("tf32_on", True)):
                torch.backends.cudnn.allow_tf32 = flag
                torch.backends.cuda.matmul.allow_tf32 = flag
                ms, clocks = env.timed_with_clocks(lambda: waveglow_oracle.waveglow_infer(sd, cfg, mel, 0.6), 2, warmup=1)
                ms /= 2
                out[name] = {"value": samples_per_step / (ms / 1e3), "ms_per_step": ms,
                             "rtf": samples_per_step / (ms / 1e3) / RATE, "clocks": clocks}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return out


def measure_denoiser(env, model, batch=8, seconds=10.0):
    """The step right after WaveGlow.infer on the CLI path (reference src/waveglow/denoiser.py:58-68 with the dense-DFT
    STFT of src/common/stft.py:79-138): STFT GEMM -> spectral subtraction -> inverse STFT GEMM on 8 x 10 s of audio,
    tensor-core form (default) and exact FFMA form."""
    from fac_via_ppg_b200.waveglow.denoiser import Denoiser
    den = Denoiser(model, mode="zeros")
    n = synth.frames_for_seconds(seconds) * 160
    audio = torch.randn(batch, n, device=env.dev) * 0.1
    out = {"workload": "Denoiser(strength 0.005) on %d x %.0f s of audio (STFT 1024 / hop 160)" % (batch, seconds),
           "unit": "samples/s", "timing": "CUDA events, mean of 5 after 1 warm-up"}
    for precision in ("fp16x3", "fp32"):
        den.stft.precision = precision
        ms, clocks = env.timed_with_clocks(lambda: den(audio, strength=0.005), 5, warmup=1)
        out[precision] = {"value": batch * n / (ms / 5e3), "ms": ms / 5, "clocks": clocks}
    return out


def measure_cfg4(env, model, batch=32, seconds=10.0):
    """BASELINE configs[3]: WaveGlow.infer on 256 x 10 s sharded over 8 GPUs = 32 utterances per GPU (weak scaling:
    every rank runs its own 32), mel resident, and through pinned host buffers."""
    dev, F = env.dev, synth.frames_for_seconds(seconds)
    mel_host = synth.synthetic_mel(batch, F, seed=synth.SEED + 100 + env.rank).pin_memory()
    mel = mel_host.to(dev)
    out_host = torch.empty(batch, F * 160).pin_memory()

    def e2e():
        audio = model.infer(mel_host.to(dev, non_blocking=True), sigma=0.6)
        out_host.copy_(audio, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms, clocks = env.timed_with_clocks(lambda: model.infer(mel, sigma=0.6), 3, warmup=1)
    ms_e2e, _ = env.timed_with_clocks(e2e, 3, warmup=1)
    ms, ms_e2e = ms / 3, ms_e2e / 3
    n = env.world * batch * F * 160
    del mel
    torch.cuda.empty_cache()
    return {"workload": "WaveGlow.infer mel->wav, %d x %.0f s per GPU = %d utterances on %d GPU(s) (BASELINE configs[3])"
                        % (batch, seconds, batch * env.world, env.world),
            "precision": model.precision, "value": n / (ms / 1e3), "unit": "samples/s", "rtf": n / (ms / 1e3) / RATE,
            "ms_per_step": ms, "timing": "CUDA events, max over ranks, mean of 3 after 1 warm-up", "clocks": clocks,
            "tflops_algorithmic": n / (ms / 1e3) * WG_FLOP_PER_SAMPLE / 1e12,
            "e2e": {"value": n / (ms_e2e / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": mel_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4}}


def measure_cfg5(env, model, batch=8, seconds=60.0):
    """BASELINE configs[4]: long-form 60 s utterances, 64 across 8 GPUs = 8 per GPU, END TO END like the CLI
    (reference src/script/generate_synthesis.py:86-98): PPG (host) -> Tacotron2.inference -> WaveGlow.infer
    (sigma 0.6) -> Denoiser (strength 0.005) -> waveform (host).  Whole-job RTF = audio seconds of all ranks per
    second of the slowest rank."""
    from fac_via_ppg_b200.waveglow.denoiser import Denoiser
    dev, F = env.dev, synth.frames_for_seconds(seconds)
    taco = build_tacotron(dev, F)
    denoiser = Denoiser(model, mode="zeros")
    gen = torch.Generator(device=dev).manual_seed(synth.SEED + 200 + env.rank)
    logits = torch.randn((batch, F, 5816), generator=gen, device=dev) * 3.0       # synth.synthetic_ppg's recipe,
    ppg_host = torch.empty(batch, 5816, F).pin_memory()                             # drawn on the device (1.5 GB)
    ppg_host.copy_(torch.softmax(logits, dim=-1).transpose(1, 2))
    del logits
    # Through the public batch-stream helper (fac_via_ppg_b200/pipeline.py): the PPG upload (1.5 GB over PCIe) of step
    # k + 1 runs on a copy stream while step k is in the vocoder.  Every step still uploads its own input from pinned
    # host memory inside the timed region (the upload that overlaps the last timed step is that of a step not run).
    from fac_via_ppg_b200.pipeline import BatchStream
    stream = BatchStream(taco, model, denoiser, sigma=0.6, denoiser_strength=0.005)

    def endless():
        while True:
            yield ppg_host

    results = stream.run(endless(), to_host=True, record_phases=True)
    marks = []

    def step():
        next(results)                       # one batch: host PPG in, host waveform out (synchronised)
        marks.append(stream.phase_events)

    with quiet_stderr():
        ms, clocks = env.timed_with_clocks(step, 2, warmup=1)
    ms /= 2
    ev = marks[-1]
    n = env.world * batch * F * 160
    results.close()
    torch.cuda.synchronize()
    del taco, denoiser, stream, results
    torch.cuda.empty_cache()
    return {"workload": "PPG->Mel->WaveGlow->Denoiser end to end, %d x %.0f s per GPU = %d utterances on %d GPU(s), "
                        "host PPG in, host waveform out (BASELINE configs[4]; generate_synthesis.py:86-98)"
                        % (batch, seconds, batch * env.world, env.world),
            "precision": "acoustic model fp16x3 (fp32-grade), vocoder %s" % model.precision,
            "value": n / (ms / 1e3), "unit": "samples/s", "rtf_whole_job": n / (ms / 1e3) / RATE, "ms_per_step": ms,
            "timing": "CUDA events, max over ranks, mean of 2 after 1 warm-up; the PPG upload of step k+1 runs on a copy "
                      "stream behind the vocoder of step k", "clocks": clocks,
            "phases_ms_rank0": {"ppg2mel": ev[0].elapsed_time(ev[1]), "mel2wav": ev[1].elapsed_time(ev[2]),
                                "denoiser+d2h": ev[2].elapsed_time(ev[3])},
            "frames": F, "decoder_steps": F,
            "h2d_bytes_per_step": ppg_host.numel() * 4, "d2h_bytes_per_step": batch * F * 160 * 4}


REF_BATCH = 2      # utterances per reference step (see run_reference)


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
    """Reference arm: the reference's own CPU implementation of WaveGlow.infer (oracle port: the reference is pure
    Python/PyTorch, there is nothing to compile into oracle/_ref) on all host threads.  Each step is a BOUNDED
    sample of the workload: REF_BATCH = 2 full-length 10 s utterances instead of 8 -- at ~4.5 s per utterance on 16
    cores the full batch would make `--steps 20` a 12-minute run; per-utterance shapes are those of the config and
    CPU throughput per sample does not depend on the batch size beyond 2."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    frames = synth.frames_for_seconds(args.seconds)
    cfg = synth.WAVEGLOW_CONFIG
    from oracle import waveglow_oracle
    torch.set_num_threads(threads)
    sd = synth.waveglow_state(cfg=cfg)
    mel = synth.synthetic_mel(REF_BATCH, frames)
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 1))):
            waveglow_oracle.waveglow_infer(sd, cfg, mel, 0.6)
        t0 = time.perf_counter()
        n = 0
        for _ in range(args.steps):
            n += waveglow_oracle.waveglow_infer(sd, cfg, mel, 0.6).numel()
        dt = time.perf_counter() - t0
    value = n / dt
    sample = ("%d x %d frames (%d x %.2f s of audio) per step instead of %d x: bounded sample, oracle port of glow.py "
              "on CPU fp32, 1 warm-up step" % (REF_BATCH, frames, REF_BATCH, frames * 160 / RATE, args.batch))
    line = {
        "impl": "reference", "metric": "audio samples/sec @22.05 kHz (WaveGlow.infer mel->wav)", "value": value,
        "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "rtf": value / RATE,
        "config": dict(workload_config(args, args.batch, synth.frames_for_seconds(args.seconds)), precision="fp32 (torch CPU)",
                       reference_batch_per_step=REF_BATCH),
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
    ap.add_argument("--no-side", action="store_true", help="skip every side object (headline line only)")
    args = ap.parse_args()

    env = Env()
    rank, world = env.rank, env.world

    if args.impl == "reference":
        run_reference(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: this package has no CPU path")
    env.init_cuda()
    dev, distributed = env.dev, env.distributed
    if distributed:
        import torch.distributed as dist

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

    def step_resident():
        model.infer(mel, sigma=0.6)

    def step_e2e():
        m = mel_host.to(dev, non_blocking=True)
        audio = model.infer(m, sigma=0.6)
        out_host.copy_(audio, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()
    lib.fac_reset_launch_count()
    torch.cuda.nvtx.range_push("fac_timed")      # ncu --nvtx --nvtx-include "fac_timed/" isolates this region
    ms = env.timed(step_resident, args.steps)
    torch.cuda.nvtx.range_pop()
    launches = lib.fac_launch_count()
    clocks = sampler.stop() if rank == 0 else None

    step_e2e()
    ms_e2e = env.timed(step_e2e, args.steps)

    value = world * samples_per_step * args.steps / (ms / 1e3)
    e2e_value = world * samples_per_step * args.steps / (ms_e2e / 1e3)

    # ---- the two multi-GPU configs BASELINE.json names (collective timing: every rank takes part) ----
    cfg4 = cfg5 = None
    if not args.no_side:
        cfg4 = measure_cfg4(env, model)
        cfg5 = measure_cfg5(env, model)

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the WN layer GEMMs), timed live with CUDA events
    peaks = measured_peaks()
    roof = model.profile_dominant_kernel(mel, peaks, passes=max(3, min(args.steps, 10)))
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if roof is not None and os.path.isfile(traffic_path):
        with open(traffic_path) as fh:
            t = json.load(fh).get(precision)
        if t:
            roof["traffic"] = t["bytes_per_launch"]                         # DRAM bytes per launch (ncu --set full)
            roof["traffic_source"] = t["source"]

    side = world == 1 and not args.no_side
    fp32 = library = ppg2mel = pipeline = cpu = denoiser = None
    if side and precision != "fp32":
        fp32 = measure_fp32_mode(env, model, mel, peaks, samples_per_step)
    if side:
        library = measure_library_gpu(env, mel, samples_per_step)
        denoiser = measure_denoiser(env, model)
    # ---- PPG -> Mel side metric (mel frames/s) and BASELINE configs[2] ----------
    if side and not args.no_ppg2mel:
        ppg2mel = measure_ppg2mel(env)
        pipeline = measure_pipeline(env, model)

    # ---- CPU baseline (bounded sample of the same workload, rank 0 only) ----------------
    if side and not args.no_cpu_baseline:
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
        "fp32": fp32,
        "library_gpu": library,
        "ppg2mel": ppg2mel,
        "pipeline": pipeline,
        "denoiser": denoiser,
        "cfg4": cfg4,
        "cfg5": cfg5,
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

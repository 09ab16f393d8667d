"""Batch streaming of the synthesis path of reference ``src/script/generate_synthesis.py:86-98``
(PPG -> Tacotron2.inference -> WaveGlow.infer -> Denoiser) over a sequence of HOST batches.

An extension, not a reference symbol: the reference CLI handles one utterance per process.  A service that feeds
batch after batch leaves the PCIe link idle while the vocoder runs (0.5 s for 8 x 60 s) and the GPU idle while the
next posteriorgrams (23 KB per frame: 1.5 GB for 8 x 60 s) come up.  Here the upload of batch k + 1 runs on a copy
stream behind the vocoder of batch k, into the second of two device buffers; the results of a batch are those of
the three module calls made one after the other (same kernels, same stream order).
"""
from __future__ import annotations

import torch


class BatchStream:
    """``for wav in BatchStream(tacotron, waveglow, denoiser).run(host_batches): ...``

    host_batches: iterable of (B, n_symbols, T) float32 CPU tensors (pinned memory makes the upload asynchronous);
    all batches of one run() must have the same shape.  Yields (B, T_out * hop) waveforms on the GPU (``to_host``:
    on the CPU, in a pinned buffer that is reused by the next batch -- copy it if you keep it).
    """

    def __init__(self, tacotron, waveglow, denoiser=None, sigma=0.6, denoiser_strength=0.005, mel_clip=(-11.5, 2.0)):
        self.tacotron, self.waveglow, self.denoiser = tacotron, waveglow, denoiser
        self.sigma, self.strength, self.mel_clip = sigma, denoiser_strength, mel_clip
        self.phase_events = None          # (start, mel done, vocoder done, end) CUDA events of the last batch

    def run(self, host_batches, to_host=False, record_phases=False):
        it = iter(host_batches)
        nxt = next(it, None)
        if nxt is None:
            return
        dev = next(self.waveglow.parameters()).device
        main = torch.cuda.current_stream(dev)
        copy_stream = torch.cuda.Stream(device=dev)
        bufs = [torch.empty(nxt.shape, device=dev, dtype=torch.float32) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        out_host = None

        def upload(batch, b):
            if batch.shape != bufs[b].shape:
                raise ValueError("BatchStream: every batch of a run must have the shape of the first one")
            copy_stream.wait_stream(main)                 # the last reader of buffer b has been queued on `main`
            with torch.cuda.stream(copy_stream):
                bufs[b].copy_(batch, non_blocking=True)
                ready[b].record(copy_stream)

        upload(nxt, 0)
        k = 0
        while nxt is not None:
            b = k & 1
            k += 1
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record_phases else None
            if ev:
                ev[0].record()
            main.wait_event(ready[b])
            mel = self.tacotron.inference(bufs[b])[1]
            if ev:
                ev[1].record()
            nxt = next(it, None)
            if nxt is not None:
                upload(nxt, b ^ 1)                        # behind this batch's vocoder
            if self.mel_clip is not None:
                mel = mel.clamp(*self.mel_clip)
            wav = self.waveglow.infer(mel.contiguous(), sigma=self.sigma)
            if ev:
                ev[2].record()
            if self.denoiser is not None:
                wav = self.denoiser(wav, strength=self.strength)[:, 0]
            if to_host:
                if out_host is None or out_host.shape != wav.shape:
                    out_host = torch.empty(wav.shape, dtype=wav.dtype).pin_memory()
                out_host.copy_(wav, non_blocking=True)
            if ev:
                ev[3].record()
                self.phase_events = ev
            if to_host:
                main.synchronize()
                yield out_host
            else:
                yield wav

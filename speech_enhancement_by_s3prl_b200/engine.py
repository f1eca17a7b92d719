"""Fused evaluation / training step of the enhancement path.

``EnhancementEngine.eval_step`` computes what the reference's evaluation step computes
(runner.py:556-602): features of the noisy channel -> mask head -> enhanced waveform
(``_decode_wav``: iSTFT with the noisy phase, zero-pad, level-match to the clean reference)
-> spectral SISDR criterion -> per-utterance waveform SI-SDR -- in four kernels:

  K1  se_stft            noisy wav -> (log-)power features               read 4H  write 4K  /frame
  K2  se_cmvn_stats + se_linear_head_fwd   features -> mask             read 4K  write 4K
  K3  se_mask_istft      noisy+clean wav, mask -> enhanced wav + sums    read 8H+4K write 4H
  K3' se_finalize_metrics  gain, SI-SDR, loss terms; wav *= gain (in place)

Nothing else is materialised: no noise-channel STFT, no phase, no `predicted`, no int64 masks.
"""
import ctypes

import torch

from . import _lib, dp, ops


class EnhancementEngine:
    def __init__(self, preprocessor, head, log_features=True, precision=0):
        """preprocessor: se_b200 OnlinePreprocessor (gives n_fft / hop / window, channel_inp/tar);
        head: se_b200 LinearResidual on the (log-)power spectrum of the input channel."""
        self.pre = preprocessor
        self.head = head
        self.log_features = bool(log_features)
        self.precision = precision
        self.n_fft = preprocessor._win_args["n_fft"]
        self.hop = preprocessor._win_args["hop_length"]
        self.ch_inp = int(getattr(preprocessor, "channel_inp", 0))
        self.ch_tar = int(getattr(preprocessor, "channel_tar", 1))
        self._graphs = {}

    # ------------------------------------------------------------------ device-resident step
    def eval_step(self, lengths, wavs, want_spec_loss=True):
        """lengths (B,) int64, wavs (B, C, T) fp32, both on the GPU.
        Returns dict(loss_per_utt (B,), sisdr (B,), wav_predicted (B, T), gain (B,))."""
        B, C, T = wavs.shape
        dev = wavs.device
        window = self.pre._frame_window
        if window.device != dev:
            self.pre.to(dev)
            window = self.pre._frame_window
        head = self.head
        with torch.no_grad():
            spec = ops.stft(wavs, self.ch_inp, self.n_fft, self.hop, window, power=not self.log_features,
                            logpower=self.log_features, log_eps=self.pre.eps)
            feats = spec["logpower"] if self.log_features else spec["power"]
            mean = std = None
            if head.cmvn:
                mean, std = ops.cmvn_stats(feats)
            mask, _ = ops.linear_head_fused(feats, head.linear.weight, head.linear.bias, head.activation, mean, std,
                                            head.eps, precision=self.precision)
            wav, sums = ops.mask_istft(wavs, self.ch_inp, self.ch_tar, mask, lengths, self.n_fft, self.hop, window,
                                       pad_to=T, want_sums=True, want_spec=want_spec_loss)
            gain, sisdr, loss = ops.finalize_metrics(sums, lengths, T, wav=wav, target_db=None)
        return {"loss_per_utt": loss, "sisdr": sisdr, "wav_predicted": wav, "gain": gain, "mask": mask}

    # ------------------------------------------------------------------ CUDA-graph replay of the step
    def capture(self, B, C, T, device):
        """Capture eval_step for a fixed (B, C, T) into a CUDA graph; returns the static buffers."""
        key = (B, C, T, str(device))
        if key in self._graphs:
            return self._graphs[key]
        ops.prepare(self.n_fft)
        static = {"lengths": torch.full((B,), T, dtype=torch.int64, device=device),
                  "wavs": torch.zeros(B, C, T, device=device)}
        static["wavs"].normal_(0, 0.05)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(2):                                   # warm up allocator + lazy init outside capture
                self.eval_step(static["lengths"], static["wavs"])
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.eval_step(static["lengths"], static["wavs"])
        static.update(out)
        static["graph"] = graph
        self._graphs[key] = static
        return static

    def eval_step_graph(self, lengths, wavs):
        """Same result as eval_step through the captured graph (inputs are copied into the static buffers)."""
        B, C, T = wavs.shape
        st = self.capture(B, C, T, wavs.device)
        st["lengths"].copy_(lengths, non_blocking=True)
        st["wavs"].copy_(wavs, non_blocking=True)
        st["graph"].replay()
        return st

    # ------------------------------------------------------------------ host-facing step (e2e)
    def eval_step_host(self, lengths_cpu, wavs_cpu, use_graph=True):
        """The call a user makes with collate_fn's output (CPU tensors; pinned for speed):
        H2D of the batch, the fused step, D2H of the per-utterance loss terms and SI-SDR.
        Returns (mean_loss, mean_sisdr, wav_predicted on device)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        if use_graph:
            B, C, T = wavs_cpu.shape
            st = self.capture(B, C, T, dev)
            st["lengths"].copy_(lengths_cpu, non_blocking=True)
            st["wavs"].copy_(wavs_cpu, non_blocking=True)
            st["graph"].replay()
            out = st
        else:
            out = self.eval_step(lengths_cpu.to(dev, non_blocking=True), wavs_cpu.to(dev, non_blocking=True))
        res = torch.stack([out["loss_per_utt"], out["sisdr"]]).cpu()          # D2H + sync
        return res[0].mean().item(), res[1].mean().item(), out["wav_predicted"]

    # ------------------------------------------------------------------ training step (head fwd + bwd)
    def train_step(self, lengths, wavs, objective, optimizer=None, grad_clip=None):
        """runner.py:431-471 on the kernels: preprocessor tensors -> head -> criterion -> backward
        (+ gradient all-reduce under DP) -> optimizer step.  Returns the loss tensor."""
        c = self.pre.get_feat_config
        feat_cfg = c("linear", self.ch_inp, log=self.log_features)
        feats, linear_inp, linear_tar = self.pre(wavs, [feat_cfg, c("linear", self.ch_inp), c("linear", self.ch_tar)])
        predicted, extra = self.head(features=feats, linears=linear_inp)
        frames = lengths // self.hop + 1
        loss, _ = objective(predicted=predicted, linear_tar=linear_tar, linear_inp=linear_inp, stft_lengths=frames, **extra)
        if optimizer is not None:
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            dp.allreduce_gradients(self.head.parameters())
            if grad_clip is not None:
                torch.nn.utils.clip_grad_norm_(list(self.head.parameters()), grad_clip)
            optimizer.step()
        return loss

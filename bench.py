#!/usr/bin/env python
"""bench.py -- enhanced audio-seconds per second through the fused STFT -> mask -> iSTFT + loss path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): VoiceBank-DEMAND-shaped synthetic batch, 16 kHz, 64 x 4 s
utterances per GPU, n_fft=512 hop=256, LinearResidual mask head on the log-power spectrum, the
reference's evaluation step (runner.py:556-602): enhanced waveform level-matched to the clean
reference, spectral SISDR criterion, per-utterance waveform SI-SDR.  One step = one batch.

Own arm: `value` = device-timed graph replays with the batches already in HBM (inputs rotate over
a ring of distinct batches larger than L2); `e2e` = the same through EnhancementEngine's host
pipeline with pinned HOST batches, H2D and D2H inside the timed region.  `roofline` is for the
dominant kernel (fused mask -> iSTFT), timed with CUDA events around its launches.  `cpu_baseline`
(rank 0, N=1) times the CPU oracle -- the reference's un-fused sequence on torch CPU ops -- on the
host cores.  `--impl reference` runs that oracle alone as the reference arm (the reference is pure
Python whose STFT lives in an un-vendored dependency).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "enhanced audio-sec/sec (STFT->mask->iSTFT+loss)"
UNIT = "audio-s/s"
SR = 16000
N_UTT, SECONDS, N_FFT, HOP = 64, 4.0, 512, 256
PRE_KW = dict(sample_rate=SR, win_ms=32, hop_ms=16, n_freq=257, n_mels=40, n_mfcc=13)
WORKLOAD = "configs[1]: VoiceBank-DEMAND-shaped synthetic, 16 kHz, 64x4s per GPU, n_fft=512 hop=256, LinearResidual(257) on log-power, eval step + SISDR + waveform SI-SDR"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled by NVML in a background thread while the GPU is busy."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.index = index
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": "nvmlClocksEventReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksEventReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksEventReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksEventReasonSwPowerCap"}
        bits = {k: getattr(nv, v, None) for k, v in names.items()}
        legacy = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
                  "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}
        for k, v in legacy.items():
            if bits[k] is None:
                bits[k] = getattr(nv, v, None)
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, b in bits.items():
                    if b is not None and mask & b:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self.thread.start()

    def stop(self):
        self._stop.set()
        if self.nv is not None and self.thread.is_alive():
            self.thread.join(timeout=1.0)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def oracle_objects():
    from oracle import signal_path as sp
    from oracle.preprocessor import OnlinePreprocessor as OraclePre
    ora = OraclePre(**PRE_KW)
    ora.channel_inp, ora.channel_tar = 0, 1
    c = ora.get_feat_config
    # the six tensors of run_downstream.py:150-157 with the log-power spectrum as up/downstream feature
    ora.feat_list = [c("linear", 0, log=True), c("linear", 0, log=True), c("linear", 0), c("phase", 0), c("linear", 1), c("phase", 1)]
    return sp, ora


def time_cpu_oracle(lengths, wavs, weight, bias, budget_s, min_reps=2, max_reps=50):
    """Best-of wall-clock of the un-fused reference sequence on the host cores (torch CPU, fp32, no_grad)."""
    sp, ora = oracle_objects()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    head = dict(weight=weight, bias=bias)
    times = []
    t_start = time.perf_counter()
    with torch.no_grad():
        sp.eval_step(ora, head, lengths, wavs)                     # warm-up (thread pools, FFT plans)
        while len(times) < max_reps and (len(times) < min_reps or time.perf_counter() - t_start < budget_s):
            t0 = time.perf_counter()
            sp.eval_step(ora, head, lengths, wavs)
            times.append(time.perf_counter() - t0)
    return min(times), cores, len(times)


def run_reference_arm(args, rank):
    """--impl reference: the CPU oracle port of the reference path on all host threads."""
    if rank != 0:
        return
    from speech_enhancement_by_s3prl_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sp, ora = oracle_objects()
    torch.manual_seed(1337)
    lin = torch.nn.Linear(257, 257)
    head = dict(weight=lin.weight.detach(), bias=lin.bias.detach())
    n_utt = N_UTT
    lengths, wavs = synth.batch(n_utt, SECONDS)
    with torch.no_grad():
        t0 = time.perf_counter()
        sp.eval_step(ora, head, lengths, wavs)
        first = time.perf_counter() - t0
        while first * (args.steps + args.warmup) > 150.0 and n_utt > 4:   # keep the whole run within a few minutes
            n_utt //= 2
            first /= 2
        lengths, wavs = lengths[:n_utt], wavs[:n_utt].contiguous()
        for _ in range(args.warmup):
            sp.eval_step(ora, head, lengths, wavs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sp.eval_step(ora, head, lengths, wavs)
        total = time.perf_counter() - t0
    audio_s = n_utt * SECONDS
    value = audio_s * args.steps / total
    sample = f"{n_utt}x{SECONDS:g}s per step ({'full batch' if n_utt == N_UTT else 'sub-sampled batch'}), {args.steps} steps"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "where": "host CPU, torch ops (torch.stft/istft) restating the reference's un-fused sequence"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ring", type=int, default=8, help="distinct batches the device-resident run rotates over")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU oracle timing for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch kernels eagerly instead of replaying CUDA graphs")
    ap.add_argument("--streams", type=int, default=2, help="independent steps in flight (CUDA streams) in the device-resident run")
    ap.add_argument("--head-precision", type=int, default=1, help="0 = fp32 SIMT head, 1 = TF32 tcgen05 head")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch.distributed as dist
    import speech_enhancement_by_s3prl_b200 as se
    from speech_enhancement_by_s3prl_b200 import dp, ops, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---------------------------------------------------------------- model + data
    pre = se.OnlinePreprocessor(**PRE_KW).to(dev)
    pre.channel_inp, pre.channel_tar = 0, 1
    torch.manual_seed(1337)
    head = se.LinearResidual(input_size=257, output_size=257).to(dev)
    engine = se.EnhancementEngine(pre, head, log_features=True, precision=args.head_precision)
    T = int(SECONDS * SR)
    ring_host = [synth.batch(N_UTT, SECONDS, first_index=(rank * args.ring + r) * N_UTT) for r in range(min(args.ring, 4))]
    # device ring: the first few slots come from the host batches, the rest are level-preserving
    # circular shifts of them (distinct bytes; generation on the host is the slow part)
    ring = []
    for r in range(args.ring):
        lengths, wavs = ring_host[r % len(ring_host)]
        d = wavs.to(dev)
        if r >= len(ring_host):
            d = torch.roll(d, shifts=1777 * r, dims=-1).contiguous()
        ring.append((lengths.to(dev), d))
    audio_s_per_step = N_UTT * SECONDS
    F = T // HOP + 1
    K = N_FFT // 2 + 1

    # ---------------------------------------------------------------- device-resident timed run
    if args.eager:
        steps = [lambda l=l, w=w: engine.eval_step(l, w) for l, w in ring]
    else:
        graphs = [engine.capture_bound(l, w) for l, w in ring]
        steps = [g["graph"].replay for g in graphs]
    sampler = ClockSampler(local_rank)
    # Consecutive steps work on independent batches: with --streams S > 1, step i is enqueued on stream i % S, so the
    # ramp-up and tail of one step's kernels overlap the neighbouring step's kernels (every ring slot owns its buffers).
    main = torch.cuda.current_stream()
    lanes = [torch.cuda.Stream(device=dev) for _ in range(args.streams)] if args.streams > 1 else [main]

    def run_steps(n):
        if len(lanes) > 1:
            fork = torch.cuda.Event()
            fork.record(main)
            for st in lanes:
                st.wait_event(fork)
        for i in range(n):
            with torch.cuda.stream(lanes[i % len(lanes)]):
                steps[i % len(steps)]()
        if len(lanes) > 1:
            for st in lanes:
                main.wait_stream(st)

    run_steps(args.warmup)
    if world > 1:                                      # communicator setup and the first collective stay out of the timed region
        warm = graphs[0] if not args.eager else engine.eval_step(*ring[0])
        for _ in range(3):
            dp.global_means(warm["loss_per_utt"], warm["sisdr"])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(args.steps)
    if world > 1:                                      # the one collective of an evaluation pass: metric sums
        out = graphs[0] if not args.eager else engine.eval_step(*ring[0])
        dp.global_means(out["loss_per_utt"], out["sisdr"])
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    value = world * audio_s_per_step * args.steps / (ms_total * 1e-3)

    # ---------------------------------------------------------------- per-kernel timing, roofline of the dominant kernel
    # Each kernel is captured alone in a CUDA graph that launches it once per ring slot (distinct, HBM-cold
    # inputs; no CPU launch gaps between the launches) and timed with CUDA events around the replays.
    window = pre._frame_window
    wpad = engine._padded_weight()
    LD = ops.round4(K)
    fused = engine.precision == 1 and ops.linear_head_tma_supported(N_UTT, F, K, K, LD, wpad.shape[1], LD)
    use_ws = fused and engine.use_spec_ws and ops.spec_ws_supported(N_FFT, HOP)
    slots = []
    with torch.no_grad():
        for lengths, wavs in ring:
            sl = dict(lengths=lengths, wavs=wavs)
            if fused:
                sl["stat_sums"] = torch.zeros(N_UTT, LD, 2, device=dev, dtype=torch.float64)
                sl["spec_ws"] = torch.empty(N_UTT, F, ops.SPEC_WS_FLOATS, device=dev) if use_ws else None
                sl["feats"], _ = ops.stft_features(wavs, 0, N_FFT, HOP, window, logpower=True, stat_sums=sl["stat_sums"],
                                                   spec_ws=sl["spec_ws"])
                sl["mask"] = ops.linear_head_tma(sl["feats"], K, wpad, head.linear.bias, head.activation, sl["stat_sums"], head.eps)
            else:
                sl["feats"] = ops.stft_padded(wavs, 0, N_FFT, HOP, window, logpower=True)
                sl["mean"], sl["std"] = ops.cmvn_stats_padded(sl["feats"], K)
                sl["mask"] = ops.linear_head_padded(sl["feats"], K, wpad, head.linear.bias, head.activation, sl["mean"], sl["std"], head.eps,
                                                    precision=engine.precision)
            sl["wav"], sl["sums"] = ops.mask_istft(wavs, 0, 1, sl["mask"], lengths, N_FFT, HOP, window, pad_to=T, mask_padded=True,
                                                   spec_ws=sl.get("spec_ws"))
            slots.append(sl)
    torch.cuda.synchronize()
    if fused:
        # (the sums buffer keeps accumulating across timing launches: the values are not used here)
        launchers = {
            "stft": lambda s: ops.stft_features(s["wavs"], 0, N_FFT, HOP, window, logpower=True, stat_sums=s["stat_sums"],
                                                spec_ws=s["spec_ws"]),
            "head": lambda s: ops.linear_head_tma(s["feats"], K, wpad, head.linear.bias, head.activation, s["stat_sums"], head.eps),
        }
    else:
        launchers = {
            "stft": lambda s: ops.stft_padded(s["wavs"], 0, N_FFT, HOP, window, logpower=True),
            "cmvn_stats": lambda s: ops.cmvn_stats_padded(s["feats"], K),
            "head": lambda s: ops.linear_head_padded(s["feats"], K, wpad, head.linear.bias, head.activation, s["mean"], s["std"], head.eps,
                                                     precision=engine.precision),
        }
    launchers["mask_istft"] = lambda s: ops.mask_istft(s["wavs"], 0, 1, s["mask"], s["lengths"], N_FFT, HOP, window, pad_to=T,
                                                       mask_padded=True, out=s["wav"], sums=s["sums"], spec_ws=s.get("spec_ws"))
    launchers["finalize"] = lambda s: ops.finalize_metrics(s["sums"], s["lengths"], T, wav=s["wav"])
    kernel_ms = {}
    reps = max(3, min(20, args.steps // len(slots)))
    for name, fn in launchers.items():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for sl in slots[:2]:
                fn(sl)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g), torch.no_grad():
            for sl in slots:
                fn(sl)
        g.replay()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            g.replay()
        a1.record()
        torch.cuda.synchronize()
        kernel_ms[name] = a0.elapsed_time(a1) / (reps * len(slots))
    # algorithmic bytes per launch (DESIGN.md "kernels"): K3 reads noisy + clean (2 x 4T), the mask (4FK), writes 4T per utterance
    alg_bytes = {"stft": N_UTT * (4 * T + 4 * F * K), "cmvn_stats": N_UTT * 4 * F * K, "head": N_UTT * 8 * F * K,
                 "mask_istft": N_UTT * (12 * T + 4 * F * K), "finalize": N_UTT * 8 * T}
    dominant = max(kernel_ms, key=kernel_ms.get)
    peak, peak_src = peaks()
    achieved = alg_bytes[dominant] / (kernel_ms[dominant] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(dominant, {}).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes[dominant],
                "kernel_ms": {k: round(v, 5) for k, v in kernel_ms.items()},
                "note": "K1/K3 are bound by the fp32 pipe + shared-memory wavefronts, not HBM (DESIGN.md 4.1: 22 M fp32-pipe cycles "
                        "and 19 M issue slots per K3 launch = ~48 % / ~50 % busy; DRAM 10-24 %)",
                "step_algorithmic_bytes": int(audio_s_per_step * 16 * SR * (1 + K / HOP)),
                "step_frac_of_hbm_peak": (audio_s_per_step * 16 * SR * (1 + K / HOP)) / (ms_total / args.steps * 1e-3) / 1e9 / peak}

    # ---------------------------------------------------------------- end-to-end through the host pipeline
    pipe = engine.host_pipeline(N_UTT, 3, T, depth=2, device=dev)
    pinned = [(l.clone().pin_memory(), w.clone().pin_memory()) for l, w in ring_host]
    for i in range(max(3, min(args.warmup, 10))):
        pipe.submit(*pinned[i % len(pinned)])
    pipe.drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        pipe.submit(*pinned[i % len(pinned)])
    results = pipe.drain()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * audio_s_per_step * args.steps / t.item()
    clocks = sampler.stop()
    mean_sisdr = float(torch.stack([r[1] for r in results]).mean())
    mean_loss = float(torch.stack([r[0] for r in results]).mean())

    # ---------------------------------------------------------------- CPU baseline on the host cores (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        lengths, wavs = ring_host[0]
        best, cores, reps = time_cpu_oracle(lengths, wavs, head.linear.weight.detach().cpu(), head.linear.bias.detach().cpu(), args.cpu_budget)
        cpu = {"value": audio_s_per_step / best, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"full batch {N_UTT}x{SECONDS:g}s, best of {reps} runs of the un-fused oracle sequence (torch CPU fp32)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "head": "tf32 tcgen05 (fp32 accumulate)" if args.head_precision == 1 else "fp32 SIMT", "utterances_per_gpu": N_UTT, "seconds": SECONDS, "n_fft": N_FFT, "hop": HOP,
                           "launch": ("eager" if args.eager else "cuda-graph replay") + f", {args.streams} step(s) in flight (streams)",
                           "l2": f"inputs rotate over {args.ring} distinct device batches ({args.ring * N_UTT * 3 * T * 4 / 1e6:.0f} MB > 126 MB L2); no explicit flush",
                           "parallelism": f"dp{world} (utterance-sharded, no data-path collective)"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                        "ms_per_step": 1e3 * t.item() / args.steps, "pipeline_depth": 2},
                "gpu_launches": engine.launches_per_step * args.steps, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
                "check": {"mean_sisdr_db": mean_sisdr, "mean_loss": mean_loss}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

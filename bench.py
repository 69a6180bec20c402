#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: audio-seconds processed per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3]): standard mode, --gate_ui 50, a batch of 5-minute 44.1 kHz stereo
synthetic tracks (pink noise + on/off envelope + tone bursts), sharded by whole track across ranks:
128 tracks per GPU, so N = 8 is the full 1024-track batch (1024 tracks of fp32 in + out do not fit one
GPU's HBM; the per-GPU shard does).  A "step" is one pass of the whole hot path over the rank's tracks:
levels -> gate scan -> fp64 edge frames -> fused STFT/OLA with in-kernel per-chunk limiter.  Weak scaling, no data-path
collective; the only cross-rank traffic is the timing reduction.

One JSON line on stdout (rank 0).  `value` is device-resident throughput, `e2e` the same metric through
the host-buffer API (pinned host -> H2D -> kernels -> D2H -> pinned host, copies inside the timed region),
`roofline` the fused STFT kernel against the measured HBM copy bandwidth using ALGORITHMIC bytes
(16 B per stereo sample-frame: read once + write once), `cpu_baseline` the NumPy port of the reference
(oracle/) timed on the host cores.  `--impl reference` times that CPU port on all host cores instead.
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

SR = 44100
TRACK_SECONDS = 300.0
N_SAMPLES = int(SR * TRACK_SECONDS)          # 13 230 000 sample-frames
ALG_BYTES_PER_SF = 16                        # fp32 stereo read once + written once
METRIC = "audio_seconds_per_second"
UNIT = "audio-s/s"


def workload_config(tracks_per_gpu, n_gpus):
    return {
        "workload": f"BASELINE configs[3]: batch of {tracks_per_gpu * n_gpus} synthetic 5 min 44.1 kHz stereo tracks "
                    f"({tracks_per_gpu}/GPU, sharded by track), process_tomatis standard mode --gate_ui 50",
        "mode": "standard", "gate_ui": 50, "sample_rate": SR, "track_seconds": TRACK_SECONDS,
        "tracks_per_gpu": tracks_per_gpu, "n_fft": 4096, "hop": 2048,
        "l2_policy": "inputs larger than L2 (13.5 GB in + 13.5 GB out per GPU per step, streamed once)",
        "parallelism": f"track-sharded x{n_gpus}, no collective on the data path",
    }


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ts, line in self.rows:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in timed region"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
_CPU_TRACKS = {}


def _cpu_track(seed):
    from tomatis_audio_processor_b200 import synth
    if seed not in _CPU_TRACKS:
        _CPU_TRACKS[seed] = synth.recipe_gated_pink(TRACK_SECONDS, SR, seed)
    return _CPU_TRACKS[seed]


def _cpu_worker(seed):
    from oracle import tomatis_oracle as orc           # the reference's algorithm, restated in NumPy
    x = _cpu_track(seed)
    t = time.perf_counter()
    res = orc.process_standard(x, SR, gate_ui=50)
    return time.perf_counter() - t, int(res["out"].shape[0])


def cpu_baseline_single(n_tracks=3):
    """Oracle port on ONE host core, bounded sample (n_tracks x 5 min)."""
    total_t, total_sf = 0.0, 0
    for i in range(n_tracks):
        dt, n = _cpu_worker(1000 + (i % 2))
        total_t += dt
        total_sf += n
    return {"value": total_sf / SR / total_t, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n_tracks} tracks x {TRACK_SECONDS:.0f} s @ {SR} Hz through oracle.process_standard (NumPy "
                      f"restatement of src/process_tomatis.py, float32 pocketfft), one process"}


def parity_vs_oracle(x0, y0, states0, chunk_peaks0):
    """GPU track 0 of the timed batch against the oracle on the same samples (checker only): PCM max-abs error against the
    float64-FFT evaluation of the reference source (the north star's yardstick) and against its float32-FFT output, gate
    mismatches, limiter chunks that differ in their over-the-limit decision."""
    import numpy as np
    from oracle import tomatis_oracle as orc
    o32 = orc.process_standard(x0, SR, gate_ui=50)
    o64 = orc.run("standard", x0, SR, gate_ui=50, fft_dtype="float64")
    y = y0.astype(np.float64)
    d64 = np.abs(y - o64["out"].astype(np.float64)).max(axis=1)
    d32 = np.abs(y - o32["out"].astype(np.float64)).max(axis=1)
    self_noise = np.abs(o32["out"].astype(np.float64) - o64["out"].astype(np.float64)).max(axis=1)
    lens = np.asarray(o32["chunk_lengths"])
    ends = np.cumsum(lens)
    o_over = np.array([float(np.abs(o32["out"][e - n:e]).max()) >= 0.999 - 1e-6 for n, e in zip(lens, ends)])
    return {"parity_max_abs": float(d64.max()), "parity_max_abs_vs_f32_fft": float(d32.max()),
            "parity_pointwise_ok": bool(np.all(d32 <= 1e-5 + self_noise)),
            "oracle_self_noise": float(self_noise.max()),
            "gate_mismatches": int((np.asarray(states0) != np.asarray(o32["states"])).sum()),
            "gate_frames": int(len(o32["states"])),
            "limited_chunks_gpu": int((np.asarray(chunk_peaks0) > np.float32(0.999)).sum()), "limited_chunks_oracle": int(o_over.sum()),
            "parity_sample": "track 0 of the timed batch (13 230 000 sample-frames) against oracle.process_standard on the same samples; "
                             "parity_max_abs is against the float64-FFT evaluation of the reference source, bar 1e-5"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import numpy as np
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workers = max(1, min(cores, args.cpu_workers or cores))
    seeds = [1000, 1001]
    for s in seeds:                       # synthesise before forking so workers share the pages
        _cpu_track(s)
    ctx = mp.get_context("fork")
    jobs = [seeds[i % len(seeds)] for i in range(workers)]
    times = []
    with ctx.Pool(workers) as pool:
        for it in range(args.warmup + args.steps):
            t = time.perf_counter()
            res = pool.map(_cpu_worker, jobs)
            dt = time.perf_counter() - t
            if it >= args.warmup:
                times.append(dt)
    sf = sum(r[1] for r in res)
    total = sum(times)
    value = sf / SR * len(times) / total
    sample = (f"each step: {workers} x one 5 min 44.1 kHz track (one process per host core, NumPy {np.__version__}) "
              f"through oracle.process_standard; bounded sample of the {args.tracks_per_gpu * args.gpus}-track workload")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.tracks_per_gpu, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per stft_kernel launch from the committed ncu capture, scaled per sample-frame."""
    p = os.path.join(ROOT, "profiles", "stft_kernel_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return None


def fp32_secondary(sf_per_s, sm_mhz):
    """Secondary ceiling: the kernel's FP32 work against the non-tensor FP32 peak.  Instruction counts are the
    static SASS counts of the frame loop (profiles/stft_kernel_flops.json, tools/sass_count.py); the pipe executes 128 lanes per
    clock and SM whether an instruction is an add, a multiply or a fused multiply-add, so `frac` is lane-ops against lane slots
    (what bounds the kernel) and `achieved` the real flop rate (an FMA counts 2) against the 2-flop-per-slot peak."""
    try:
        with open(os.path.join(ROOT, "profiles", "stft_kernel_flops.json")) as f:
            c = json.load(f)
    except Exception:
        return None
    mhz = float(sm_mhz or 1965.0)
    slots = 128 * 148 * mhz * 1e6                        # FP32 lane slots per second
    return {"bound": "fp32", "achieved": c["flop_per_sample_frame"] * sf_per_s / 1e12, "peak": 2 * slots / 1e12, "unit": "TFLOP/s",
            "frac": c["lane_ops_per_sample_frame"] * sf_per_s / slots,
            "flop_per_sample_frame": c["flop_per_sample_frame"], "lane_ops_per_sample_frame": c["lane_ops_per_sample_frame"],
            "packed_fp32_instructions_per_thread_frame": c["packed_fp32_per_thread_frame"],
            "peak_source": f"128 FP32 lanes x 148 SMs x {mhz:.0f} MHz (clock sampled during the run), 2 flop per lane slot",
            "source": c["source"]}


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    from tomatis_audio_processor_b200 import synth, _lib as L
    from tomatis_audio_processor_b200.batch import DeviceBatch, HostBatchPipeline

    T = args.tracks_per_gpu
    x = synth.device_batch(T, N_SAMPLES, SR, 1000 + rank * T, dev)       # seeds 1000.. as in SURVEY 8(d) C4
    y = torch.empty_like(x)
    db = DeviceBatch(x, y, SR, "standard", device=local, unit_blocks=args.unit_blocks, gate_ui=50)
    sf_rank = T * N_SAMPLES
    audio_s_rank = sf_rank / SR

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        db.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = db.plan.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for k in range(args.steps):
        db.levels(); db.gate(); db.edges()
        ev[k][0].record()
        db.stft()                      # fused STFT/OLA + per-chunk limiter
        ev[k][1].record()
    e1.record()
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    stft_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    launches = db.plan.launch_count() - l0
    tmax = torch.tensor([ms_total, stft_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total, stft_ms = float(tmax[0]), float(tmax[1])
    value = audio_s_rank * world * args.steps / (ms_total * 1e-3)
    states = db.plan.read(L.ARR_STATE)
    peaks = db.plan.read(L.ARR_CHUNK_PEAK)
    c2_frac = float((states == 2).mean())
    lim_frac = float((peaks > np.float32(0.999)).mean())
    out_peak = float(y.abs().max())
    finite = bool(torch.isfinite(y[0]).all())
    parity_in = None
    if rank == 0 and not args.no_cpu:          # track 0 of the LAST timed step goes to the oracle below (checker leg, not timed)
        nf0, nc0 = db.plan.track_frames[0], db.plan.track_chunks[0]
        parity_in = (x[0].cpu().numpy(), y[0].cpu().numpy(), states[:nf0].copy(), peaks[:nc0].copy())
    db.close()

    # ---- the reference documentation's faster setting, --n_fft 2048 --hop 1024, on the same resident tracks (the second build of
    # the library: pair mode of the same fused kernels); a short sub-record, the headline stays the default frame size
    alt = None
    if not args.no_alt_size:
        db2 = DeviceBatch(x, y, SR, "standard", device=local, unit_blocks=args.unit_blocks, gate_ui=50, n_fft=2048, hop=1024)
        for _ in range(2):
            db2.step()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_alt = max(3, args.steps // 2)
        a0.record()
        for _ in range(n_alt):
            db2.step()
        a1.record()
        barrier()
        t_alt = torch.tensor([a0.elapsed_time(a1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_alt, op=dist.ReduceOp.MAX)
        alt = {"n_fft": 2048, "hop": 1024, "value": audio_s_rank * world * n_alt / (float(t_alt[0]) * 1e-3), "unit": UNIT,
               "steps": n_alt, "ms_per_step": float(t_alt[0]) / n_alt, "output_peak": float(y.abs().max()),
               "library": "libtomatis_b200_n2048.so (TMT_NFFT=2048: two 2048-point frames per 4096-wide pass of stft_kernel)"}
        db2.close()

    # ---- end to end through the host-buffer API (pinned host memory, copies inside the timed region)
    e2e = None
    e2e_pcm = None
    if not args.no_e2e:
        # same tracks per GPU as `value` on one GPU (27 GB of pinned host memory); with several ranks on one host the pinned
        # footprint would be N x 27 GB, so the multi-GPU runs keep a 32-track sample per rank and say so
        want = args.e2e_tracks if args.e2e_tracks > 0 else (T if world == 1 else 32)
        e2e_note = None
        while True:
            Te = min(T, want)
            Te -= Te % args.wave_tracks
            Te = max(Te, args.wave_tracks)
            try:
                h_in = torch.empty((Te, N_SAMPLES, 2), dtype=torch.float32, pin_memory=True)
                h_out = torch.empty((Te, N_SAMPLES, 2), dtype=torch.float32, pin_memory=True)
                break
            except RuntimeError as exc:         # not enough pinnable host memory on this box: smaller sample, stated in the line
                if want <= 32:
                    raise
                e2e_note = f"pinned host allocation for {Te} tracks failed ({str(exc)[:80]}); 32-track sample"
                want = 32
        if Te != T and e2e_note is None:
            e2e_note = (f"{Te}-track sample per rank ({world} ranks share one host: {T} tracks per rank would pin "
                        f"{2 * T * N_SAMPLES * 8 * world / 1e9:.0f} GB)") if world > 1 else f"{Te}-track sample (--e2e-tracks)"
        h_in.copy_(x[:Te])
        y = None
        pipe = HostBatchPipeline(N_SAMPLES, SR, "standard", device=local, wave_tracks=args.wave_tracks,
                                 unit_blocks=args.unit_blocks, gate_ui=50)
        for _ in range(max(1, min(args.warmup, 2))):
            pipe.process(h_in, h_out)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.e2e_steps):
            pipe.process(h_in, h_out)
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        e2e_value = Te * N_SAMPLES / SR * world * args.e2e_steps / (float(ms[0]) * 1e-3)
        # spot check: the host result equals the resident result of the same tracks
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(Te * N_SAMPLES * 8 * world),
               "d2h_bytes_per_step": int(Te * N_SAMPLES * 8 * world), "tracks_per_gpu_per_step": Te,
               "steps": args.e2e_steps, "ms_per_step": float(ms[0]) / args.e2e_steps,
               "api": "tomatis_audio_processor_b200.batch.HostBatchPipeline.process (pinned host float32 in/out, "
                      f"waves of {args.wave_tracks} tracks, H2D/compute/D2H on 3 streams)",
               "host_peak_out": max(float(h_out[i].abs().max()) for i in sorted({0, Te // 2, Te - 1})), "note": e2e_note}
        pipe.close()
        # same pipeline with integer PCM crossing PCIe (int16 in as decoded from a 16-bit source, PCM_24 out as the
        # reference's output files store it): informational, the headline e2e above moves float32 like the reference arm
        if not args.no_e2e_pcm:
            s_in = torch.empty((Te, N_SAMPLES, 2), dtype=torch.int16, pin_memory=True)
            s_in.copy_((h_in * 32767.0).round_().to(torch.int16))
            del h_in, h_out
            s_out = torch.empty((Te, N_SAMPLES, 6), dtype=torch.uint8, pin_memory=True)
            pipe = HostBatchPipeline(N_SAMPLES, SR, "standard", device=local, wave_tracks=args.wave_tracks, unit_blocks=args.unit_blocks,
                                     in_format="s16", out_format="s24", gate_ui=50)
            pipe.process(s_in, s_out)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.e2e_steps):
                pipe.process(s_in, s_out)
            b.record()
            barrier()
            ms = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            e2e_pcm = {"value": Te * N_SAMPLES / SR * world * args.e2e_steps / (float(ms[0]) * 1e-3), "unit": UNIT,
                       "h2d_bytes_per_step": int(Te * N_SAMPLES * 4 * world), "d2h_bytes_per_step": int(Te * N_SAMPLES * 6 * world),
                       "ms_per_step": float(ms[0]) / args.e2e_steps,
                       "api": "HostBatchPipeline(in_format='s16', out_format='s24'): int16 PCM in, packed PCM_24 out, conversions on the device"}
            pipe.close()
            del s_in, s_out
        else:
            del h_in, h_out

    # ---- the second named scaling target in the same line: one 2-hour 96 kHz file (BASELINE configs[4]), strong scaling
    lf = None
    if not args.no_longfile:
        x = y = None
        torch.cuda.empty_cache()
        if world == 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(dev))
        lf = longfile_measure(args, dist, rank, world, local, args.steps, False)
        for k in ("metric", "unit", "higher_is_better", "vs_baseline", "dtype", "data", "warmup"):
            lf.pop(k, None)

    if rank == 0:
        peak_gbs, peak_src = measured_peak()
        achieved = ALG_BYTES_PER_SF * sf_rank / (stft_ms * 1e-3) / 1e9
        traffic = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(T, world), "clocks": clocks,
            "e2e": e2e, "e2e_pcm": e2e_pcm, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "stft_kernel (fused gather+window+FFT+gain+IFFT+window+OLA+peak+chunk limiter)",
                         "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": ALG_BYTES_PER_SF * sf_rank,
                         "kernel_ms": stft_ms, "kernel_share_of_step": stft_ms * args.steps / ms_total,
                         "traffic": (traffic or {}).get("dram_bytes_per_sf", None) and
                                    (traffic["dram_bytes_per_sf"] * sf_rank),
                         "traffic_source": (traffic or {}).get("source"),
                         "secondary": fp32_secondary(sf_rank / (stft_ms * 1e-3), (clocks or {}).get("sm_mhz"))},
            "checks": {"c2_fraction": c2_frac, "chunks_over_limit": lim_frac, "output_peak": out_peak,
                       "finite": finite},
        }
        if parity_in is not None:
            line["checks"].update(parity_vs_oracle(*parity_in))
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_single(args.cpu_tracks)
        if alt is not None:
            line["frame_size_2048_1024"] = alt
        if lf is not None:
            line["longfile"] = lf
        print(json.dumps(line), flush=True)
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ long file (configs[4])
LF_SR = 96000
LF_SECONDS = 7200.0
LF_TOTAL = int(LF_SR * LF_SECONDS)            # 691 200 000 sample-frames = 2048 * 337 500 (pad_end = 0)


def longfile_config(n_gpus, total=LF_TOTAL):
    return {
        "workload": f"BASELINE configs[4]: single {total / LF_SR / 3600:.2f}-hour 96 kHz stereo synthetic file, process_tomatis standard mode "
                    f"--gate_ui 50, time-chunk sharded across {n_gpus} GPU(s) on limiter-chunk boundaries with one-hop halos",
        "mode": "standard", "gate_ui": 50, "sample_rate": LF_SR, "file_seconds": total / LF_SR, "n_fft": 4096, "hop": 2048,
        "l2_policy": f"inputs larger than L2 ({total * 8 / 1e9:.2f} GB in + out per pass, streamed once)",
        "parallelism": f"time-chunk sharded x{n_gpus} on limiter-chunk boundaries; per pass every rank's hop-block sums go to every rank "
                       f"({total // 2048 * 4 / 1e6:.2f} MB assembled) for the redundant gate scan and one hop (16 KB) to each neighbour: "
                       "peer-memory stores + flags over NVLink (CUDA IPC) when available, else NCCL all-gather + all-reduce; "
                       "the per-chunk limiter is rank-local (fused in stft_kernel)",
    }


def cpu_baseline_longfile(excerpt_seconds=240.0):
    """Oracle port on ONE host core (the reference processes one file on one thread), bounded excerpt."""
    from tomatis_audio_processor_b200 import synth
    from oracle import tomatis_oracle as orc
    x = synth.recipe_gated_pink(excerpt_seconds, LF_SR, 5)
    t = time.perf_counter()
    orc.process_standard(x, LF_SR, gate_ui=50)
    dt = time.perf_counter() - t
    return {"value": excerpt_seconds / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{excerpt_seconds:.0f} s excerpt @ {LF_SR} Hz through oracle.process_standard, one process (the reference's "
                      f"streaming loop is single-threaded and sequential in the gate state)"}


def run_longfile_reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    times = []
    for it in range(args.warmup + args.steps):
        cb = cpu_baseline_longfile(args.lf_cpu_seconds)
        if it >= args.warmup:
            times.append(args.lf_cpu_seconds / cb["value"])
    value = args.lf_cpu_seconds * len(times) / sum(times)
    cb["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sum(times) / len(times) * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": longfile_config(args.gpus, args.lf_total),
            "cpu_baseline": cb, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def longfile_measure(args, dist, rank, world, local, steps, want_cpu):
    """One 2-hour 96 kHz file, time-chunk sharded over the ranks of the (already initialised) process group; every rank
    returns the record (only rank 0's is printed).  Timed like the batch arm: barrier + synchronize on both sides, CUDA
    events, max over ranks."""
    import numpy as np
    import torch

    dev = f"cuda:{local}"
    from tomatis_audio_processor_b200 import sharded, synth, _lib as L
    total = args.lf_total
    comm = sharded.Comm(None, dev)
    me = sharded.plan_shards(total, world, sharded.STREAMING)[rank]
    own = synth.device_long_file_range(me.own_lo, me.own_hi, LF_SR, 5000, dev)
    sess = sharded.StreamingShardSession("standard", own, LF_SR, total, comm, device_index=local, unit_blocks=args.unit_blocks, gate_ui=50)
    own = sess.own                      # the session keeps the samples inside its window buffer
    plan = sess.be.plan

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sess.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = plan.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for k in range(steps):
        sess.step()                         # CUDA-graph replay of the whole pass (kernels + collectives)
    e1.record()
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    launches_per_step = None
    for k in range(steps):                  # the same pass run eagerly: STFT kernel time (roofline) and the launch count
        lk = plan.launch_count()
        sess.step(ev[k])
        launches_per_step = plan.launch_count() - lk
    torch.cuda.synchronize()
    stft_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    launches = (launches_per_step or 0) * steps
    marks = []
    bytes0 = comm.bytes_sent
    sess.step(marks=marks)
    comm_step_bytes = comm.bytes_sent - bytes0
    torch.cuda.synchronize()
    breakdown = {b[0]: round(a[1].elapsed_time(b[1]), 4) for a, b in zip(marks, marks[1:])}
    tmax = torch.tensor([ms_total, stft_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total, stft_ms = float(tmax[0]), float(tmax[1])
    value = total / LF_SR * steps / (ms_total * 1e-3)
    states = plan.read(L.ARR_STATE)
    peaks = plan.read(L.ARR_CHUNK_PEAK)
    out_peak = float(sess.out.abs().max()) if sess.out.numel() else 0.0

    # ---- end to end: the rank's own range from pinned host memory and back, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        n_own = me.own_hi - me.own_lo
        h_in = torch.empty((n_own, 2), dtype=torch.float32, pin_memory=True)
        h_out = torch.empty((n_own, 2), dtype=torch.float32, pin_memory=True)
        h_in.copy_(own)

        def e2e_step():
            own.copy_(h_in, non_blocking=True)
            sess.step()
            h_out.copy_(sess.out, non_blocking=True)
        e2e_step()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.e2e_steps):
            e2e_step()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        e2e = {"value": total / LF_SR * args.e2e_steps / (float(ms[0]) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(total * 8), "d2h_bytes_per_step": int(total * 8), "steps": args.e2e_steps,
               "ms_per_step": float(ms[0]) / args.e2e_steps,
               "api": "sharded.StreamingShardSession.step with each rank's own range copied from / to pinned host float32 buffers",
               "host_peak_out": float(h_out.abs().max()) if n_own else 0.0}
        if world == 1:
            # one GPU: the streamed form of the same path -- time slabs on limiter-chunk boundaries through three device slots,
            # copy-in / kernels / copy-out of consecutive slabs on three streams (streamed.HostFileStreamer), bounded device memory
            from tomatis_audio_processor_b200.streamed import HostFileStreamer
            whole = h_out.clone()
            st = HostFileStreamer("standard", total, LF_SR, device=local, slab_seconds=args.lf_slab_seconds,
                                  unit_blocks=args.unit_blocks, gate_ui=50)
            st.process(h_in, h_out)
            barrier()
            a.record()
            for _ in range(args.e2e_steps):
                st.process(h_in, h_out)
            b.record()
            barrier()
            ms_st = a.elapsed_time(b)
            e2e = {"value": total / LF_SR * args.e2e_steps / (ms_st * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": int(sum(sh.in_hi - sh.in_lo for sh in st.shards) * 8), "d2h_bytes_per_step": int(total * 8),
                   "steps": args.e2e_steps, "ms_per_step": ms_st / args.e2e_steps,
                   "api": f"streamed.HostFileStreamer.process: pinned host float32 in / out, {len(st.slabs)} time slabs of "
                          f"~{args.lf_slab_seconds:.0f} s through {len(st.slots)} device slots ({st.device_bytes() / 1e9:.2f} GB of HBM), "
                          "H2D / kernels / D2H on 3 streams",
                   "host_peak_out": float(h_out.abs().max()), "equals_whole_file_output": bool(torch.equal(whole, h_out)),
                   "whole_file_in_hbm": {"ms_per_step": e2e["ms_per_step"], "value": e2e["value"], "api": e2e["api"]}}
            st.close()
            del whole
        del h_in, h_out

    peak_gbs, peak_src = measured_peak()
    sf_rank = me.own_hi - me.own_lo
    achieved = ALG_BYTES_PER_SF * sf_rank / (stft_ms * 1e-3) / 1e9
    rec = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": longfile_config(world, total), "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches), "comm_bytes_per_step_rank0": int(comm_step_bytes),
        "roofline": {"bound": "hbm", "kernel": "stft_kernel (fused gather+window+FFT+gain+IFFT+window+OLA+peak), rank 0 shard",
                     "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ALG_BYTES_PER_SF * sf_rank, "kernel_ms": stft_ms,
                     "kernel_share_of_step": stft_ms * steps / ms_total},
        "phase_ms_rank0": breakdown, "graph_replay": bool(getattr(sess, "_graph", None) is not None),
        "exchange": ("none (one rank)" if world == 1 else "peer memory (publish / wait kernels, CUDA IPC over NVLink)"
                     if getattr(sess, "peer", None) is not None else "NCCL collectives (all-gather + all-reduce)"),
        "peer_status": (sess.peer.status() if getattr(sess, "peer", None) is not None else None),
        "checks": {"c2_fraction": float((states == 2).mean()), "chunks_over_limit": float((peaks > np.float32(0.999)).mean()),
                   "output_peak": out_peak, "frames": int(states.size)},
    }
    if want_cpu and rank == 0:
        rec["cpu_baseline"] = cpu_baseline_longfile(args.lf_cpu_seconds)
    sess.close()
    del own, sess
    torch.cuda.empty_cache()
    return rec


def run_longfile_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))
    rec = longfile_measure(args, dist, rank, world, local, args.steps, world == 1 and not args.no_cpu)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--tracks-per-gpu", dest="tracks_per_gpu", type=int, default=128)
    ap.add_argument("--unit-blocks", dest="unit_blocks", type=int, default=0)
    ap.add_argument("--e2e-tracks", dest="e2e_tracks", type=int, default=0, help="0 = all tracks of the rank on one GPU, 32 per rank otherwise")
    ap.add_argument("--e2e-steps", dest="e2e_steps", type=int, default=3)
    ap.add_argument("--wave-tracks", dest="wave_tracks", type=int, default=2)
    ap.add_argument("--cpu-tracks", dest="cpu_tracks", type=int, default=3)
    ap.add_argument("--cpu-workers", dest="cpu_workers", type=int, default=0)
    ap.add_argument("--workload", choices=["batch", "longfile"], default="batch",
                    help="batch = BASELINE configs[3] (default, weak scaling); longfile = configs[4] (one 2 h file, strong scaling)")
    ap.add_argument("--lf-total", dest="lf_total", type=int, default=LF_TOTAL, help="long-file length in sample-frames")
    ap.add_argument("--lf-cpu-seconds", dest="lf_cpu_seconds", type=float, default=240.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-pcm", dest="no_e2e_pcm", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-longfile", dest="no_longfile", action="store_true", help="skip the long-file sub-record of the default line")
    ap.add_argument("--no-alt-size", dest="no_alt_size", action="store_true", help="skip the --n_fft 2048 --hop 1024 sub-record")
    ap.add_argument("--lf-slab-seconds", dest="lf_slab_seconds", type=float, default=300.0,
                    help="long file, one GPU, end to end: slab length of the streamed pipeline")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3                       # timing rule: at least 3 warm-up steps
    if args.workload == "longfile":
        (run_longfile_reference_arm if args.impl == "reference" else run_longfile_arm)(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

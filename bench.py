#!/usr/bin/env python3
"""Benchmark of the HPSS feature front-end (BASELINE.json metric: audio-seconds per second).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference's algorithm

Workload (BASELINE.json configs[1]): a batch of 4096 synthetic 1 s 16 kHz segments per GPU,
n_fft = 400, hop = 160, median kernels 31/31, 120 mel bands, feature LogMelHarmPercSpec.
One step = waveform -> (240, 98) float32 featuregram for every clip of the batch + the raw
feature moments of get_data_stats (hpss_featuregram_moments) (all-reduced over ranks when N > 1: the only collective).

  value  device-resident waveform -> device-resident features, CUDA events on the launching
         stream, max over ranks; inputs (262 MB) + intermediates (1.3 GB/step) exceed the
         126 MB L2, so no explicit flush is needed between iterations.
  e2e    the same through the host-buffer C-ABI entry (hpss_featuregram_host): pinned host
         waveform in, pinned host features out, H2D and D2H inside the timed region.
  roofline / stages   per-kernel CUDA-event times from a separate pass over the same batch.
  cpu_baseline        oracle (librosa's algorithm on scipy/numpy) on all host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line: while the job runs, file descriptor 1 points at stderr (NCCL prints its version
# banner with a plain printf when NCCL_DEBUG=VERSION is in the environment); emit() restores it for the result line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(fs=16000, Tw=25, Ts=10, n_fft=400, win=400, hop=160, l_harm=31, l_perc=31, n_mels=120,
           featName="LogMelHarmPercSpec", clip_samples=16000, n_clips=4096)
WORKLOAD = ("4096 x 1 s synthetic 16 kHz segments per GPU, n_fft=400 hop=160 win=400, median 31/31, "
            "n_mels=120, LogMelHarmPercSpec (BASELINE.json configs[1])")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu capture (profiles/r1_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f)[kernel]
        return int(t["read"]) + int(t["write"])
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, device copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def count_since(self, t0):
        return sum(1 for t, _ in self.lines if t >= t0)

    def stop(self, t0=None, t1=None):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for t, ln in self.lines:
            if (t0 is not None and t < t0) or (t1 is not None and t > t1):
                continue                      # only samples taken while the GPU was under this benchmark's load
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ============================================================================= CPU arm
def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0            # one CPU arm per box: the other ranks exit without work
    from oracle.cpu_baseline import CpuArm, usable_cores
    cores = usable_cores()
    per_core = args.cpu_clips_per_core or 64
    arm = CpuArm(CFG, cores=cores, n_per_worker=per_core)
    for _ in range(max(args.warmup, 1)):
        arm.step()
    times = [arm.step() for _ in range(args.steps)]
    arm.close()
    audio_s = arm.clips_per_step * CFG["clip_samples"] / CFG["fs"]
    total = sum(times)
    value = audio_s * args.steps / total
    sample = (f"{arm.clips_per_step} of the 4096 clips per step ({cores} single-threaded workers x "
              f"{per_core} clips), oracle = librosa algorithm on scipy.ndimage/numpy.fft/np.dot")
    line = {
        "impl": "reference", "metric": "audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 FFT)",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def cpu_baseline_block(args):
    from oracle.cpu_baseline import CpuArm, usable_cores
    cores = usable_cores()
    arm = CpuArm(CFG, cores=cores, n_per_worker=args.cpu_clips_per_core or 256)
    arm.step()
    reps = 2
    t = sum(arm.step() for _ in range(reps))
    arm.close()
    audio_s = arm.clips_per_step * CFG["clip_samples"] / CFG["fs"] * reps
    return {"value": audio_s / t, "unit": "audio-s/s", "cores": cores, "kind": "port",
            "sample": f"{arm.clips_per_step} clips x {reps} passes of the same 1 s workload on {cores} single-threaded "
                      f"workers ({t:.1f} s of wall time); oracle = librosa algorithm on scipy/numpy"}


def bind_to_gpu_numa_node(local):
    """Best effort: run this rank (and first-touch its pinned host buffers) on the CPUs of the NUMA node its GPU
    hangs off, so that the H2D / D2H traffic of the e2e leg does not cross the socket interconnect."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{dev}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


# ============================================================================= GPU arm
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_block(args)        # before CUDA is initialised in this process

    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sm_hpss_mtl_b200 import engine, synth
    from sm_hpss_mtl_b200.dist import allreduce_moments

    ctx = engine.get_context(local)
    n_clips, L = args.clips, CFG["clip_samples"]
    prm = engine.make_params(n_fft=CFG["n_fft"], win_length=CFG["win"], hop_length=CFG["hop"], l_harm=CFG["l_harm"],
                             l_perc=CFG["l_perc"], n_mels=CFG["n_mels"], mel_sr=22050, feature="LOGMEL_HARMPERC")
    batch = engine.Batch(ctx, clip_lengths=[L] * n_clips, n_fft=CFG["n_fft"], hop_length=CFG["hop"])
    D = engine.feature_rows(prm)
    F = CFG["n_fft"] // 2 + 1
    M = CFG["n_mels"]
    frames = batch.total_frames
    audio_s = n_clips * L / CFG["fs"]

    # rank r owns the contiguous slice [r*n_clips, (r+1)*n_clips) of the synthetic corpus
    wave_host = engine.host_alloc(n_clips * L)
    wave_host[:] = synth.synth_batch_fast(n_clips, L, first_index=rank * n_clips).ravel()
    out_host = engine.host_alloc(D * frames)
    wave = torch.from_numpy(wave_host).cuda()
    out = torch.empty(D * frames, dtype=torch.float32, device="cuda")
    classes = (np.arange(n_clips) % 3).astype(np.int32)
    acc = torch.zeros(3 * D + D + 3 + 1, dtype=torch.float64, device="cuda")

    def step():
        acc.zero_()
        engine.featuregram_moments(batch, wave, prm, classes, 3, out=out, acc=acc)
        if world > 1:
            allreduce_moments(acc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()                      # nvidia-smi needs a moment to come up: start it before the warm-up
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    t_load0 = time.perf_counter()
    l0 = engine.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = engine.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    # keep the sampler running over the e2e region as well (both are "under load")

    # ---- e2e: host buffers through the C-ABI host entry
    for _ in range(2):
        engine.featuregram_host(batch, wave_host, prm, out_host)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        engine.featuregram_host(batch, wave_host, prm, out_host)
        _ = float(out_host[0])                     # the result is in host memory
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    extra = False
    t_wait = time.perf_counter()
    while sampler.proc is not None and sampler.count_since(t_load0) < 3 and time.perf_counter() - t_wait < 3.0:
        # short runs: keep the same kernels running until nvidia-smi has reported (untimed; no collective here, the
        # ranks leave this loop at different times)
        engine.featuregram_moments(batch, wave, prm, classes, 3, out=out, acc=acc)
        torch.cuda.synchronize()
        extra = True
    clocks = sampler.stop(t_load0, time.perf_counter())
    if extra:
        clocks["note"] = "timed region shorter than the sampling period: the same step was kept running (untimed) until 3 samples arrived"

    # ---- per-stage pass (rank 0 only reports it): same batch, one CUDA-event pair per kernel
    stages = None
    if rank == 0:
        fused = os.environ.get("HPSS_USE_FUSED", "0") not in ("", "0")     # the path hpss_featuregram takes
        if fused:
            names = ["K1 stft_mag", "K2h median_time", "K2p+K3 perc_mask_mel_log", "K3b+K5 topdb_moments"]
            bytes_per_frame = [4 * CFG["hop"] + 4 * F, 8 * F, 8 * F + 8 * M, 8 * M]
        else:
            names = ["K1 stft_mag", "K2h median_time", "K2p median_freq", "K3 mask_mel_log", "K3b+K5 topdb_moments"]
            # K3b+K5 reads every feature once and writes back only the values the top_db clip changes: 8*M, not 16*M
            bytes_per_frame = [4 * CFG["hop"] + 4 * F, 8 * F, 8 * F, 12 * F + 8 * M, 8 * M]
        tot = [0.0] * len(names)
        reps = max(args.steps, 5)
        for it in range(reps + 2):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            evs[0].record()
            S = engine.stft_mag(batch, wave, CFG["n_fft"], CFG["win"], CFG["hop"]); evs[1].record()
            harm = engine.median_time(batch, S, F, CFG["l_harm"]); evs[2].record()
            if fused:
                o, cmax = engine.perc_mask_mel_log(batch, S, harm, F, CFG["l_perc"], 22050, M, log_power=1); evs[3].record()
                nxt = 4
            else:
                perc = engine.median_freq(batch, S, F, CFG["l_perc"]); evs[3].record()
                o, cmax = engine.mask_mel_log(batch, S, harm, perc, F, mel_sr=22050, n_mels=M, log_power=1); evs[4].record()
                nxt = 5
                del perc
            acc.zero_()
            engine.topdb_moments(batch, o, M, 2, cmax, 80.0, classes, 3, acc=acc); evs[nxt].record()
            torch.cuda.synchronize()
            if it >= 2:
                for i in range(len(names)):
                    tot[i] += evs[i].elapsed_time(evs[i + 1])
            del S, harm, o, cmax
        peak, peak_src = measured_peak_gbs()
        stages = []
        for i, nme in enumerate(names):
            ms = tot[i] / reps
            gbs = bytes_per_frame[i] * frames / (ms * 1e-3) / 1e9
            stages.append({"kernel": nme, "ms": round(ms, 4), "algorithmic_bytes_per_frame": bytes_per_frame[i],
                           "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 4)})
        dom = max(stages, key=lambda s: s["ms"])
        roofline = {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak,
                    "unit": "GB/s", "frac": dom["frac_of_hbm_peak"], "traffic": measured_traffic(dom["kernel"]),
                    "algorithmic_bytes": dom["algorithmic_bytes_per_frame"] * frames, "peak_source": peak_src,
                    "traffic_source": "profiles/r1_traffic.json (ncu --set full of the same kernels, bytes per launch)",
                    "note": "the two median kernels are bound by the ALU pipe (half-rate FMNMX selection networks: "
                            "profiles/README.md), HBM is their secondary bound; K1/K3/K3b+K5 are the HBM-side "
                            "stages; per-stage numbers in 'stages'"}

    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = world * audio_s / (ms_per_step * 1e-3)
        e2e_value = world * audio_s * e2e_steps / e2e_s
        line = {
            "metric": "audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_gpu": n_clips, "frames_per_gpu": frames,
                       "l2": "inputs+intermediates per step (1.6 GB) exceed the 126 MB L2; no explicit flush",
                       "step": "hpss_featuregram_moments = K1, K2h, K2p, K3, K3b+K5 (top_db clip and moments share one pass)"
                               + (" + NCCL all-reduce of the 968-double moment vector" if world > 1 else "")},
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(wave_host.nbytes),
                    "d2h_bytes_per_step": int(out_host.nbytes), "steps": e2e_steps,
                    "api": "hpss_featuregram_host (pinned host buffers, chunked H2D/compute/D2H pipeline)",
                    "rank0_numa_node": numa_node},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "stages": stages,
        }
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=CFG["n_clips"], help="clips per GPU (default: the named 4096)")
    ap.add_argument("--cpu-clips-per-core", type=int, default=0,
                    help="clips per worker and pass of the CPU legs (default: 256 for cpu_baseline = the whole 4096-clip "
                         "workload on 16 cores, ~15 s; 64 per step for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(v, "1")
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())

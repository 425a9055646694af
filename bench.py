#!/usr/bin/env python
"""Headline benchmark: RVQ encode+decode frames/s on synthetic latents (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A *step* is one pass of the hot path over one batch: encode (nearest-codeword search over all
stages) followed by decode (codebook gather-accumulate).  Default workload = BASELINE.json
configs[1] (`cfg2_enc24k_32d_vq1`: Encodec_24k_32d single-codebook VQ on long sequences,
x [8, 512, 45000] per GPU, K = 1024).  Multi-GPU runs shard clips across ranks (no collective on
the data path): every rank processes its own batch, so scaling is weak.

JSON keys beyond the base contract:
  roofline      dominant kernel (the search) -- algorithmic flops / CUDA-event time vs the
                measured bf16 tensor peak of MEASURED_PEAKS.json
  cpu_baseline  the oracle port (= the reference's own ATen calls) timed on this host's cores
  e2e           the same step through the host-buffer C ABI (acq_*_host): pinned host latents
                -> H2D -> kernels -> D2H, copies inside the timed region
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from academicodec_b200 import synth  # noqa: E402

METRIC = "rvq_encode_decode_frames_per_sec"
UNIT = "frames/s"


# ------------------------------------------------------------------------------------ utilities
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_burst=p["bf16_tflops"],
                    bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def dist_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def workload(name: str):
    w = dict(synth.WORKLOADS[name])
    w["name"] = name
    return w


def algorithmic(w):
    """Per-frame algorithmic work (SURVEY.md 8d / BASELINE.md section 5)."""
    g = w.get("G", 1)
    s = w["n_q"]
    flops_enc = 2.0 * w["bins"] * w["D"] * s          # 2*K*Dg*G*S
    bytes_enc = 4.0 * w["D"] + 8.0 * s * g
    bytes_dec = 8.0 * s * g + 4.0 * w["D"]
    return flops_enc, bytes_enc, bytes_dec


# ------------------------------------------------------------------------------------ CPU arm
def cpu_port_throughput(w, clips: int, reps: int, warm: int = 1, device: str = "cpu"):
    """Time the oracle port (the reference's ATen call sequence) on this host's cores for
    `clips` clips of the workload; returns frames/s of encode+decode and the thread count.
    device="cuda" runs the same eager op sequence on the GPU (SURVEY.md 8d's secondary baseline:
    the reference's algorithm as PyTorch eager kernels on the same B200)."""
    from oracle import grvq_oracle, rvq_oracle
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    t_frames = w["T"]
    x = torch.from_numpy(synth.latents(clips, w["D"], t_frames, 1234)).to(device)
    if w["kind"] == "grvq":
        ws = synth.grvq_codebooks(w["G"], w["bins"], 777, "randn")
        ws = [[torch.from_numpy(a).to(device) for a in st] for st in ws]

        def step():
            q, loss, ids = grvq_oracle.grvq_forward(x, ws)
            codes = torch.stack(ids, -1).reshape(clips, t_frames, -1)
            return grvq_oracle.grvq_embed(codes, ws)
    else:
        cb = list(torch.from_numpy(synth.rvq_codebooks(w["n_q"], w["bins"], w["D"], 4321, "decay")).to(device))

        def step():
            codes = rvq_oracle.rvq_encode(x, cb)
            return rvq_oracle.rvq_decode(codes, cb)
    sync = torch.cuda.synchronize if device != "cpu" else (lambda: None)
    with torch.no_grad():
        for _ in range(warm):
            step()
        sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            step()
        sync()
        dt = (time.perf_counter() - t0) / reps
    return clips * t_frames / dt, cores, dt


def run_reference(args):
    world, rank, _ = dist_env()
    if rank != 0:
        return 0
    w = workload(args.workload)
    clips = max(1, min(w["B"], args.ref_clips))
    # each step = `clips` clips of the workload through the CPU port
    vals = []
    fps, cores, dt = cpu_port_throughput(w, clips, reps=max(1, args.steps), warm=max(1, min(args.warmup, 2)))
    vals.append(fps)
    sample = f"{clips} of {w['B']} clips [{w['D']}x{w['T']}] per step, oracle port (torch CPU, {cores} threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "audio_sec_per_sec": fps / w["frame_rate"],
        "config": {"workload": w["name"], "D": w["D"], "n_q": w["n_q"], "bins": w["bins"],
                   "clips_per_step": clips, "frames_per_clip": w["T"]},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    world, rank, local = dist_env()
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; academicodec_b200 has no CPU path"}))
        return 2
    import torch.distributed as dist
    from academicodec_b200 import _lib, ops
    _lib.load()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = workload(args.workload)
    b, d, t, s, k = w["B"], w["D"], w["T"], w["n_q"], w["bins"]
    g = w.get("G", 1)
    grvq = w["kind"] == "grvq"
    n_frames = b * t
    flags = (ops.ACQ_STE | ops.ACQ_LOSS_RAW) if grvq else 0

    # ---- synthetic inputs: generated on the host, pinned (the e2e leg starts from them) --------
    x_host = torch.from_numpy(synth.latents(b, d, t, 1234 + rank)).pin_memory()
    if grvq:
        ws = synth.grvq_codebooks(g, k, 777, "randn")
        cbs = [torch.from_numpy(a).to(dev) for st in ws for a in st]
    else:
        cbs = [c.contiguous() for c in torch.from_numpy(synth.rvq_codebooks(s, k, d, 4321, "decay")).to(dev)]
    hn = ops.codebook_half_norms(cbs)
    pack = None
    if args.kernel != 1 and ops.tc_supported(k, d, g):
        pack = ops.tc_pack_codebooks(cbs)      # tcgen05 operand images (built once per codebook)
    x_dev = x_host.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    codes_dev = torch.empty((s * g, n_frames), dtype=torch.int64, device=dev)
    out_dev = torch.empty((b, d, t), dtype=torch.float32, device=dev)

    def step_resident():
        ops.rvq_search(x_dev, cbs, s, g, half_norms=hn, flags=flags, impl=args.kernel, tc_pack=pack,
                       codes_out=codes_dev)
        ops.vq_decode(codes_dev, n_frames, 1, cbs, s, g, b, t, check=False, out=out_dev)

    for _ in range(max(3, args.warmup)):
        step_resident()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    # ---- timed region: K steps, CUDA events on the launching stream ----------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    torch.cuda.synchronize()
    start = torch.cuda.Event(enable_timing=True)
    stop = torch.cuda.Event(enable_timing=True)
    start.record()
    host_t0 = time.perf_counter()
    for i in range(args.steps):
        ev[i][0].record()
        ops.rvq_search(x_dev, cbs, s, g, half_norms=hn, flags=flags, impl=args.kernel, tc_pack=pack,
                       codes_out=codes_dev)
        ev[i][1].record()
        ops.vq_decode(codes_dev, n_frames, 1, cbs, s, g, b, t, check=False, out=out_dev)
        ev[i][2].record()
    host_ms = (time.perf_counter() - host_t0) * 1e3 / args.steps     # launch cost per step
    stop.record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = start.elapsed_time(stop)
    enc_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    dec_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    t_ms = torch.tensor([total_ms, enc_ms, dec_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms, enc_ms, dec_ms = t_ms.tolist()
    ms_per_step = total_ms / args.steps
    value = world * n_frames / (ms_per_step * 1e-3)

    # ---- effective SM clock inside the search kernel (outside the timed region) -----------------
    # nvidia-smi's 100 ms samples cannot resolve a 20 ms timed region; the tensor-core kernel counts its
    # own cycles (ACQ_TC_DBG bit 512: clock64 in its MMA-issuing thread), which with the event time of
    # the same launches gives the clock it really ran at.  Sustained, this kernel sits at the board's
    # software power cap (scripts/power_probe.py: ~1.0-1.2 GHz at 990-1000 W).
    sm_mhz_in_kernel = None
    if rank == 0 and pack is not None and not os.environ.get("ACQ_TC_KERNEL"):
        wsp = ops.tc_workspace(d, dev)
        base = int(_lib.load().acq_tc_workspace_bytes(d)) - 256 + 64
        prev = os.environ.get("ACQ_TC_DBG")
        os.environ["ACQ_TC_DBG"] = "512"
        try:
            wsp[base:base + 72].zero_()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nk = 5
            k0.record()
            for _ in range(nk):
                ops.rvq_search(x_dev, cbs, s, g, half_norms=hn, flags=flags, impl=args.kernel, tc_pack=pack,
                               codes_out=codes_dev)
            k1.record()
            torch.cuda.synchronize()
            cyc = wsp[base:base + 72].view(torch.int64)[7].item() / (nk * min(148, (n_frames + 127) // 128))
            sm_mhz_in_kernel = cyc / (k0.elapsed_time(k1) / nk) / 1e3
        finally:
            if prev is None:
                os.environ.pop("ACQ_TC_DBG", None)
            else:
                os.environ["ACQ_TC_DBG"] = prev

    # ---- end to end through the host-buffer C ABI -------------------------------------------------
    pipe = ops.HostPipeline(local, args.chunk_mb << 20)
    codes_host = torch.empty((s * g, n_frames), dtype=torch.int64).pin_memory()
    out_host = torch.empty((b, d, t), dtype=torch.float32).pin_memory()

    def step_e2e():
        # one public call: host latents -> codes + reconstructed latents on the host
        pipe.rvq_codec(x_host, cbs, s, g, hn, flags=flags, impl=args.kernel, codes_out=codes_host,
                       out=out_host, tc_pack=pack)
        return pipe.last_launches

    e2e_launches = 0
    for _ in range(2):
        e2e_launches = step_e2e()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * n_frames / float(e2e_s.item())
    e2e_ok = bool(torch.equal(codes_host, codes_dev.cpu()) and torch.equal(out_host, out_dev.cpu()))
    h2d = x_host.numel() * 4
    d2h = codes_host.numel() * 8 + out_host.numel() * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (the search) ------------------------------------------
    peaks = load_peaks()
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and pack is not None:
        traffic = json.load(open(tpath)).get(w["name"], {})

    def dram_traffic(kernel):
        t = traffic.get(kernel)
        return (t["dram_read"] + t["dram_write"]) if t else None

    flops_enc, bytes_enc, bytes_dec = algorithmic(w)
    achieved_tf = flops_enc * n_frames / (enc_ms * 1e-3) / 1e12
    peak_tf = peaks["bf16_sustained"] if total_ms > 1000 else peaks["bf16_burst"]
    dec_gbs = bytes_dec * n_frames / (dec_ms * 1e-3) / 1e9
    roof = {"bound": "tensor", "kernel": "rvq_search", "achieved": achieved_tf, "peak": peak_tf,
            "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": dram_traffic("rvq_search"),
            "traffic_note": "DRAM bytes per launch from the committed ncu capture (profiles/traffic.json); "
                            "algorithmic bytes per launch = %d" % int(bytes_enc * n_frames),
            "peak_source": f"{peaks['source']} bf16 dense ({'sustained' if total_ms > 1000 else 'burst'})",
            "ms_per_launch": enc_ms,
            # the tensor-core kernel executes three fp16 MMAs per algorithmic product (hi.lo + lo.hi + hi.hi,
            # fp32-class scores): what the tensor pipe actually delivers, next to the algorithmic figure
            "executed_tflops": (3.0 * achieved_tf) if (pack is not None and args.kernel != 1
                                                         and os.environ.get("ACQ_TC_KERNEL") != "1") else achieved_tf,
            "executed_frac": ((3.0 * achieved_tf) if (pack is not None and args.kernel != 1
                                                       and os.environ.get("ACQ_TC_KERNEL") != "1")
                              else achieved_tf) / peak_tf,
            "decode": {"bound": "hbm", "kernel": "vq_decode", "achieved": dec_gbs,
                       "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": dec_gbs / peaks["hbm_gbs"],
                       "traffic": dram_traffic("vq_decode"), "ms_per_launch": dec_ms}}

    eager = None
    if world == 1 and not args.no_cpu_baseline:
        # secondary baseline: the reference's op sequence as PyTorch eager kernels on this GPU
        try:
            efps, _, edt = cpu_port_throughput(w, b, reps=3, warm=2, device=f"cuda:{local}")
            eager = {"value": efps, "unit": UNIT, "kind": "port (torch eager CUDA ops, fp32, TF32 off)",
                     "sample": f"all {b} clips, 3 timed reps after 2 warm-ups, {edt * 1e3:.1f} ms per rep"}
        except torch.cuda.OutOfMemoryError:
            eager = {"value": None, "unit": UNIT, "kind": "port (torch eager CUDA ops)", "sample": "out of memory"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        clips = max(1, min(b, args.ref_clips))
        fps, cores, dt = cpu_port_throughput(w, clips, reps=2, warm=1)
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{clips} of {b} clips [{d}x{t}], 2 timed reps after 1 warm-up, oracle port "
                         f"(torch CPU ops = the reference's ATen calls), {dt * 1e3:.0f} ms per rep"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "audio_sec_per_sec": value / w["frame_rate"],
        "config": {"workload": w["name"], "kind": w["kind"], "D": d, "n_q": s, "groups": g, "bins": k,
                   "clips_per_gpu": b, "frames_per_clip": t, "frame_rate": w["frame_rate"],
                   "l2": "inputs+outputs per step (%.0f MB) exceed the 126 MB L2" % ((h2d + d2h) / 1e6),
                   "kernel": "tcgen05" if pack is not None else "simt"},
        "encode_ms": enc_ms, "decode_ms": dec_ms, "host_launch_ms_per_step": host_ms,
        "roofline": roof, "cpu_baseline": cpu, "eager_gpu_baseline": eager,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "matches_resident": e2e_ok, "chunk_mb": args.chunk_mb,
                "api": "acq_rvq_codec_host (encode -> decode, codes stay on the device in between)"},
        "gpu_launches": 2 * args.steps + e2e_launches * e2e_steps,
        "clocks": clocks,
    }
    if clocks is not None and sm_mhz_in_kernel is not None:
        clocks["sm_mhz_in_search_kernel"] = round(sm_mhz_in_kernel, 1)
        clocks["note"] = ("in-kernel clock = the search kernel's own cycle count / its event time; it is below "
                          "sm_max_mhz because the tensor-core kernel draws the board's full power budget "
                          "(sw_power_cap when sustained, scripts/power_probe.py) -- kept and noted, not a clock lock")
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_enc24k_32d_vq1", choices=list(synth.WORKLOADS))
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 SIMT, 2 tensor-core")
    ap.add_argument("--chunk-mb", type=int, default=96,
                    help="staging bytes per in-flight chunk of the e2e pipeline; >= one clip keeps the PCIe copies 1-D")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-clips", type=int, default=2,
                    help="clips per step of the CPU arm / cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

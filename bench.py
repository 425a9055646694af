#!/usr/bin/env python
"""Headline benchmark: RVQ encode+decode frames/s on synthetic latents (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A *step* is one pass of the hot path over one batch: encode (nearest-codeword search over all
stages) followed by decode (codebook gather-accumulate).  Default workload = BASELINE.json
configs[1] (`cfg2_enc24k_32d_vq1`: Encodec_24k_32d single-codebook VQ on long sequences,
x [8, 512, 45000] per GPU, K = 1024).  Multi-GPU runs shard clips across ranks (no collective on
the data path): every rank processes its own batch, so scaling is weak.

Every BASELINE.json config is a --workload (academicodec_b200/synth.py WORKLOADS): cfg1 / cfg1-recipe / cfg2 /
cfg3 (GRVQ) / cfg4 (quantizer slice of the SoundStream codec) run the encode+decode step; the cfg5 workloads
(`*_train`) run the training forward of the drop-in module -- tcgen05 search, replay pass (straight-through sum,
commitment loss, EMA statistics), ONE NCCL all-reduce of the flat statistics buffer, EMA apply -- which is the
only collective of the path.

JSON keys beyond the base contract:
  roofline      dominant kernel (the search) -- algorithmic flops / CUDA-event time vs the
                measured bf16 tensor peak of MEASURED_PEAKS.json; `decode` = the gather vs the HBM peak
  cpu_baseline  the reference's own modules (baseline/_ref, kind "reference") -- or, if that install is absent,
                the oracle port (kind "port") -- timed on this host's cores on a bounded sample
  e2e           the same step through the host-buffer C ABI (acq_*_host): pinned host latents
                -> H2D -> kernels -> D2H, copies inside the timed region
  module        the drop-in surface itself: ResidualVectorQuantizer.encode / .decode (or Quantizer.forward /
                .embed) per call, CUDA events
  sustained     the resident step back to back for >= 2 s (power-capped steady state) with its clocks
  secondary     (N = 1) the other BASELINE configs: ms per launch, frames/s and roofline fractions
  collective    (cfg5) all-reduce time of the statistics buffer and its bus bandwidth vs NVLink
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
DEFAULT_WORKLOAD = "cfg2_enc24k_32d_vq1"
# the other BASELINE.json configs at their throughput sizes, reported under `secondary`
SECONDARY = ["cfg1_enc24k_240d_rvq", "cfg1_recipe_rvq", "cfg1_b4096", "cfg3_hifi16k_320d_grvq", "cfg3_b4096",
             "cfg4_ss24k_240d_rvq", "cfg4_b64", "cfg5_rvq_ema_train", "cfg5_b640_train"]


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
def _reference_modules():
    """The unmodified reference package from baseline/_ref (baseline/install_ref.py), or None."""
    try:
        from baseline import install_ref
        return install_ref.load()
    except Exception:
        return None


def cpu_throughput(w, clips: int, reps: int, warm: int = 1, device: str = "cpu", prefer_reference: bool = True):
    """Time the reference path for `clips` clips of the workload on this host's cores (or, device="cuda", the
    same eager op sequence on the GPU: SURVEY.md 8d's secondary baseline).  Uses the reference's own modules
    when baseline/_ref is installed (kind "reference"), the oracle port otherwise (kind "port").
    -> (frames/s of the step, threads, seconds per step, kind)"""
    import types
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    t_frames = w["T"]
    train = w["kind"] == "rvq_train"
    x = torch.from_numpy(synth.latents(clips, w["D"], t_frames, 1234)).to(device)
    ref = _reference_modules() if prefer_reference else None
    kind = "reference" if ref is not None else "port"
    if w["kind"] == "grvq":
        ws = synth.grvq_codebooks(w["G"], w["bins"], 777, "randn")
        ws = [[torch.from_numpy(a).to(device) for a in st] for st in ws]
        if ref is not None:
            from academicodec.models.hificodec.models import Quantizer as RefQuantizer
            h = types.SimpleNamespace(n_code_groups=w["G"], n_codes=w["bins"], codebook_loss_lambda=1.0,
                                      commitment_loss_lambda=0.25)
            q = RefQuantizer(h).to(device)
            with torch.no_grad():
                for g in range(w["G"]):
                    q.quantizer_modules[g].embedding.weight.copy_(ws[0][g])
                    q.quantizer_modules2[g].embedding.weight.copy_(ws[1][g])

            def step():
                qo, loss, ids = q(x)
                codes = torch.stack(ids, -1).reshape(clips, t_frames, -1)
                return q.embed(codes)
        else:
            from oracle import grvq_oracle

            def step():
                qo, loss, ids = grvq_oracle.grvq_forward(x, ws)
                codes = torch.stack(ids, -1).reshape(clips, t_frames, -1)
                return grvq_oracle.grvq_embed(codes, ws)
    else:
        cb = torch.from_numpy(synth.rvq_codebooks(w["n_q"], w["bins"], w["D"], 4321, "decay")).to(device)
        if ref is not None:
            from academicodec.quantization import ResidualVectorQuantizer as RefRVQ
            q = RefRVQ(dimension=w["D"], n_q=w["n_q"], bins=w["bins"], kmeans_init=False).to(device)
            for i, layer in enumerate(q.vq.layers):
                layer._codebook.embed.data.copy_(cb[i])
                layer._codebook.embed_avg.data.copy_(cb[i])
                layer._codebook.inited.data.fill_(1.0)
            q.train(train)

            def step():
                if train:
                    return q(x, w["frame_rate"])
                codes = q.encode(x, w["frame_rate"])
                return q.decode(codes)
        else:
            from oracle import rvq_oracle
            cbl = list(cb)
            states = rvq_oracle.make_states(cb) if train else None

            def step():
                if train:
                    return rvq_oracle.rvq_forward(x, states, None, training=True)
                codes = rvq_oracle.rvq_encode(x, cbl)
                return rvq_oracle.rvq_decode(codes, cbl)
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
    return clips * t_frames / dt, cores, dt, kind


def metric_of(w):
    return "rvq_ema_train_frames_per_sec" if w["kind"] == "rvq_train" else METRIC


def run_reference(args):
    world, rank, _ = dist_env()
    if rank != 0:
        return 0
    w = workload(args.workload)
    clips = max(1, min(w["B"], args.ref_clips))
    # each step = `clips` clips of the workload through the reference's CPU path
    fps, cores, dt, kind = cpu_throughput(w, clips, reps=max(1, args.steps), warm=max(1, min(args.warmup, 2)),
                                          prefer_reference=not args.ref_port)
    what = ("the reference's own modules (baseline/_ref, unmodified)" if kind == "reference"
            else "oracle port (torch CPU ops = the reference's ATen calls)")
    sample = f"{clips} of {w['B']} clips [{w['D']}x{w['T']}] per step, {what}, {cores} threads"
    line = {
        "impl": "reference", "metric": metric_of(w), "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "audio_sec_per_sec": fps / w["frame_rate"],
        "config": {"workload": w["name"], "D": w["D"], "n_q": w["n_q"], "bins": w["bins"],
                   "clips_per_step": clips, "frames_per_clip": w["T"]},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------ GPU arm
def _events(n):
    return [torch.cuda.Event(enable_timing=True) for _ in range(n)]


class Resident:
    """One workload resident on one GPU: inputs, codebooks, derived tables and the step functions."""

    def __init__(self, w, dev, rank, kernel, pin=True):
        from academicodec_b200 import ops
        self.ops, self.w, self.dev, self.kernel = ops, w, dev, kernel
        b, d, t, s, k = w["B"], w["D"], w["T"], w["n_q"], w["bins"]
        g = w.get("G", 1)
        self.b, self.d, self.t, self.s, self.k, self.g = b, d, t, s, k, g
        self.grvq = w["kind"] == "grvq"
        self.train = w["kind"] == "rvq_train"
        self.n_frames = b * t
        self.flags = (ops.ACQ_STE | ops.ACQ_LOSS_RAW) if self.grvq else (ops.ACQ_STE if self.train else 0)
        x_host = torch.from_numpy(synth.latents(b, d, t, 1234 + rank))
        self.x_host = x_host.pin_memory() if pin else x_host
        if self.grvq:
            ws = synth.grvq_codebooks(g, k, 777, "randn")
            self.cbs = [torch.from_numpy(a).to(dev) for st in ws for a in st]
        else:
            self.cbs = [c.contiguous() for c in torch.from_numpy(synth.rvq_codebooks(s, k, d, 4321, "decay")).to(dev)]
        self.hn = ops.codebook_half_norms(self.cbs)
        self.pack = None
        if kernel != 1 and ops.tc_supported(k, d, g):
            self.pack = ops.tc_pack_codebooks(self.cbs)      # tcgen05 operand images (built once per codebook)
        self.x_dev = self.x_host.to(dev, non_blocking=True)
        self.codes_dev = torch.empty((s * g, self.n_frames), dtype=torch.int64, device=dev)
        self.out_dev = torch.empty((b, d, t), dtype=torch.float32, device=dev)
        self.module = None
        torch.cuda.synchronize()

    # -- the two kernels of the encode+decode step, preallocated outputs (ops layer)
    def search(self):
        self.ops.rvq_search(self.x_dev, self.cbs, self.s, self.g, half_norms=self.hn, flags=self.flags,
                            impl=self.kernel, tc_pack=self.pack, codes_out=self.codes_dev)

    def decode(self):
        self.ops.vq_decode(self.codes_dev, self.n_frames, 1, self.cbs, self.s, self.g, self.b, self.t,
                           check=False, out=self.out_dev)

    # -- the drop-in module (reference class names and call signatures)
    def build_module(self):
        if self.module is not None:
            return self.module
        w = self.w
        if self.grvq:
            import types
            from academicodec_b200.grvq import Quantizer
            h = types.SimpleNamespace(n_code_groups=self.g, n_codes=self.k, codebook_loss_lambda=1.0,
                                      commitment_loss_lambda=0.25)
            q = Quantizer(h).to(self.dev)
            with torch.no_grad():
                for gi in range(self.g):
                    q.quantizer_modules[gi].embedding.weight.copy_(self.cbs[gi])
                    q.quantizer_modules2[gi].embedding.weight.copy_(self.cbs[self.g + gi])
            q.invalidate_caches()
        else:
            from academicodec_b200.quantization import ResidualVectorQuantizer
            q = ResidualVectorQuantizer(dimension=self.d, n_q=self.s, bins=self.k, kmeans_init=False).to(self.dev)
            with torch.no_grad():
                for i, layer in enumerate(q.vq.layers):
                    layer._codebook.embed.copy_(self.cbs[i])
                    layer._codebook.embed_avg.copy_(self.cbs[i])
                    layer._codebook.invalidate_caches()
            q.train(self.train)
        self.module = q
        return q

    def step(self):
        if self.train:
            return self.module(self.x_dev, self.w["frame_rate"])
        self.search()
        self.decode()


def time_phases(res, steps):
    """K steps on the current stream; -> (total ms, search ms per step, decode ms per step, host ms per step)."""
    ev = [_events(3) for _ in range(steps)]
    start, stop = _events(2)
    torch.cuda.synchronize()
    start.record()
    h0 = time.perf_counter()
    for i in range(steps):
        if res.train:
            ev[i][0].record()
            res.step()
            ev[i][1].record()
            ev[i][2].record()
        else:
            ev[i][0].record()
            res.search()
            ev[i][1].record()
            res.decode()
            ev[i][2].record()
    host_ms = (time.perf_counter() - h0) * 1e3 / steps
    stop.record()
    torch.cuda.synchronize()
    total = start.elapsed_time(stop)
    a = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
    b = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
    return total, a, b, host_ms


def time_fn(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = _events(2)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def effective_variant(w, cfg):
    """(variant, cluster) the library's automatic choice resolves to for this workload (c_api.cu run_tc)."""
    variant, cluster, split = cfg
    dg = w["D"] // w.get("G", 1)
    tiles = (w["B"] * w["T"] + 127) // 128
    n_pass = w["bins"] // 256
    small = bool(split) and ((n_pass % 4 == 0 and tiles * 4 <= 148) or (n_pass % 2 == 0 and tiles * 2 <= 148))
    v = variant or (1 if dg >= 512 and not small else 3)
    c = cluster or (2 if v == 1 and not small else 1)
    return (3 if small else v), c


def rooflines(w, n_frames, enc_ms, dec_ms, peaks, sustained, pack, kernel, variant, traffic):
    flops_enc, bytes_enc, bytes_dec = algorithmic(w)
    achieved_tf = flops_enc * n_frames / (enc_ms * 1e-3) / 1e12
    peak_tf = peaks["bf16_sustained"] if sustained else peaks["bf16_burst"]
    three = pack is not None and kernel != 1 and variant == 3
    executed = (3.0 if three else 1.0) * achieved_tf

    def dram(kernel_name):
        t = traffic.get(kernel_name)
        return (t["dram_read"] + t["dram_write"]) if t else None

    roof = {"bound": "tensor", "kernel": "rvq_search", "achieved": achieved_tf, "peak": peak_tf,
            "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": dram("rvq_search"),
            "traffic_note": "DRAM bytes per launch from the committed ncu capture (profiles/traffic.json); "
                            "algorithmic bytes per launch = %d" % int(bytes_enc * n_frames),
            "peak_source": f"{peaks['source']} bf16 dense ({'sustained' if sustained else 'burst'})",
            "ms_per_launch": enc_ms,
            # MMAs issued per algorithmic product: 3 for the hi/lo-split kernel (fp32-class scores), 1 for the
            # single-product filter + exact re-score kernel
            "executed_tflops": executed, "executed_frac": executed / peak_tf}
    if dec_ms and dec_ms > 0:
        dec_gbs = bytes_dec * n_frames / (dec_ms * 1e-3) / 1e9
        roof["decode"] = {"bound": "hbm", "kernel": "vq_decode", "achieved": dec_gbs, "peak": peaks["hbm_gbs"],
                          "unit": "GB/s", "frac": dec_gbs / peaks["hbm_gbs"], "traffic": dram("vq_decode"),
                          "ms_per_launch": dec_ms}
        sg = w["n_q"] * w.get("G", 1)
        if sg > 1:
            # multi-stage decode gathers S codeword rows per output row out of L2 (the tables do not fit shared
            # memory: S x K x D x 4 B), so its real bound is the L2 -> SM path, not HBM: 4*D*S bytes per frame
            # against the chip's L2 throughput cap (B300_MICROARCH.md: ~6300 B/clk over all slices)
            gather = 4.0 * w["D"] * w["n_q"] * n_frames / (dec_ms * 1e-3) / 1e9
            l2_peak = 6300.0 * 1.965
            roof["decode"]["l2_gather"] = {"achieved": gather, "peak": l2_peak, "unit": "GB/s",
                                           "frac": gather / l2_peak,
                                           "note": "codeword bytes gathered from L2 per second vs 6300 B/clk x 1.965 GHz"}
    return roof


def run_ours(args):
    world, rank, local = dist_env()
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; academicodec_b200 has no CPU path"}))
        return 2
    import torch.distributed as dist
    from academicodec_b200 import _lib, ops
    lib = _lib.load()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = None
    if world > 1:
        # pin this rank to the CPUs next to its GPU before any pinned host buffer is allocated (first touch
        # decides the NUMA node of the staging memory the e2e leg copies from / to)
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
            affinity = "nvml ideal CPUs: %d of them" % len(os.sched_getaffinity(0))
        except Exception as exc:
            affinity = "unchanged (%s)" % type(exc).__name__
        dist.init_process_group("nccl", device_id=dev)
    tc_cfg = _lib.tc_config_defaults()

    w = workload(args.workload)
    variant, cluster = effective_variant(w, tc_cfg)
    res = Resident(w, dev, rank, args.kernel)
    b, d, t, s, k, g = res.b, res.d, res.t, res.s, res.k, res.g
    n_frames = res.n_frames
    train = res.train
    if train:
        res.build_module()

    for _ in range(max(3, args.warmup)):
        res.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    # ---- timed region: K steps, CUDA events on the launching stream ----------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    if world > 1:
        # (rank 0 slept while it started the clock sampler: without this barrier the other ranks' timed regions
        #  would include that skew wherever a step has a cross-rank dependency, i.e. the cfg5 exchange)
        dist.barrier()
        torch.cuda.synchronize()
    total_ms, enc_ms, dec_ms, host_ms = time_phases(res, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    if train:
        # the search inside the training forward, timed by itself (same launch: STE residuals, all stages)
        enc_ms = time_fn(res.search, max(3, min(args.steps, 10)))
        dec_ms = 0.0
    t_ms = torch.tensor([total_ms, enc_ms, dec_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms, enc_ms, dec_ms = t_ms.tolist()
    ms_per_step = total_ms / args.steps
    value = world * n_frames / (ms_per_step * 1e-3)

    # ---- the collective of the path (cfg5): all-reduce of the flat EMA statistics buffer ------------------
    collective = None
    if train:
        nbytes = s * k * (d + 1) * 4
        if world > 1:
            from academicodec_b200.quantization.distrib import PeerExchange

            def timed(fn):
                ms = time_fn(fn, 20, warm=5)
                tt = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms = float(tt.item())
                return {"us": ms * 1e3, "bus_gbs": 2.0 * (world - 1) / world * nbytes / (ms * 1e-3) / 1e9}

            buf = torch.zeros(nbytes // 4, dtype=torch.float32, device=dev)
            nccl = timed(lambda: dist.all_reduce(buf))
            # the path the module takes: this package's kernel over NVLink peer memory (NVLS multimem when the
            # platform has a multicast mapping), bracketed by the symmetric-memory barriers
            peer = None
            exch = PeerExchange.create(nbytes // 4, dev)
            if exch is not None:
                exch.buf.zero_()
                peer = timed(exch.all_reduce_)
                peer["mode"] = exch.mode
            used = peer or nccl
            collective = {"op": "all_reduce(SUM) of [n_q, K, D+1] fp32 EMA statistics", "bytes": nbytes,
                          "us": used["us"], "bus_gbs": used["bus_gbs"], "nvlink_peak_gbs_per_dir": 900.0,
                          "frac_of_nvlink": used["bus_gbs"] / 900.0, "share_of_step": used["us"] * 1e-3 / ms_per_step,
                          "path": ("acq_peer_allreduce (%s)" % peer["mode"]) if peer else "nccl all_reduce",
                          "peer_memory": peer, "nccl": nccl, "ranks": world}
        else:
            collective = {"op": "all_reduce(SUM) of [n_q, K, D+1] fp32 EMA statistics", "bytes": nbytes,
                          "us": 0.0, "ranks": 1, "note": "single rank: no collective is issued"}

    # ---- sustained: the same resident step back to back for >= 2 s (power-capped steady state) -------------
    sustained = None
    if rank == 0 and args.sustained_s > 0:
        n_s = max(args.steps, int(args.sustained_s * 1e3 / max(ms_per_step, 1e-3)) + 1)
        if world > 1:
            n_s = min(n_s, 200)      # (other ranks idle meanwhile; keep it short)
        if world == 1:
            smp = ClockSampler(local)
            smp.start()
            s_ms = time_fn(res.step, n_s, warm=0)
            sclk = smp.stop()
            sustained = {"ms_per_step": s_ms, "value": n_frames / (s_ms * 1e-3), "unit": UNIT, "steps": n_s,
                         "seconds": s_ms * n_s / 1e3, "clocks": sclk}

    # ---- effective SM clock inside the search kernel (outside the timed region) -----------------
    # nvidia-smi's 100 ms samples cannot resolve a 20 ms timed region; the tensor-core kernels count their
    # own cycles (ACQ_TC_DBG bit 512: clock64 in the MMA-issuing warp), which with the event time of
    # the same launches gives the clock they really ran at.
    sm_mhz_in_kernel = None
    if rank == 0 and res.pack is not None:
        wsp = ops.tc_workspace(d, dev)
        base = int(lib.acq_tc_workspace_bytes(d)) - 256 + 64
        prev = os.environ.get("ACQ_TC_DBG")
        os.environ["ACQ_TC_DBG"] = "512"
        try:
            wsp[base:base + 72].zero_()
            nk = 5
            k_ms = time_fn(res.search, nk, warm=0)
            ctas = min(148, (n_frames + 127) // 128)
            cyc = wsp[base:base + 72].view(torch.int64)[7].item() / (nk * ctas)
            if cyc > 0:
                sm_mhz_in_kernel = cyc / k_ms / 1e3
        finally:
            if prev is None:
                os.environ.pop("ACQ_TC_DBG", None)
            else:
                os.environ["ACQ_TC_DBG"] = prev

    # ---- end to end -------------------------------------------------------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    if train:
        # the public call of the training path is the module forward: pinned host latents -> H2D -> forward
        # (search, replay, all-reduce, EMA apply) -> D2H of the commitment loss
        x_in = torch.empty_like(res.x_dev)
        q = res.module

        def step_e2e():
            x_in.copy_(res.x_host, non_blocking=True)
            out = q(x_in, w["frame_rate"])
            return float(out.penalty.detach().cpu()) if hasattr(out, "penalty") else float(out[-1].detach().cpu())

        for _ in range(2):
            step_e2e()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        torch.cuda.synchronize()
        e2e_dt = (time.perf_counter() - t0) / e2e_steps
        h2d, d2h = res.x_host.numel() * 4, 4
        e2e_api = "ResidualVectorQuantizer.forward (train mode): H2D latents, forward + EMA update, D2H loss"
        e2e_ok, e2e_launches = True, None
    else:
        pipe = ops.HostPipeline(local, args.chunk_mb << 20)
        codes_host = torch.empty((s * g, n_frames), dtype=torch.int64).pin_memory()
        out_host = torch.empty((b, d, t), dtype=torch.float32).pin_memory()

        def step_e2e():
            # one public call: host latents -> codes + reconstructed latents on the host
            pipe.rvq_codec(res.x_host, res.cbs, s, g, res.hn, flags=res.flags, impl=args.kernel,
                           codes_out=codes_host, out=out_host, tc_pack=res.pack)
            return pipe.last_launches

        e2e_launches = 0
        for _ in range(2):
            e2e_launches = step_e2e()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        torch.cuda.synchronize()
        e2e_dt = (time.perf_counter() - t0) / e2e_steps
        e2e_ok = bool(torch.equal(codes_host, res.codes_dev.cpu()) and torch.equal(out_host, res.out_dev.cpu()))
        h2d = res.x_host.numel() * 4
        d2h = codes_host.numel() * 8 + out_host.numel() * 4
        e2e_api = "acq_rvq_codec_host (encode -> decode, codes stay on the device in between)"
        # The ceiling of this leg is the host link: the same bytes as plain copies, both directions at once, on
        # every rank at the same time (the box's aggregate H2D / D2H capacity is shared by its GPUs).
        s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        x_sink = torch.empty_like(res.x_dev)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        c0 = time.perf_counter()
        n_probe = 3
        for _ in range(n_probe):
            with torch.cuda.stream(s_up):
                x_sink.copy_(res.x_host, non_blocking=True)
            with torch.cuda.stream(s_dn):
                out_host.copy_(res.out_dev, non_blocking=True)
                codes_host.copy_(res.codes_dev, non_blocking=True)
        torch.cuda.synchronize()
        probe_dt = (time.perf_counter() - c0) / n_probe
        del x_sink
    e2e_s = torch.tensor([e2e_dt, probe_dt if not train else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * n_frames / float(e2e_s[0].item())
    e2e_ceiling = None
    if not train and float(e2e_s[1].item()) > 0:
        pdt = float(e2e_s[1].item())
        e2e_ceiling = {"value": world * n_frames / pdt, "unit": UNIT,
                       "h2d_gbs_aggregate": world * h2d / pdt / 1e9, "d2h_gbs_aggregate": world * d2h / pdt / 1e9,
                       "frac": e2e_value / (world * n_frames / pdt),
                       "note": "the step's H2D and D2H bytes as plain pinned-memory copies, both directions at once, "
                               "on all %d ranks simultaneously (max over ranks): the host-link bound of this leg" % world}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- the drop-in surface itself: module calls, per call, CUDA events --------------------------------
    module = None
    if not train and not args.no_module:
        q = res.build_module()
        with torch.no_grad():
            if res.grvq:
                fwd = lambda: q(res.x_dev)                                   # noqa: E731
                codes_m = torch.stack(q(res.x_dev)[2], -1).reshape(b, t, -1)
                f_ms = time_fn(fwd, 5)
                e_ms = time_fn(lambda: q.embed(codes_m), 5)
                module = {"Quantizer.forward_ms": f_ms, "Quantizer.embed_ms": e_ms,
                          "frames_per_s": n_frames / ((f_ms + e_ms) * 1e-3)}
            else:
                codes_m = q.encode(res.x_dev, w["frame_rate"])
                f_ms = time_fn(lambda: q.encode(res.x_dev, w["frame_rate"]), 5)
                e_ms = time_fn(lambda: q.decode(codes_m), 5)
                module = {"ResidualVectorQuantizer.encode_ms": f_ms, "ResidualVectorQuantizer.decode_ms": e_ms,
                          "frames_per_s": n_frames / ((f_ms + e_ms) * 1e-3),
                          "note": "outputs allocated per call by torch's caching allocator; the code-range check "
                                  "of decode() is deferred (no host sync)"}
            del codes_m

    # ---- roofline of the dominant kernel (the search) ------------------------------------------
    peaks = load_peaks()
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and res.pack is not None:
        traffic = json.load(open(tpath)).get(w["name"], {})
    roof = rooflines(w, n_frames, enc_ms, dec_ms, peaks, total_ms > 1000, res.pack, args.kernel, variant, traffic)
    if sustained is not None and not train:
        roof["sustained_note"] = "see `sustained`: the same step for >= 2 s at the board's power cap"

    # ---- the other BASELINE configs (N = 1, default workload only) -----------------------------------------
    secondary = None
    if world == 1 and not args.no_secondary and args.workload == DEFAULT_WORKLOAD:
        secondary = {}
        del res.x_dev, res.out_dev, res.codes_dev
        torch.cuda.empty_cache()
        for name in SECONDARY:
            try:
                w2 = workload(name)
                r2 = Resident(w2, dev, 0, args.kernel, pin=False)
                if r2.train:
                    r2.build_module()
                for _ in range(3):
                    r2.step()
                n2 = 10
                tot2, a2, b2, _ = time_phases(r2, n2)
                if r2.train:
                    a2, b2 = time_fn(r2.search, 5), 0.0
                v2, c2 = effective_variant(w2, tc_cfg)
                rf2 = rooflines(w2, r2.n_frames, a2, b2, peaks, False, r2.pack, args.kernel, v2, {})
                secondary[name] = {
                    "kind": w2["kind"], "D": w2["D"], "n_q": w2["n_q"], "groups": w2.get("G", 1),
                    "clips": w2["B"], "frames_per_clip": w2["T"], "ms_per_step": tot2 / n2,
                    "value": r2.n_frames / (tot2 / n2 * 1e-3), "unit": UNIT, "search_ms": a2, "decode_ms": b2,
                    "search_tflops": rf2["achieved"], "search_frac": rf2["frac"],
                    "search_kernel": "tcgen05 %d-product, cluster %d" % (v2, c2) if r2.pack is not None else "simt",
                    "decode_gbs": rf2.get("decode", {}).get("achieved"),
                    "decode_frac": rf2.get("decode", {}).get("frac")}
                del r2
                torch.cuda.empty_cache()
            except Exception as exc:          # a secondary line must never take the headline down
                secondary[name] = {"error": repr(exc)[:200]}

    eager = None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        if not train:
            # secondary baseline: the reference's op sequence as PyTorch eager kernels on this GPU
            try:
                efps, _, edt, ekind = cpu_throughput(w, b, reps=3, warm=2, device=f"cuda:{local}")
                eager = {"value": efps, "unit": UNIT,
                         "kind": f"{ekind} (torch eager CUDA ops, fp32, TF32 off)",
                         "sample": f"all {b} clips, 3 timed reps after 2 warm-ups, {edt * 1e3:.1f} ms per rep"}
            except torch.cuda.OutOfMemoryError:
                eager = {"value": None, "unit": UNIT, "kind": "torch eager CUDA ops", "sample": "out of memory"}
        clips = max(1, min(b, args.ref_clips))
        fps, cores, dt, kind = cpu_throughput(w, clips, reps=2, warm=1, prefer_reference=not args.ref_port)
        what = ("the reference's own modules (baseline/_ref, unmodified)" if kind == "reference"
                else "oracle port (torch CPU ops = the reference's ATen calls)")
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{clips} of {b} clips [{d}x{t}], 2 timed reps after 1 warm-up, {what}, "
                         f"{dt * 1e3:.0f} ms per rep"}

    launches_per_step = (3 if world == 1 else 3) if train else 2     # search + replay + EMA apply | search + decode
    line = {
        "metric": metric_of(w), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "audio_sec_per_sec": value / w["frame_rate"],
        "config": {"workload": w["name"], "kind": w["kind"], "D": d, "n_q": s, "groups": g, "bins": k,
                   "clips_per_gpu": b, "frames_per_clip": t, "frame_rate": w["frame_rate"],
                   "l2": "inputs+outputs per step (%.0f MB) %s the 126 MB L2" %
                         ((h2d + d2h) / 1e6, "exceed" if (h2d + d2h) > 126e6 else "fit in"),
                   "kernel": ("tcgen05 (%s, cluster %d)" % ("1 product + filter + exact re-score" if variant == 1
                                                             else "3 products hi/lo", cluster))
                             if res.pack is not None else "simt"},
        "encode_ms": enc_ms, "decode_ms": dec_ms, "host_launch_ms_per_step": host_ms,
        "roofline": roof, "cpu_baseline": cpu, "eager_gpu_baseline": eager,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "matches_resident": e2e_ok, "chunk_mb": args.chunk_mb, "api": e2e_api,
                "ceiling": e2e_ceiling, "cpu_affinity": affinity},
        "gpu_launches": launches_per_step * args.steps + (e2e_launches or launches_per_step) * e2e_steps,
        "clocks": clocks,
    }
    if module is not None:
        line["module"] = module
    if sustained is not None:
        line["sustained"] = sustained
    if collective is not None:
        line["collective"] = collective
    if secondary is not None:
        line["secondary"] = secondary
    if (h2d + d2h) <= 126e6:
        line["config"]["l2"] += " (the reference's own batch size; the throughput variants of this shape are in `secondary`)"
    if clocks is not None and sm_mhz_in_kernel is not None:
        clocks["sm_mhz_in_search_kernel"] = round(sm_mhz_in_kernel, 1)
        clocks["note"] = ("in-kernel clock = the search kernel's own cycle count / its event time; when it is below "
                          "sm_max_mhz the tensor-core kernel is drawing the board's power budget "
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
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(synth.WORKLOADS))
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 SIMT, 2 tensor-core")
    ap.add_argument("--chunk-mb", type=int, default=96,
                    help="staging bytes per in-flight chunk of the e2e pipeline; >= one clip keeps the PCIe copies 1-D")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-clips", type=int, default=2,
                    help="clips per step of the CPU arm / cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-port", action="store_true",
                    help="CPU arm: time the oracle port even when baseline/_ref (the reference's own modules) is installed")
    ap.add_argument("--no-secondary", action="store_true", help="skip the `secondary` block (other BASELINE configs)")
    ap.add_argument("--no-module", action="store_true", help="skip the module-level timing block")
    ap.add_argument("--sustained-s", type=float, default=2.0, help="seconds of the back-to-back sustained leg (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Headline benchmark: info Gb/s (frames/s) of batched flooding min-sum decoding at 10 iterations.

Workload (BASELINE.json configs[1] on the largest named code shape): Neural2DMinSumDecoder,
weight_sharing_type=2, 10 iterations, 65 536 frames per GPU, synthetic (16200,7200)-shaped code
(E = 48 599), AWGN LLRs at 2 dB in the reference's own sign convention (ldpc_decoder.py:289), under which
no frame satisfies the parity checks, so every frame executes exactly 10 full iterations with the
per-iteration posterior / hard decision / syndrome / early-stop test of the reference still running.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 our arm
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] CPU arm (oracle port, all host threads)

A step = one decode of the whole batch.  `value` times ldpc_decode_device with the LLRs resident in HBM
(CUDA events on the launching stream); `e2e` times ldpc_decode_host (pinned host LLRs in, hard decisions
+ iteration counts + success flags out, copies inside the timed region).  One JSON line on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "info_gbps_at_10_iters"
UNIT = "info Gb/s"
T_ITERS = 10
SNR_DB = 2.0
# dram__bytes_read.sum + dram__bytes_write.sum per FRAME of one launch, from the committed `ncu --set full`
# captures (profiles/r01b_ncu_full_*_8192frames_raw.csv: bytes of one launch / 8192 frames); scaled by the
# frames of the benchmarked launch.  Only the captured (decoder, code, kernel) pairs have an entry.
NCU_TRAFFIC_BYTES_PER_FRAME = {
    ("n2d2", "dvbs2", "vn_kernel"): (2.123882e9 + 1.558320e9) / 8192,
    ("n2d2", "dvbs2", "cn_kernel"): (1.592799e9 + 1.542210e9) / 8192,
    ("rcq", "dvbs2", "vn_kernel"): (0.929480e9 + 1.551163e9) / 8192,
    ("rcq", "dvbs2", "cn_kernel"): (1.592856e9 + 0.371589e9) / 8192,
    ("n2d2", "qc", "vn_kernel"): (1.552332e9 + 1.204139e9) / 8192,
    ("n2d2", "qc", "cn_kernel"): (1.242029e9 + 1.192600e9) / 8192,       # cn_wide_kernel (row ring)
    ("wrcq1", "qc", "vn_kernel"): (0.621112e9 + 1.193412e9) / 8192,
    ("wrcq1", "qc", "cn_kernel"): (1.241724e9 + 0.280145e9) / 8192,      # cn_wide_kernel (row ring)
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=65536, help="frames per GPU per step")
    ap.add_argument("--code", default="dvbs2", choices=["dvbs2", "qc", "dv12"])
    ap.add_argument("--decoder", default="n2d2", choices=["n2d2", "n2d1", "nnms", "oms2", "rcq", "wrcq1", "basic"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-mc", action="store_true", help="skip the early-stop Monte-Carlo leg (T=50, frames converge)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def make_code(L, name):
    if name == "dv12":   # not a BASELINE shape: exercises the variable-node path for degrees above 8
        return L.codes.ira_code({12: 1620, 3: 4860}, {5: 4861, 6: 4859}, max_iterations=T_ITERS)
    return L.codes.dvbs2_shaped(max_iterations=T_ITERS) if name == "dvbs2" else L.codes.qc_shaped(max_iterations=T_ITERS)


def workload_name(args):
    shape = {"dvbs2": "(16200,7200)-shaped E=48599", "qc": "(9472,8192)-shaped QC E=37888",
             "dv12": "(16200,6480) IRA dv 12/3/2 E=53459"}[args.code]
    dec = {"n2d2": "Neural2DMinSumDecoder type 2", "n2d1": "Neural2DMinSumDecoder type 1", "nnms": "NeuralMinSumDecoder (per-edge weights)", "oms2": "Neural2DOffsetMinSumDecoder type 2", "rcq": "RCQMinSumDecoder bc=3", "wrcq1": "WeightedRCQ type 1 bc=3",
           "basic": "BasicMinSumDecoder f64 factor 0.7"}[args.decoder]
    return f"{dec}, {T_ITERS} iters, {shape}, AWGN {SNR_DB} dB reference sign convention"


def det_weights(T):
    t = np.arange(T, dtype=np.float64)
    return (0.75 + t / 64).astype(np.float32), (1 - t / 32).astype(np.float32)


def build_decoder(L, code, kind):
    import torch
    qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
    b, a = det_weights(T_ITERS)
    if kind == "n2d2":
        dec = L.Neural2DMinSumDecoder(code, weight_sharing_type=2, max_iterations=T_ITERS)
        with torch.no_grad():
            dec._beta_table.copy_(torch.from_numpy(b)[:, None].expand_as(dec._beta_table))
            dec._alpha_table.copy_(torch.from_numpy(a)[:, None].expand_as(dec._alpha_table))
    elif kind == "n2d1":
        dec = L.Neural2DMinSumDecoder(code, weight_sharing_type=1, max_iterations=T_ITERS)
        with torch.no_grad():
            dec._beta_table.copy_(torch.from_numpy(b)[:, None].expand_as(dec._beta_table))
    elif kind == "oms2":
        dec = L.Neural2DOffsetMinSumDecoder(code, weight_sharing_type=2, max_iterations=T_ITERS)
        with torch.no_grad():
            dec._beta_table.fill_(0.15)
            dec._alpha_table.fill_(0.02)
    elif kind == "nnms":
        dec = L.NeuralMinSumDecoder(code, max_iterations=T_ITERS)
        with torch.no_grad():
            dec._beta_table.copy_(torch.from_numpy(b)[:, None].expand_as(dec._beta_table))
    elif kind == "rcq":
        dec = L.RCQMinSumDecoder(code, bc=3, bv=8, quantizer_params=qp, max_iterations=T_ITERS)
    elif kind == "wrcq1":
        dec = L.WeightedRCQDecoder(code, bc=3, bv=8, quantizer_params=qp, weight_sharing_type=1, max_iterations=T_ITERS)
        with torch.no_grad():
            dec._beta_table.copy_(torch.from_numpy(b)[:, None].expand_as(dec._beta_table))
    else:
        dec = L.BasicMinSumDecoder(code, factor=0.7)
    return dec


def algorithmic_bytes(code, kind):
    """bytes per frame-iteration at storage width, and per kernel launch per frame."""
    g = code.graph
    E, n = g.E, g.n
    if kind in ("rcq", "wrcq1"):
        return dict(frame_iter=10 * E + 4 * n, cn=4 * E + 1 * E, vn=1 * E + 4 * n + 4 * E, vn_final=1 * E + 4 * n)
    w = 8 if kind == "basic" else 4
    return dict(frame_iter=4 * w * E + w * n, cn=2 * w * E, vn=2 * w * E + w * n, vn_final=w * E + w * n)


def oracle_setup(L, code, kind):
    from oracle.restatement import MODE_NMS, MODE_RCQ, MODE_WRCQ, SparseGraph, quantizer_schedule, quantizer_thresholds
    g = code.graph
    og = SparseGraph.from_coo(g.n, g.m, g.edge_check, g.check_var)
    b, a = det_weights(T_ITERS)
    kw = dict(T=T_ITERS, want_posterior=False)
    qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
    thr = np.array([quantizer_thresholds(3, C, gm) for C, gm in qp]).astype(np.float32)
    if kind == "n2d2":
        kw.update(mode=MODE_NMS, beta=np.tile(b[:, None], (1, g.E)), alpha=np.tile(a[:, None], (1, g.n)))
    elif kind in ("n2d1", "nnms"):
        kw.update(mode=MODE_NMS, beta=np.tile(b[:, None], (1, g.E)))
    elif kind == "oms2":
        from oracle.restatement import MODE_OFFSET
        kw.update(mode=MODE_OFFSET, beta=np.full((T_ITERS, g.E), np.float32(0.15)), alpha=np.full((T_ITERS, g.n), np.float32(0.02)))
    elif kind == "rcq":
        kw.update(mode=MODE_RCQ, bc=3, thresholds=thr, quantizer_of_iter=quantizer_schedule(T_ITERS, 3))
    elif kind == "wrcq1":
        kw.update(mode=MODE_WRCQ, bc=3, thresholds=thr, quantizer_of_iter=quantizer_schedule(T_ITERS, 3),
                  beta=np.tile(b[:, None], (1, g.E)), alpha=np.ones((T_ITERS, g.n), np.float32))
    else:
        kw.update(mode=MODE_NMS, dtype=np.float64, beta=np.full((T_ITERS, g.E), 0.7))
    return og, kw


def host_llrs(code, frames, seed, dtype=np.float32):
    rng = np.random.default_rng(seed)
    s2 = 10 ** (-SNR_DB / 10)
    return (2 * (-1.0 + np.sqrt(s2) * rng.standard_normal((frames, code.n), dtype=np.float32)) / s2).astype(dtype)


def cpu_leg(L, code, kind, seconds, threads):
    """Time the oracle port on a bounded sample of the same workload.  Returns dict for `cpu_baseline`."""
    from oracle import capi as O
    og, kw = oracle_setup(L, code, kind)
    dtype = kw.get("dtype", np.float32)
    probe = host_llrs(code, max(threads, 8), 99, dtype)
    t0 = time.perf_counter()
    O.decode(og, probe, nthreads=threads, **kw)
    per_frame = (time.perf_counter() - t0) / probe.shape[0]
    frames = int(max(threads, min(200000, seconds / max(per_frame, 1e-9))))
    frames = max(threads, frames // threads * threads)
    llr = host_llrs(code, frames, 100, dtype)
    t0 = time.perf_counter()
    res = O.decode(og, llr, nthreads=threads, **kw)
    dt = time.perf_counter() - t0
    fps = frames / dt
    return dict(value=fps * code.k / 1e9, unit=UNIT, cores=threads, kind="port", frames_per_s=fps,
                sample=f"{frames} frames of the same workload ({dt:.1f} s), oracle/minsum_oracle.c with {threads} OpenMP threads, "
                       f"avg iterations {float(res.iterations.mean()):.2f}"), dt, frames


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads to the CPUs next to its GPU (NVML's ideal affinity) BEFORE any pinned host
    buffer is allocated, so that the end-to-end leg's staging memory is local to the GPU's PCIe root.  With
    several ranks per box the host-to-device copies otherwise share one socket's memory controllers."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ldpc_b200 as L
    from oracle import capi as O
    code = make_code(L, args.code)
    threads = os.cpu_count() or 1
    og, kw = oracle_setup(L, code, args.decoder)
    dtype = kw.get("dtype", np.float32)
    # step = a bounded sample sized so the whole run stays within a few minutes
    budget = 150.0 / max(1, args.steps + args.warmup)
    probe = host_llrs(code, max(threads, 8), 1, dtype)
    t0 = time.perf_counter()
    O.decode(og, probe, nthreads=threads, **kw)
    per_frame = (time.perf_counter() - t0) / probe.shape[0]
    frames = max(threads, int(min(budget, 20.0) / per_frame) // threads * threads)
    llr = host_llrs(code, frames, 2, dtype)
    for _ in range(args.warmup):
        O.decode(og, llr, nthreads=threads, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = O.decode(og, llr, nthreads=threads, **kw)
    dt = time.perf_counter() - t0
    fps = frames * args.steps / dt
    val = fps * code.k / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if args.decoder == "basic" else "f32", "data": "synthetic",
        "frames_per_s": fps,
        "config": {"workload": workload_name(args), "frames_per_step": frames,
                   "note": "reference algorithm timed as its C port (the reference is pure Python and cannot travel to the GPU box)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{frames} frames per step x {args.steps} steps, avg iterations {float(res.iterations.mean()):.2f}"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def mc_leg(L, code, kind, B, local_rank, rank, world, barrier, dev, T=50, snrs=(2.0, 3.0), rounds=3):
    import torch
    import torch.distributed as dist
    t = np.arange(T, dtype=np.float64)
    beta = np.minimum(0.75 + t / 64, 1.0).astype(np.float32)
    qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
    out = []
    for compact in (0, 1):
        os.environ["LDPC_COMPACT"] = str(compact)      # read when the decoder handle is created
        if kind == "n2d2":
            dec = L.Neural2DMinSumDecoder(code, weight_sharing_type=2, max_iterations=T)
            with torch.no_grad():
                dec._beta_table.copy_(torch.from_numpy(beta)[:, None].expand_as(dec._beta_table))
                dec._alpha_table.fill_(1.0)
        elif kind == "rcq":
            dec = L.RCQMinSumDecoder(code, bc=3, bv=8, quantizer_params=qp, max_iterations=T)
        else:
            dec = L.WeightedRCQDecoder(code, bc=3, bv=8, quantizer_params=qp, weight_sharing_type=1, max_iterations=T)
            with torch.no_grad():
                dec._beta_table.copy_(torch.from_numpy(beta)[:, None].expand_as(dec._beta_table))
        eng = dec._engine(local_rank)
        counters = torch.zeros(4, dtype=torch.int64, device=dev)
        for snr in snrs:
            eng.mc_round(snr, B, seed=7, frame0=rank * B, llr_sign=1, counters=counters)     # warm-up / allocation
            barrier()
            counters.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for r in range(rounds):
                eng.mc_round(snr, B, seed=11, frame0=(r * world + rank) * B, llr_sign=1, counters=counters)
            e1.record()
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                dist.all_reduce(counters)
            fe, be, it, nf = (int(v) for v in counters.tolist())
            fps = nf / (float(ms.item()) / 1e3)
            out.append({"snr_db": snr, "max_iterations": T, "compaction": bool(compact), "frames": nf,
                        "frames_per_s": fps, "info_gbps": fps * code.k / 1e9, "avg_iterations": it / max(nf, 1),
                        "fer": fe / max(nf, 1), "ber": be / max(nf * code.n, 1)})
        del dec, eng
    os.environ.pop("LDPC_COMPACT", None)
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import ldpc_b200 as L
    code = make_code(L, args.code)
    g = code.graph
    kind = args.decoder
    dec = build_decoder(L, code, kind)
    B = int(args.frames)
    f64 = kind == "basic"
    eng = dec._engine(local_rank)
    eng.reserve(B)

    # synthetic LLRs generated on the device (Philox), distinct frames per rank; resident before timing
    llr = L.awgn_llr(g.n, B, SNR_DB, seed=1234, frame0=rank * B, llr_sign=-1, device=local_rank)
    if f64:
        llr = llr.double()
    torch.cuda.synchronize()

    def step_device():
        return eng.decode_device(llr, want_posterior=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        bits, _, iters, succ = step_device()
    barrier()
    avg_iters = float(iters.float().mean().item())
    eng.profile_read(reset=True)
    eng.profile_mode(1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        bits, _, iters, succ = step_device()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    ms = e0.elapsed_time(e1)
    prof = eng.profile_read(reset=True)
    eng.profile_mode(0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    fps = world * B * args.steps / (ms_max / 1e3)
    value = fps * code.k / 1e9

    # ---- roofline of the dominant kernel (live CUDA-event durations from inside the timed region) ----
    ab = algorithmic_bytes(code, kind)
    Bp = prof["frames_padded"]
    peak, peak_src = peak_hbm()
    vn_per_step = prof["vn_launches"] / args.steps
    vn_bytes = ((vn_per_step - 1) * ab["vn"] + ab["vn_final"]) / vn_per_step * Bp
    cn_bytes = ab["cn"] * Bp
    vn_ms = prof["vn_ms"] / max(prof["vn_launches"], 1)
    cn_ms = prof["cn_ms"] / max(prof["cn_launches"], 1)
    vn_gbs = vn_bytes / (vn_ms * 1e-3) / 1e9
    cn_gbs = cn_bytes / (cn_ms * 1e-3) / 1e9
    dominant = "vn_kernel" if prof["vn_ms"] >= prof["cn_ms"] else "cn_kernel"
    ach = vn_gbs if dominant == "vn_kernel" else cn_gbs
    step_bytes = T_ITERS * ab["frame_iter"] * B
    roofline = {
        "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
        "traffic": (NCU_TRAFFIC_BYTES_PER_FRAME[(kind, args.code, dominant)] * Bp
                    if (kind, args.code, dominant) in NCU_TRAFFIC_BYTES_PER_FRAME else None),
        "traffic_source": "ncu --set full at 8192 frames per launch, scaled per frame (profiles/README.md, r01b captures)",
        "kernel": dominant, "peak_source": peak_src,
        "bytes_per_launch": vn_bytes if dominant == "vn_kernel" else cn_bytes,
        "avg_launch_ms": vn_ms if dominant == "vn_kernel" else cn_ms,
        "vn_kernel": {"gbs": vn_gbs, "frac": vn_gbs / peak, "ms_total": prof["vn_ms"], "launches": prof["vn_launches"]},
        "cn_kernel": {"gbs": cn_gbs, "frac": cn_gbs / peak, "ms_total": prof["cn_ms"], "launches": prof["cn_launches"]},
        "other_ms_total": prof["other_ms"],
        "whole_step": {"gbs": step_bytes / (ms_max / args.steps * 1e-3) / 1e9,
                       "frac": step_bytes / (ms_max / args.steps * 1e-3) / 1e9 / peak,
                       "bytes_per_frame_iter": ab["frame_iter"]},
    }

    # ---- end to end through the host-buffer C-ABI call (pinned host buffers, copies inside) ----
    e2e = None
    if not args.no_e2e:
        rdt = np.float64 if f64 else np.float32
        pin_llr = L.PinnedBuffer((B, g.n), rdt)
        pin_bits = L.PinnedBuffer((B, g.n), np.uint8)
        pin_it = L.PinnedBuffer((B,), np.int32)
        pin_su = L.PinnedBuffer((B,), np.uint8)
        chunk = 8192
        for s in range(0, B, chunk):   # fill the pinned input from the device-generated LLRs
            pin_llr.array[s:s + chunk] = llr[s:s + chunk].cpu().numpy()
        outs = dict(bits=pin_bits.array, iterations=pin_it.array, success=pin_su.array)
        for _ in range(max(1, min(args.warmup, 2))):
            eng.decode_host(pin_llr.array, out=outs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.decode_host(pin_llr.array, out=outs)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e_fps = world * B * args.steps / dt
        same = bool(np.array_equal(pin_bits.array[:256], bits[:256].cpu().numpy()))
        e2e = {"value": e2e_fps * code.k / 1e9, "unit": UNIT, "frames_per_s": e2e_fps,
               "h2d_bytes_per_step": int(B * g.n * np.dtype(rdt).itemsize),
               "d2h_bytes_per_step": int(B * g.n + B * 4 + B), "ms_per_step": 1e3 * dt / args.steps,
               "api": "ldpc_decode_host (pinned host LLR in; hard decisions + iterations + success out)",
               "matches_device_path": same}
        for p in (pin_llr, pin_bits, pin_it, pin_su):
            p.free()

    # ---- early-stop leg: BASELINE configs[4] shape -- T = 50, LLRs of the all-zero codeword in the
    # converging sign convention, AWGN generated on the device (ldpc_mc_round), frames stop at different
    # iterations; with and without frame compaction.  Reported next to the headline, not part of `value`.
    early = None
    if not args.no_mc and kind != "basic":
        early = mc_leg(L, code, kind, B, local_rank, rank, world, barrier, dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, _, _ = cpu_leg(L, code, kind, args.cpu_seconds, os.cpu_count() or 1)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if f64 else "f32", "data": "synthetic",
            "frames_per_s": fps, "edge_msgs_per_s": fps * T_ITERS * 2 * g.E, "avg_iterations": avg_iters,
            "config": {"workload": workload_name(args), "frames_per_gpu": B, "global_frames": world * B,
                       "n": g.n, "k": code.k, "E": g.E, "iterations": T_ITERS, "early_stop": True,
                       "l2": "inputs larger than L2 (message arrays %.1f GB per GPU)" % (2 * 4 * g.E * Bp / 1e9),
                       "parallelism": f"frames sharded over {world} GPU(s), no data-path collective",
                       "host_affinity": (f"{len(numa)} CPUs next to the GPU (NVML)" if numa else "unbound")},
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "early_stop_mc": early,
            "gpu_launches": int(prof["launches"]),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

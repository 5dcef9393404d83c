#!/usr/bin/env python
"""Headline benchmark: info Gb/s (frames/s) of batched flooding min-sum decoding at 10 iterations.

Workload of `value` (BASELINE.json configs[1] on the largest named code shape): Neural2DMinSumDecoder.forward,
weight_sharing_type=2, 10 iterations, 65 536 frames per GPU, synthetic (16200,7200)-shaped code (E = 48 599),
AWGN LLRs at 2 dB in the reference's own sign convention (ldpc_decoder.py:289), under which no frame satisfies
the parity checks, so every frame executes exactly 10 full iterations with the per-iteration posterior / hard
decision / syndrome / early-stop test of the reference still running.  forward() returns (decoded, posterior,
iterations) (neural_2d_decoder.py:206-225): the timed call delivers all three.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 our arm
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] CPU arm (oracle port, all host threads)

A step = one decode of the whole batch.  `value` times ldpc_decode_device with the LLRs resident in HBM (CUDA events
on the launching stream); `e2e` times ldpc_decode_host (pinned host LLRs in; hard decisions + posteriors + iteration
counts + success flags out; every copy inside the timed region).  `decode_only` / `e2e_decode_only` are the same two
calls without the posterior (the reference's decode() surface).  `configs` holds one short leg for each of the other
BASELINE.json configurations (C1 (7,4) Basic f64, C3 RCQ bc=3, C4 W-RCQ type 1 on the QC shape, C5 the Monte-Carlo
SNR sweep through LDPSimulator at T = 50); at N > 1 `mc_parity` says whether the all-reduced Monte-Carlo counters of a
sharded round equal those of the same global frame range decoded by rank 0 alone.  One JSON line on stdout.
"""
from __future__ import annotations

import argparse
import csv
import gc
import glob
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "implementation-of-neural-ldpc-decoders-with-degree-specific-weight-sharing-and-rcq-quantization_b200"

METRIC = "info_gbps_at_10_iters"
UNIT = "info Gb/s"
T_ITERS = 10
SNR_DB = 2.0
QP = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=65536, help="frames per GPU per step")
    ap.add_argument("--code", default="dvbs2", choices=["dvbs2", "qc", "dv12", "h74", "r504"])
    ap.add_argument("--decoder", default="n2d2", choices=["n2d2", "n2d1", "nnms", "oms2", "rcq", "wrcq1", "basic"])
    ap.add_argument("--decode-only", action="store_true", help="time decode() (no posterior) as `value` instead of forward()")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the legs for BASELINE configs 1, 3, 4, 5")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def host_only_package():
    """The package's host-side modules (code generators, graph construction) WITHOUT importing the package itself:
    its __init__ loads the CUDA library, which the CPU reference arm must not map."""
    name = "_ldpc_b200_hostonly"
    if name not in sys.modules:
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(ROOT, PKG)]
        sys.modules[name] = pkg
    ns = types.SimpleNamespace()
    ns.codes = importlib.import_module(name + ".codes")
    ns.create_test_ldpc_code = importlib.import_module(name + ".ldpc_decoder").create_test_ldpc_code
    return ns


def make_code(L, name, T=T_ITERS):
    if name == "dv12":   # not a BASELINE shape: exercises the variable-node path for degrees above 8
        return L.codes.ira_code({12: 1620, 3: 4860}, {5: 4861, 6: 4859}, max_iterations=T)
    if name == "h74":
        code = L.create_test_ldpc_code()
        code.max_iterations = T
        return code
    if name == "r504":
        return L.codes.regular_code(504, 3, 6, max_iterations=T)
    return L.codes.dvbs2_shaped(max_iterations=T) if name == "dvbs2" else L.codes.qc_shaped(max_iterations=T)


SHAPES = {"dvbs2": "(16200,7200)-shaped E=48599", "qc": "(9472,8192)-shaped QC E=37888",
          "dv12": "(16200,6480) IRA dv 12/3/2 E=53459", "h74": "(7,4) test code E=13", "r504": "(3,6)-regular n=504 E=1512"}
DECODERS = {"n2d2": "Neural2DMinSumDecoder type 2", "n2d1": "Neural2DMinSumDecoder type 1",
            "nnms": "NeuralMinSumDecoder (per-edge weights)", "oms2": "Neural2DOffsetMinSumDecoder type 2",
            "rcq": "RCQMinSumDecoder bc=3", "wrcq1": "WeightedRCQ type 1 bc=3", "basic": "BasicMinSumDecoder f64 factor 0.7"}


def workload_name(decoder, code, T=T_ITERS, snr=SNR_DB):
    return f"{DECODERS[decoder]}, {T} iters, {SHAPES[code]}, AWGN {snr} dB reference sign convention"


def det_weights(T):
    t = np.arange(T, dtype=np.float64)
    return (0.75 + t / 64).astype(np.float32), (1 - t / 32).astype(np.float32)


def build_decoder(L, code, kind, T=None):
    import torch
    T = T_ITERS if T is None else T
    b, a = det_weights(T)

    def fill(table, column):
        with torch.no_grad():
            table.copy_(torch.from_numpy(column)[:, None].expand_as(table))

    if kind == "n2d2":
        dec = L.Neural2DMinSumDecoder(code, weight_sharing_type=2, max_iterations=T)
        fill(dec._beta_table, b)
        fill(dec._alpha_table, a)
    elif kind == "n2d1":
        dec = L.Neural2DMinSumDecoder(code, weight_sharing_type=1, max_iterations=T)
        fill(dec._beta_table, b)
    elif kind == "oms2":
        dec = L.Neural2DOffsetMinSumDecoder(code, weight_sharing_type=2, max_iterations=T)
        with torch.no_grad():
            dec._beta_table.fill_(0.15)
            dec._alpha_table.fill_(0.02)
    elif kind == "nnms":
        dec = L.NeuralMinSumDecoder(code, max_iterations=T)
        fill(dec._beta_table, b)
    elif kind == "rcq":
        dec = L.RCQMinSumDecoder(code, bc=3, bv=8, quantizer_params=QP, max_iterations=T)
    elif kind == "wrcq1":
        dec = L.WeightedRCQDecoder(code, bc=3, bv=8, quantizer_params=QP, weight_sharing_type=1, max_iterations=T)
        fill(dec._beta_table, b)
    else:
        dec = L.BasicMinSumDecoder(code, factor=0.7)
    return dec


def algorithmic_bytes(code, kind, posterior):
    """bytes per frame-iteration at storage width, and per kernel launch per frame.  With `posterior`
    the final variable-node pass also writes the frame's posterior row entries."""
    g = code.graph
    E, n = g.E, g.n
    if kind in ("rcq", "wrcq1"):
        w, cw = 4, 1
    else:
        w = cw = 8 if kind == "basic" else 4
    return dict(frame_iter=(2 * w + 2 * cw) * E + w * n, cn=(w + cw) * E, vn=(w + cw) * E + w * n,
                vn_final=cw * E + w * n + (w * n if posterior else 0))


def oracle_setup(code, kind, T=T_ITERS):
    from oracle.restatement import (MODE_NMS, MODE_OFFSET, MODE_RCQ, MODE_WRCQ, SparseGraph, quantizer_schedule,
                                    quantizer_thresholds)
    g = code.graph
    og = SparseGraph.from_coo(g.n, g.m, g.edge_check, g.check_var)
    b, a = det_weights(T)
    kw = dict(T=T, want_posterior=False)
    thr = np.array([quantizer_thresholds(3, C, gm) for C, gm in QP]).astype(np.float32)
    if kind == "n2d2":
        kw.update(mode=MODE_NMS, beta=np.tile(b[:, None], (1, g.E)), alpha=np.tile(a[:, None], (1, g.n)))
    elif kind in ("n2d1", "nnms"):
        kw.update(mode=MODE_NMS, beta=np.tile(b[:, None], (1, g.E)))
    elif kind == "oms2":
        kw.update(mode=MODE_OFFSET, beta=np.full((T, g.E), np.float32(0.15)), alpha=np.full((T, g.n), np.float32(0.02)))
    elif kind == "rcq":
        kw.update(mode=MODE_RCQ, bc=3, thresholds=thr, quantizer_of_iter=quantizer_schedule(T, 3))
    elif kind == "wrcq1":
        kw.update(mode=MODE_WRCQ, bc=3, thresholds=thr, quantizer_of_iter=quantizer_schedule(T, 3),
                  beta=np.tile(b[:, None], (1, g.E)), alpha=np.ones((T, g.n), np.float32))
    else:
        kw.update(mode=MODE_NMS, dtype=np.float64, beta=np.full((T, g.E), 0.7))
    return og, kw


def host_llrs(code, frames, seed, dtype=np.float32):
    rng = np.random.default_rng(seed)
    s2 = 10 ** (-SNR_DB / 10)
    return (2 * (-1.0 + np.sqrt(s2) * rng.standard_normal((frames, code.n), dtype=np.float32)) / s2).astype(dtype)


def cpu_leg(code, kind, seconds, threads, posterior):
    """Time the oracle port on a bounded sample of the same workload.  Returns the `cpu_baseline` dict."""
    from oracle import capi as O
    og, kw = oracle_setup(code, kind)
    kw["want_posterior"] = posterior
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
                       f"avg iterations {float(res.iterations.mean()):.2f}")


class ClockSampler:
    """nvidia-smi polling every 200 ms (the profiling recipe's clocks line) in a child process.  Its START-UP stalls the
    GPU once for 10-100 ms (measured: a 90 ms step took 97-189 ms right after the process came up,
    profiles/r02aj_sampler_effect.log), so it is started well before the timed region and only the samples taken between
    `mark()` and `stop()` -- the timed region -- are reported."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.idx = gpu_index
        self.t_mark = None

    def start(self):
        if self.proc is not None:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def wait_first_sample(self, timeout=8.0):
        """Block until the child has written a sample: its start-up (and the GPU stall that comes with it) is over."""
        if self.proc is None:
            return
        t0 = time.time()
        while time.time() - t0 < timeout and os.path.getsize(self.tmp.name) == 0 and self.proc.poll() is None:
            time.sleep(0.05)

    def mark(self):
        self.t_mark = time.time()

    @staticmethod
    def _stamp(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self):
        t_end = time.time()
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
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.tmp.read().splitlines()
        for windowed in (True, False):     # (no sample inside the window -- clock skew, a very short region: use them all)
            sm, mx, reasons, power = [], [], set(), []
            for line in lines:
                f = [x.strip() for x in line.split(",")]
                if len(f) < 10:
                    continue
                ts = self._stamp(f[0])
                if windowed and self.t_mark is not None and ts is not None and not (self.t_mark - 0.2 <= ts <= t_end + 0.2):
                    continue
                try:
                    sm.append(float(f[2])); mx.append(float(f[3])); power.append(float(f[4]))
                except ValueError:
                    continue
                for name, val in zip(names, f[6:10]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                break
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


_UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def ncu_traffic(kind, code_name, kernel, frames_padded):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed `ncu --set full`
    raw page of this round (profiles/r02_ncu_full_<decoder>_<code>_<frames>frames_raw.csv).  The capture taken at
    the benchmarked frame count is used as is; otherwise the nearest one is scaled per frame.  (None, why) if absent."""
    files = glob.glob(os.path.join(ROOT, "profiles", f"r02_ncu_full_{kind}_{code_name}_*frames_raw.csv"))
    best = None
    for f in files:
        try:
            frames = int(os.path.basename(f).split("_")[-2].replace("frames", ""))
        except ValueError:
            continue
        if best is None or abs(frames - frames_padded) < abs(best[1] - frames_padded):
            best = (f, frames)
    if best is None:
        return None, "no r02 ncu capture committed for this (decoder, code)"
    f, frames = best
    rows = list(csv.reader(open(f, newline="")))
    head, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(head)}
    want = [kernel] if kernel != "cn_kernel" else ["cn_kernel", "cn_wide_kernel"]
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        hit = next((w for w in want if w + "<" in name), None)
        if hit is None:
            continue
        targs = [a.strip() for a in name.split(hit + "<")[1].split(">")[0].split(",")]
        if hit == "vn_kernel" and len(targs) >= 3 and targs[2] != "0":
            continue    # the dominant variable-node launch is the non-final variant <Real, QUANT, FINAL=0, POST>
        if True:
            try:
                rd = float(r[col["dram__bytes_read.sum"]]) * _UNIT_SCALE[units[col["dram__bytes_read.sum"]]]
                wr = float(r[col["dram__bytes_write.sum"]]) * _UNIT_SCALE[units[col["dram__bytes_write.sum"]]]
            except (KeyError, ValueError):
                continue
            scale = frames_padded / frames
            note = os.path.relpath(f, ROOT) + (f" (launch `{name}`, {frames} frames" +
                                              ("" if frames == frames_padded else f", scaled x{scale:g} per frame") + ")")
            return (rd + wr) * scale, note
    return None, f"{os.path.relpath(f, ROOT)} holds no launch of {kernel}"


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


# =====================================================================================================
# reference arm: the reference's algorithm on the host cores (C port in oracle/; the reference itself is pure
# Python and cannot travel to the GPU box).  No part of the CUDA package is imported here.
# =====================================================================================================
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import capi as O
    L = host_only_package()
    code = make_code(L, args.code)
    threads = os.cpu_count() or 1
    og, kw = oracle_setup(code, args.decoder)
    kw["want_posterior"] = not args.decode_only
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
        "config": {"workload": workload_name(args.decoder, args.code), "frames_per_step": frames,
                   "call": "decode() (decisions only)" if args.decode_only else "forward() (decisions + posterior + iterations)",
                   "note": "reference algorithm timed as its C port (the reference is pure Python and cannot travel to the GPU box)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{frames} frames per step x {args.steps} steps, avg iterations {float(res.iterations.mean()):.2f}"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# =====================================================================================================
# our arm
# =====================================================================================================
class Rig:
    """Per-process context of the GPU arm: rank / world, device, barrier and max-over-ranks reduction."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa = bind_to_gpu_numa_node(self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed_device_steps(self, fn, steps):
        """`steps` calls with a CUDA event between consecutive calls: (median, mean) device time per step in ms, each the
        max over ranks, and the last result.  The median is what a step costs; the mean also carries whatever stalled
        the GPU from outside during the region (measurement hygiene)."""
        torch = self.torch
        self.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        out = None
        for k in range(steps):
            out = fn()
            ev[k + 1].record()
        self.barrier()
        per = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(steps))
        median = per[len(per) // 2] if len(per) % 2 else 0.5 * (per[len(per) // 2 - 1] + per[len(per) // 2])
        return self.max_over_ranks(median), self.max_over_ranks(ev[0].elapsed_time(ev[steps]) / steps), out

    def timed_device(self, fn, steps, warmup):
        """W warm-ups, then K steps bracketed by barrier + synchronize, CUDA events on the launching stream, max over
        ranks.  Returns (ms per step, last result)."""
        torch = self.torch
        for _ in range(warmup):
            out = fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps, out


def kernel_roofline(prof, ab, Bp, steps, T, step_ms, kind, code_name):
    """Roofline of the dominant kernel from the library's per-launch CUDA-event durations (profiling mode) inside the
    timed region: algorithmic bytes per launch / average launch duration, against the measured HBM peak."""
    peak, peak_src = peak_hbm()
    if prof["vn_launches"] == 0 or prof["cn_launches"] == 0:
        # on-chip decode (small codes): one launch for all T iterations, the messages never reach HBM.  The same
        # algorithmic byte count over the step time is an EFFECTIVE bandwidth and may exceed the HBM peak.
        step_bytes = T * ab["frame_iter"] * Bp
        gbs = step_bytes / (step_ms * 1e-3) / 1e9
        zero = {"gbs": 0.0, "frac": 0.0, "ms_total": 0.0, "launches": 0}
        return {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                "traffic_source": "on-chip decode: messages stay in shared memory (effective bandwidth)", "kernel": "small_decode_kernel",
                "peak_source": peak_src, "bytes_per_launch": step_bytes, "avg_launch_ms": step_ms,
                "vn_kernel": zero, "cn_kernel": zero, "other_ms_total": prof["other_ms"],
                "whole_step": {"gbs": gbs, "frac": gbs / peak, "bytes_per_frame_iter": ab["frame_iter"]}}
    vn_per_step = prof["vn_launches"] / steps
    vn_bytes = ((vn_per_step - 1) * ab["vn"] + ab["vn_final"]) / vn_per_step * Bp
    cn_bytes = ab["cn"] * Bp
    vn_ms = prof["vn_ms"] / max(prof["vn_launches"], 1)
    cn_ms = prof["cn_ms"] / max(prof["cn_launches"], 1)
    vn_gbs = vn_bytes / (vn_ms * 1e-3) / 1e9 if vn_ms > 0 else 0.0
    cn_gbs = cn_bytes / (cn_ms * 1e-3) / 1e9 if cn_ms > 0 else 0.0
    dominant = "vn_kernel" if prof["vn_ms"] >= prof["cn_ms"] else "cn_kernel"
    ach = vn_gbs if dominant == "vn_kernel" else cn_gbs
    step_bytes = T * ab["frame_iter"] * Bp
    traffic, traffic_src = ncu_traffic(kind, code_name, dominant, Bp)
    return {
        "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
        "traffic": traffic, "traffic_source": traffic_src,
        "kernel": dominant, "peak_source": peak_src,
        "bytes_per_launch": vn_bytes if dominant == "vn_kernel" else cn_bytes,
        "avg_launch_ms": vn_ms if dominant == "vn_kernel" else cn_ms,
        "vn_kernel": {"gbs": vn_gbs, "frac": vn_gbs / peak, "ms_total": prof["vn_ms"], "launches": prof["vn_launches"]},
        "cn_kernel": {"gbs": cn_gbs, "frac": cn_gbs / peak, "ms_total": prof["cn_ms"], "launches": prof["cn_launches"]},
        "other_ms_total": prof["other_ms"],
        # device time of a step not covered by any of this library's kernels (host round trips, launch gaps)
        "uncovered_ms_per_step": step_ms - (prof["vn_ms"] + prof["cn_ms"] + prof["other_ms"]) / steps,
        "whole_step": {"gbs": step_bytes / (step_ms * 1e-3) / 1e9, "frac": step_bytes / (step_ms * 1e-3) / 1e9 / peak,
                       "bytes_per_frame_iter": ab["frame_iter"]},
    }


def device_leg(rig, L, code, code_name, kind, B, steps, warmup, posterior, sampler=None, min_timed_ms=0.0):
    """One device-resident leg: LLRs generated on the device (Philox, distinct frames per rank, reference sign
    convention at SNR_DB), resident before timing; returns the result record and the pieces the caller reuses.
    `min_timed_ms` (configuration legs only; the headline times exactly `steps`): a leg whose step is a few
    milliseconds gets enough steps that the timed region is at least this long -- measured over 10 ms, the idle gap of
    the barrier in front of it (clocks ramping back up) dominated the (7,4) leg: 1.9 ms per step on one run, 4.8 on
    the next, 32 with eight ranks."""
    torch = rig.torch
    g = code.graph
    gc.collect()                       # (buffers of an earlier leg freed inside the timed region would stall it)
    if sampler is not None:
        sampler.start()                # its start-up stalls the GPU once: long before the timed region
    dec = build_decoder(L, code, kind)
    eng = dec._engine(rig.local_rank)
    eng.reserve(B)
    llr = L.awgn_llr(g.n, B, SNR_DB, seed=1234, frame0=rig.rank * B, llr_sign=-1, device=rig.local_rank)
    if kind == "basic":
        llr = llr.double()
    torch.cuda.synchronize()

    def step():
        return eng.decode_device(llr, want_posterior=posterior)

    if sampler is not None:
        sampler.wait_first_sample()
    for _ in range(warmup):
        out = step()
    if min_timed_ms > 0:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = step()
        e1.record()
        torch.cuda.synchronize()
        est = rig.max_over_ranks(e0.elapsed_time(e1))
        steps = max(steps, min(500, int(np.ceil(min_timed_ms / max(est, 1e-3)))))
        for _ in range(min(steps, 100) if est < 10.0 else 0):    # keep the device busy up to the timed region
            out = step()
    rig.barrier()
    avg_iters = float(out[2].float().mean().item())
    on_chip = eng.profile_read(reset=True)["small_decodes"] > 0
    # per-launch CUDA events for the check / variable node kernels (a one-launch on-chip decode has nothing to split)
    eng.profile_mode(0 if on_chip else 1)
    if sampler is not None:
        sampler.mark()
    ms_mean = None
    if min_timed_ms > 0:     # configuration legs: per-step events, the median is reported (the headline: exactly K steps, mean)
        ms, ms_mean, out = rig.timed_device_steps(step, steps)
    else:
        ms, out = rig.timed_device(step, steps, 0)
    clocks = sampler.stop() if sampler is not None else None
    prof = eng.profile_read(reset=True)
    eng.profile_mode(0)
    fps = rig.world * B * steps / (ms * steps / 1e3)
    ab = algorithmic_bytes(code, kind, posterior)
    roof = kernel_roofline(prof, ab, prof["frames_padded"], steps, T_ITERS, ms, kind, code_name)
    if ms_mean is not None and "uncovered_ms_per_step" in roof:   # of the whole region, stalls included
        roof["uncovered_ms_per_step"] = ms_mean - (prof["vn_ms"] + prof["cn_ms"] + prof["other_ms"]) / steps
    rec = {"frames_per_s": fps, "info_gbps": fps * code.k / 1e9, "ms_per_step": ms, "ms_per_step_mean": ms_mean,
           "avg_iterations": avg_iters, "steps": steps,
           "edge_msgs_per_s": fps * T_ITERS * 2 * g.E, "launches": int(prof["launches"]), "roofline": roof}
    return rec, dec, eng, llr, out, clocks


def e2e_leg(rig, L, code, eng, llr, B, steps, warmup, posterior, check_bits, packed=True):
    """ldpc_decode_host[_packed] on pinned host buffers: H2D of the LLRs, decode, D2H of every output, all inside the
    timed region (host wall clock around the blocking call; max over ranks).  `packed`: the hard decisions come back as
    bit-packed rows (n/8 bytes per frame, straight from the kernels' decision words) instead of one byte per bit."""
    torch = rig.torch
    g = code.graph
    rdt = np.float64 if eng.dtype == np.float64 else np.float32
    pins = {"llr": L.PinnedBuffer((B, g.n), rdt),
            "bits": L.PinnedBuffer((B, eng.row_words), np.uint32) if packed else L.PinnedBuffer((B, g.n), np.uint8),
            "iterations": L.PinnedBuffer((B,), np.int32), "success": L.PinnedBuffer((B,), np.uint8)}
    if posterior:
        pins["posterior"] = L.PinnedBuffer((B, g.n), rdt)
    chunk = 8192
    for s in range(0, B, chunk):   # fill the pinned input from the device-generated LLRs
        pins["llr"].array[s:s + chunk] = llr[s:s + chunk].cpu().numpy()
    outs = {k: p.array for k, p in pins.items() if k != "llr"}
    for _ in range(max(1, min(warmup, 2))):
        eng.decode_host(pins["llr"].array, want_posterior=posterior, out=outs, packed_bits=packed)
    rig.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.decode_host(pins["llr"].array, want_posterior=posterior, out=outs, packed_bits=packed)
    torch.cuda.synchronize()
    dt = rig.max_over_ranks(time.perf_counter() - t0)
    fps = rig.world * B * steps / dt
    got = eng.unpack_rows(pins["bits"].array[:256]) if packed else pins["bits"].array[:256]
    same = bool(np.array_equal(got, check_bits[:256].cpu().numpy()))
    bits_bytes = B * eng.row_words * 4 if packed else B * g.n
    d2h = bits_bytes + B * 4 + B + (B * g.n * np.dtype(rdt).itemsize if posterior else 0)
    rec = {"value": fps * code.k / 1e9, "unit": UNIT, "frames_per_s": fps,
           "h2d_bytes_per_step": int(B * g.n * np.dtype(rdt).itemsize), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": 1e3 * dt / steps, "steps": steps,
           "api": ("ldpc_decode_host_packed" if packed else "ldpc_decode_host") + " (pinned host LLR in; hard decisions" +
                  (" as bit-packed rows" if packed else " one byte per bit") + (" + posteriors" if posterior else "") +
                  " + iterations + success out)",
           "matches_device_path": same}
    for p in pins.values():
        p.free()
    return rec


def link_ceiling(rig, L, nbytes):
    """What the host link of this box delivers for plain pinned copies of one step's size, all ranks at once: H2D
    alone, D2H alone, and both directions together (the e2e call moves its inputs and outputs concurrently)."""
    torch = rig.torch
    n = int(nbytes // 4)
    pin_a, pin_b = L.PinnedBuffer((n,), np.float32), L.PinnedBuffer((n,), np.float32)
    ha, hb = torch.from_numpy(pin_a.array), torch.from_numpy(pin_b.array)
    da = torch.empty(n, dtype=torch.float32, device=rig.dev)
    db = torch.zeros(n, dtype=torch.float32, device=rig.dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h, reps=2):
        rig.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    da.copy_(ha, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    hb.copy_(db, non_blocking=True)
        torch.cuda.synchronize()
        return rig.max_over_ranks(time.perf_counter() - t0) / reps

    run(True, True, 1)
    gb = n * 4 / 1e9
    t_h, t_d, t_b = run(True, False), run(False, True), run(True, True)
    out = {"bytes_per_copy": n * 4, "ranks": rig.world, "h2d_gbs_per_rank": gb / t_h, "d2h_gbs_per_rank": gb / t_d,
           "bidirectional_gbs_per_rank_each_way": gb / t_b}
    del ha, hb, da, db
    pin_a.free()
    pin_b.free()
    return out


def release(*objs):
    import gc
    for o in objs:
        close = getattr(o, "close", None)
        if close:
            close()
    gc.collect()


def config_legs(rig, L, args):
    """One short leg per remaining BASELINE.json configuration (configs[0], [2], [3], [4])."""
    torch = rig.torch
    legs = []
    steps, warmup = 5, 3

    def kernel_leg(tag, code_name, kind, B, posterior):
        code = make_code(L, code_name)
        rec, dec, eng, llr, out, _ = device_leg(rig, L, code, code_name, kind, B, steps, warmup, posterior, min_timed_ms=250.0)
        r = rec["roofline"]
        leg = {"config": tag, "workload": workload_name(kind, code_name), "frames_per_gpu": B,
               "call": "forward()" if posterior else "decode()", "steps": rec["steps"], "warmup": warmup,
               "frames_per_s": rec["frames_per_s"], "info_gbps": rec["info_gbps"], "ms_per_step": rec["ms_per_step"],
               "timing": "median of per-step CUDA-event times, max over ranks (ms_per_step_mean: the whole region / steps, "
                         "which also carries stalls from outside the process)",
               "ms_per_step_mean": rec["ms_per_step_mean"], "avg_iterations": rec["avg_iterations"],
               "roofline": {"cn_frac": r["cn_kernel"]["frac"], "vn_frac": r["vn_kernel"]["frac"],
                            "whole_step_frac": r["whole_step"]["frac"], "bytes_per_frame_iter": r["whole_step"]["bytes_per_frame_iter"],
                            "other_kernels_ms_per_step": r["other_ms_total"] / rec["steps"],
                            "uncovered_ms_per_step": r.get("uncovered_ms_per_step"), "peak_gbs": r["peak"]}}
        del llr, out
        release(eng)
        del dec, eng
        torch.cuda.empty_cache()
        return leg

    # C1: create_test_ldpc_code() + BasicMinSumDecoder factor 0.7, 10 iterations, float64 (configs[0])
    legs.append(kernel_leg("C1", "h74", "basic", 1 << 22, False))
    # C3: RCQMinSumDecoder bc=3 bv=8 on the (16200,7200) shape (configs[2]); decode() returns no posterior
    legs.append(kernel_leg("C3", "dvbs2", "rcq", args.frames, False))
    # C4: WeightedRCQ type-1 degree weights on the (9472,8192)-shaped QC code (configs[3]); forward() with posterior
    legs.append(kernel_leg("C4", "qc", "wrcq1", args.frames, True))
    # C5: FER/BER Monte-Carlo sweep through LDPSimulator.simulate_decoder (simulation_framework.py:141-176) over the
    # SimulationConfig default grid 0..6 dB step 0.5, T = 50, frames sharded over the ranks (configs[4])
    legs.append(mc_sweep_leg(rig, L, args))
    # one posterior-training step (forward with message history + backward kernels)
    legs.append(training_leg(rig, L, args))
    return legs


def training_leg(rig, L, args):
    """Forward + backward of PosteriorJointTrainer's step (training_framework.py:108-165) through ldpc_train_forward /
    ldpc_train_backward: N-2D-NMS type 2, 10 iterations, 8192 frames of the (16200,7200)-shaped code per GPU, loss =
    binary_cross_entropy_with_logits(-posterior, 0).  Bytes: forward 16E + 4n per frame-iteration (the history slices ARE
    the message arrays), backward 24E (v2c / c2v history read, gradient rows read and written)."""
    torch = rig.torch
    B, T = 8192, T_ITERS
    code = make_code(L, "dvbs2", T)
    E, n = code.graph.E, code.n
    dec = build_decoder(L, code, "n2d2", T)
    eng = dec._engine(rig.local_rank)
    llr = L.awgn_llr(n, B, 2.0, seed=99, frame0=rig.rank * B, llr_sign=1, device=rig.local_rank)

    def forward():
        return eng.train_forward(llr)

    def step():
        _, post, _, _ = eng.train_forward(llr)
        g = torch.sigmoid(-post) * (-1.0 / post.numel())     # d loss / d posterior
        return eng.train_backward(g)

    it = float(forward()[2].float().mean().item())
    t_fwd, _ = rig.timed_device(forward, 5, 3)
    t_step, _ = rig.timed_device(step, 5, 3)
    peak, _ = peak_hbm()
    fwd_gbs = (16 * E + 4 * n) * it * B / (t_fwd * 1e-3) / 1e9
    bwd_gbs = 24 * E * it * B / (max(t_step - t_fwd, 1e-6) * 1e-3) / 1e9
    leg = {"config": "TRAIN", "workload": f"posterior-training step (forward + backward), {DECODERS['n2d2']}, {SHAPES['dvbs2']}, "
                                          f"AWGN 2.0 dB converging sign convention",
           "frames_per_gpu": B, "steps": 5, "warmup": 3, "avg_iterations": it,
           "frames_per_s": rig.world * B / (t_step * 1e-3), "ms_per_step": t_step, "forward_ms": t_fwd, "backward_ms": t_step - t_fwd,
           "roofline": {"forward_frac": fwd_gbs / peak, "backward_frac": bwd_gbs / peak, "forward_gbs": fwd_gbs, "backward_gbs": bwd_gbs,
                        "peak_gbs": peak}}
    release(eng)
    del dec, eng, llr
    torch.cuda.empty_cache()
    return leg


def mc_sweep_leg(rig, L, args):
    torch, dist = rig.torch, rig.dist
    T = 50
    code = make_code(L, "dvbs2", T)
    dec = L.Neural2DMinSumDecoder(code, weight_sharing_type=2, max_iterations=T)
    t = np.arange(T, dtype=np.float64)
    with torch.no_grad():
        dec._beta_table.copy_(torch.from_numpy(np.minimum(0.75 + t / 64, 1.0).astype(np.float32))[:, None].expand_as(dec._beta_table))
        dec._alpha_table.fill_(1.0)
    per_gpu = min(args.frames, 32768)
    cfg = L.SimulationConfig(snr_range=(0.0, 6.0), snr_step=0.5, max_frames=per_gpu * rig.world * 2, max_errors=200,
                             batch_frames=per_gpu, seed=20, save_results=False)
    sim = L.LDPSimulator(cfg)
    for snr in (0.5, 1.0, 1.5, 3.0):     # warm-up: the workspaces of every compaction pattern are allocated here
        sim.simulate_single_snr(dec, code, snr, per_gpu * rig.world, 10 ** 9)
    for snr in (0.0, 1.0):               # ... and the short rounds of the adaptive schedule
        sim.simulate_single_snr(dec, code, snr, cfg.max_frames, cfg.max_errors)
    rig.barrier()
    t0 = time.perf_counter()
    res = sim.simulate_decoder(dec, code, "N-2D-NMS Type 2")
    torch.cuda.synchronize()
    dt = rig.max_over_ranks(time.perf_counter() - t0)
    frames = int(sum(res.total_frames))
    leg = {"config": "C5", "workload": f"LDPSimulator.simulate_decoder, {DECODERS['n2d2']}, T = {T}, {SHAPES['dvbs2']}, "
                                       f"SNR 0..6 dB step 0.5 (13 points), all-zero codeword, converging sign convention, Philox AWGN on device",
           "max_frames_per_point": cfg.max_frames, "max_errors": cfg.max_errors, "batch_frames_per_gpu": per_gpu,
           "rounds": "pilot round of 4 x max_errors frames, then rounds sized from the error rate so far (at most batch_frames "
                     "per GPU); the reference's sequential stop rule is applied exactly, so the counters do not depend on it",
           "frames": frames, "seconds": dt, "frames_per_s": frames / dt, "info_gbps": frames / dt * code.k / 1e9,
           "snr_db": [float(s) for s in res.snr_values], "fer": res.frame_error_rates, "ber": res.bit_error_rates,
           "avg_iterations": res.average_iterations, "frames_per_point": res.total_frames,
           "frame_errors_per_point": res.total_errors, "seconds_per_point": res.simulation_times}
    # Monte-Carlo parity across GPU counts: one sharded round of `world * R` global frames (all-reduced counters)
    # against the same global frame range decoded by rank 0 alone.  Noise is keyed by the global frame index.
    parity = None
    if rig.world > 1:
        R = 4096
        eng = dec._engine(rig.local_rank)
        parity = True
        for snr in (1.5, 2.5):
            c = torch.zeros(4, dtype=torch.int64, device=rig.dev)
            eng.mc_round(snr, R, seed=77, frame0=rig.rank * R, llr_sign=1, counters=c)
            dist.all_reduce(c)
            alone = torch.zeros(4, dtype=torch.int64, device=rig.dev)
            if rig.rank == 0:
                eng.mc_round(snr, R * rig.world, seed=77, frame0=0, llr_sign=1, counters=alone)
            dist.broadcast(alone, src=0)
            parity = parity and bool(torch.equal(c, alone))
            leg.setdefault("mc_parity_counters", []).append({"snr_db": snr, "sharded": c.tolist(), "rank0_alone": alone.tolist()})
    leg["mc_parity"] = parity
    release(*dec._engines.values())
    return leg


def run_ours(args):
    rig = Rig(args)
    torch = rig.torch
    if rig.world != args.gpus and rig.world > 1:
        args.gpus = rig.world
    import ldpc_b200 as L

    code = make_code(L, args.code)
    g = code.graph
    kind = args.decoder
    B = int(args.frames)
    f64 = kind == "basic"
    posterior = not args.decode_only and kind not in ("basic", "rcq")   # decode() of those two returns no posterior

    # ---- headline: device-resident, the reference's forward() surface ----
    sampler = ClockSampler(rig.local_rank) if rig.rank == 0 else None
    head, dec, eng, llr, out, clocks = device_leg(rig, L, code, args.code, kind, B, args.steps, args.warmup, posterior, sampler)
    bits = out[0]
    # ---- the same step without the posterior (decode() surface) ----
    decode_only = None
    if posterior:
        k = max(3, min(args.steps, 10))
        eng.profile_read(reset=True)
        eng.profile_mode(1)
        ms, _ = rig.timed_device(lambda: eng.decode_device(llr, want_posterior=False), k, 2)
        prof = eng.profile_read(reset=True)
        eng.profile_mode(0)
        fps = rig.world * B / (ms / 1e3)
        r = kernel_roofline(prof, algorithmic_bytes(code, kind, False), prof["frames_padded"], k, T_ITERS, ms, kind, args.code)
        decode_only = {"frames_per_s": fps, "info_gbps": fps * code.k / 1e9, "ms_per_step": ms, "steps": k,
                       "vn_frac": r["vn_kernel"]["frac"], "cn_frac": r["cn_kernel"]["frac"], "whole_step_frac": r["whole_step"]["frac"]}

    # ---- end to end through the host-buffer C-ABI call (pinned host buffers, copies inside) ----
    e2e = e2e_decode_only = e2e_u8 = None
    if not args.no_e2e:
        e2e = e2e_leg(rig, L, code, eng, llr, B, args.steps, args.warmup, posterior, bits)
        few = max(3, min(args.steps, 5))
        if posterior:
            e2e_decode_only = e2e_leg(rig, L, code, eng, llr, B, few, 1, False, bits)
        e2e_u8 = e2e_leg(rig, L, code, eng, llr, B, few, 1, False, bits, packed=False)   # round 1's call, for continuity
        e2e["link_ceiling"] = link_ceiling(rig, L, e2e["h2d_bytes_per_step"])
    del llr, out, bits
    release(eng)
    del dec, eng
    torch.cuda.empty_cache()

    configs = None if args.no_configs else config_legs(rig, L, args)

    cpu = None
    if rig.rank == 0 and rig.world == 1 and not args.no_cpu:
        cpu = cpu_leg(code, kind, args.cpu_seconds, os.cpu_count() or 1, posterior)

    if rig.rank == 0:
        roof = head["roofline"]
        line = {
            "metric": METRIC, "value": head["info_gbps"], "unit": UNIT, "n_gpus": rig.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if f64 else "f32", "data": "synthetic",
            "frames_per_s": head["frames_per_s"], "edge_msgs_per_s": head["edge_msgs_per_s"],
            "avg_iterations": head["avg_iterations"],
            "config": {"workload": workload_name(kind, args.code),
                       "call": "forward(): decisions + posterior + iterations" if posterior else "decode(): decisions + success + iterations",
                       "frames_per_gpu": B, "global_frames": rig.world * B,
                       "n": g.n, "k": code.k, "E": g.E, "iterations": T_ITERS, "early_stop": True,
                       "l2": "inputs larger than L2 (message arrays %.1f GB per GPU)" % (algorithmic_bytes(code, kind, False)["cn"] * B / 1e9),
                       "parallelism": f"frames sharded over {rig.world} GPU(s), no data-path collective",
                       "host_affinity": (f"{len(rig.numa)} CPUs next to the GPU (NVML)" if rig.numa else "unbound")},
            "roofline": roof, "e2e": e2e, "decode_only": decode_only, "e2e_decode_only": e2e_decode_only,
            "e2e_decode_only_byte_per_bit": e2e_u8,
            "cpu_baseline": cpu, "configs": configs,
            "mc_parity": (next((c.get("mc_parity") for c in configs if c["config"] == "C5"), None) if configs else None),
            "gpu_launches": head["launches"],
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if rig.world > 1:
        rig.dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

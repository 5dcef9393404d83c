"""Monte-Carlo FER/BER harness -- B200-native counterpart of the reference's
``simulation_framework.py`` (SimulationConfig :27-38, SimulationResult :40-69, LDPSimulator :71-382,
create_test_decoders :384-420).

The reference draws one frame at a time on the CPU (:110-131).  Here a *round* of ``batch_frames``
frames per GPU is generated (Philox AWGN), decoded and counted entirely on the device by
``ldpc_mc_round``; frames are sharded over the ranks of ``torch.distributed`` when it is initialised
(one process per GPU, NCCL) and the only collective is an all-reduce of the four int64 counters
{frame_errors, bit_errors, total_iterations, total_frames} per round.  Noise is a pure function of
(seed, global frame index, variable), so results do not depend on the number of GPUs, and the
reference's sequential stop rule (`while total_frames < max_frames and frame_errors < max_errors`) is
reproduced exactly by truncating the last round in global frame order.

Plotting (:218-336) is out of scope (matplotlib is not part of the hot path).
"""
from __future__ import annotations

import json
import logging
import os
import time
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import _plotting
from .ldpc_decoder import BasicMinSumDecoder, LDPCCode
from .neural_2d_decoder import Neural2DMinSumDecoder, Neural2DOffsetMinSumDecoder
from .neural_minsum_decoder import NeuralMinSumDecoder, NeuralOffsetMinSumDecoder
from .rcq_decoder import RCQMinSumDecoder, WeightedRCQDecoder

logger = logging.getLogger(__name__)


@dataclass
class SimulationConfig:
    """Fields of simulation_framework.py:27-38 plus the batching knobs of the device harness."""
    snr_range: Tuple[float, float] = (0.0, 6.0)
    snr_step: float = 0.5
    max_frames: int = 10000
    max_errors: int = 100
    min_frames: int = 1000            # unused by the reference too
    parallel_workers: int = 4         # reference: threads over decoders; here decoders run back to back
    device: str = 'cuda'
    save_results: bool = True
    results_dir: str = 'simulation_results'
    # ---- additive ----
    batch_frames: int = 8192          # frames per GPU per round
    seed: int = 0
    reference_convention: bool = False  # True: bit 0 -> -1 like ldpc_decoder.py:289 (FER ~ 1)
    exact_stop: bool = True           # truncate the last round at the frame where max_errors is reached
    adaptive_rounds: bool = True      # size the rounds from the error rate seen so far (needs exact_stop; same results)


class SimulationResult:
    """Container with the reference's fields (simulation_framework.py:40-69)."""

    def __init__(self, decoder_name: str, snr_values: List[float]):
        self.decoder_name = decoder_name
        self.snr_values = snr_values
        self.frame_error_rates: List[float] = []
        self.bit_error_rates: List[float] = []
        self.average_iterations: List[float] = []
        self.simulation_times: List[float] = []
        self.total_frames: List[int] = []
        self.total_errors: List[int] = []

    _LISTS = ("frame_error_rates", "bit_error_rates", "average_iterations", "simulation_times",
              "total_frames", "total_errors")

    def add_result(self, snr_idx: int, fer: float, ber: float, avg_iter: float, sim_time: float,
                   total_frames: int, total_errors: int):
        values = (fer, ber, avg_iter, sim_time, total_frames, total_errors)
        for name, value in zip(self._LISTS, values):
            lst = getattr(self, name)
            zero = 0 if name.startswith("total") else 0.0
            lst.extend([zero] * (snr_idx + 1 - len(lst)))
            lst[snr_idx] = value


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def split_round(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous share of a round of ``total`` frames for ``rank``: (offset, count)."""
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def next_round_frames(full: int, left: int, max_errors: int, total_frames: int, frame_errors: int,
                      world: int = 1, fer_hint: Optional[float] = None) -> int:
    """Frames of the next Monte-Carlo round.  The reference stops at the frame that brings the error count to
    ``max_errors`` (simulation_framework.py:110): with the exact stop rule a round larger than what is still needed
    decodes frames whose results are thrown away -- at 0 dB every frame fails, 200 errors need 200 frames, and a
    full round would decode tens of thousands at the maximum iteration count.  So: a pilot round of a few times
    ``max_errors`` frames, then rounds sized from the error rate seen so far (30 % head-room), never above ``full``
    (= ``batch_frames`` per GPU).  Depends only on the all-reduced counters, so every rank computes the same
    schedule, and -- frames being keyed by their global index and the last round truncated in frame order -- the
    reported numbers do not depend on it.  ``fer_hint``: the frame error rate of the same decoder at the previous
    (lower) SNR point of a sweep, an upper bound in expectation, replaces the pilot round."""
    if total_frames == 0 and fer_hint is not None:
        want = full if fer_hint <= 0.0 else int(1.3 * max_errors / fer_hint) + 64
    elif total_frames == 0:
        want = max(4 * max_errors, 1024)
    elif frame_errors == 0:
        want = full
    else:
        want = int(1.3 * (max_errors - frame_errors) * total_frames / frame_errors) + 64
    unit = 128 * world                       # whole warps of 4-frame lanes on every rank
    want = (max(want, unit) + unit - 1) // unit * unit
    return max(1, min(full, left, want))


def truncate_in_frame_order(bit_errors: np.ndarray, iterations: np.ndarray, errors_before: int,
                            max_errors: int) -> Tuple[int, int, int, int]:
    """Sequential stop rule on one round given per-frame results in global frame order: keep frames
    up to and including the one whose error brings the count to ``max_errors``.
    Returns (frame_errors, bit_errors, iterations, frames) of the kept prefix."""
    is_err = bit_errors > 0
    cum = errors_before + np.cumsum(is_err)
    hit = np.nonzero(cum >= max_errors)[0]
    keep = int(hit[0]) + 1 if hit.size else bit_errors.size
    return (int(is_err[:keep].sum()), int(bit_errors[:keep].sum()), int(iterations[:keep].sum()), keep)


class LDPSimulator:
    """Batched, multi-GPU Monte-Carlo simulator with the reference's method names."""

    def __init__(self, config: SimulationConfig):
        self.config = config
        self.results: Dict[str, SimulationResult] = {}
        self._fer_seen: Dict[int, Tuple[float, float, int]] = {}   # id(decoder) -> (snr_db, fer, max_errors) of its last point
        if config.save_results:
            os.makedirs(config.results_dir, exist_ok=True)

    # simulation_framework.py:85-139
    def simulate_single_snr(self, decoder: Callable, code: LDPCCode, snr_db: float, max_frames: int,
                            max_errors: int) -> Tuple[float, float, float, float, int, int]:
        cfg = self.config
        start = time.time()
        dist, rank, world = _dist()
        device = self._device()
        runner = self._make_round_runner(decoder, code, snr_db, device)
        per_gpu = int(cfg.batch_frames)
        round_counters = torch.zeros(4, dtype=torch.int64, device=device)
        fbe = torch.zeros(per_gpu, dtype=torch.int32, device=device)
        fit = torch.zeros(per_gpu, dtype=torch.int32, device=device)
        frame_errors = bit_errors = total_iterations = total_frames = 0
        # sweeps go up in SNR: the error rate of the previous point bounds this one's (round sizing only)
        prev = self._fer_seen.get(id(decoder))
        hint = prev[1] if prev is not None and prev[0] <= snr_db and prev[2] == max_errors else None
        while total_frames < max_frames and frame_errors < max_errors:
            round_frames = min(per_gpu * world, max_frames - total_frames)
            if cfg.adaptive_rounds and cfg.exact_stop:
                round_frames = next_round_frames(per_gpu * world, max_frames - total_frames, max_errors,
                                                 total_frames, frame_errors, world, hint)
            off, cnt = split_round(round_frames, world, rank)
            round_counters.zero_()
            if cnt > 0:
                runner(cnt, total_frames + off, round_counters, fbe, fit)
            if dist is not None:
                dist.all_reduce(round_counters)
            fe, be, it, nf = (int(v) for v in round_counters.tolist())
            if cfg.exact_stop and frame_errors + fe >= max_errors:
                be_all, it_all = self._gather_round(dist, world, round_frames, cnt, fbe, fit)
                fe, be, it, nf = truncate_in_frame_order(be_all, it_all, frame_errors, max_errors)
            frame_errors += fe
            bit_errors += be
            total_iterations += it
            total_frames += nf
        fer = frame_errors / total_frames if total_frames > 0 else 0.0
        self._fer_seen[id(decoder)] = (snr_db, fer, max_errors)
        ber = bit_errors / (total_frames * code.n) if total_frames > 0 else 0.0
        avg_iterations = total_iterations / total_frames if total_frames > 0 else 0.0
        return fer, ber, avg_iterations, time.time() - start, total_frames, frame_errors

    def _device(self) -> torch.device:
        return torch.device("cuda", torch.cuda.current_device())

    def _make_round_runner(self, decoder, code, snr_db, device):
        """Returns run(count, frame0, counters, frame_bit_errors, frame_iterations): one device round on
        global frames [frame0, frame0+count) accumulating into ``counters`` (int64[4])."""
        engine = decoder._engine(device.index)
        sign = -1 if self.config.reference_convention else 1
        seed = self.config.seed

        def run(count, frame0, counters, fbe, fit):
            engine.mc_round(snr_db, count, seed=seed, frame0=frame0, llr_sign=sign, counters=counters,
                            frame_bit_errors=fbe, frame_iterations=fit)
        return run

    @staticmethod
    def _gather_round(dist, world, round_frames, cnt, fbe, fit):
        """Per-frame results of the whole round in global frame order (only needed in the last round)."""
        if dist is None:
            return fbe[:cnt].cpu().numpy(), fit[:cnt].cpu().numpy()
        width = (round_frames + world - 1) // world
        mine = torch.zeros(2, width, dtype=torch.int32, device=fbe.device)
        mine[0, :cnt] = fbe[:cnt]
        mine[1, :cnt] = fit[:cnt]
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        be, it = [], []
        for r, p in enumerate(parts):
            _, c = split_round(round_frames, world, r)
            p = p.cpu().numpy()
            be.append(p[0, :c])
            it.append(p[1, :c])
        return np.concatenate(be), np.concatenate(it)

    # simulation_framework.py:141-176
    def simulate_decoder(self, decoder: Union[Callable, torch.nn.Module], code: LDPCCode,
                         decoder_name: str) -> SimulationResult:
        lo, hi = self.config.snr_range
        snr_values = np.arange(lo, hi + self.config.snr_step, self.config.snr_step)
        result = SimulationResult(decoder_name, snr_values.tolist())
        for snr_idx, snr_db in enumerate(snr_values):
            out = self.simulate_single_snr(decoder, code, float(snr_db), self.config.max_frames, self.config.max_errors)
            result.add_result(snr_idx, *out)
            logger.info("%s SNR %.1f dB: FER=%.2e BER=%.2e avg_iter=%.1f time=%.2fs frames=%d",
                        decoder_name, snr_db, out[0], out[1], out[2], out[3], out[4])
        self.results[decoder_name] = result
        return result

    # simulation_framework.py:178-216 (the reference fans decoders out to GIL-bound threads; one GPU
    # is already saturated by one decoder's batch, so decoders run back to back)
    def simulate_multiple_decoders(self, decoders: Dict[str, Union[Callable, torch.nn.Module]],
                                   code: LDPCCode) -> Dict[str, SimulationResult]:
        return {name: self.simulate_decoder(dec, code, name) for name, dec in decoders.items()}

    # simulation_framework.py:218-336 -- the four figures (matplotlib imported on demand, see _plotting.py)
    def plot_fer_curves(self, results: Dict[str, SimulationResult], save_path: Optional[str] = None, log_scale: bool = True):
        _plotting.curves(results, "fer", save_path, log_scale)

    def plot_ber_curves(self, results: Dict[str, SimulationResult], save_path: Optional[str] = None, log_scale: bool = True):
        _plotting.curves(results, "ber", save_path, log_scale)

    def plot_iteration_curves(self, results: Dict[str, SimulationResult], save_path: Optional[str] = None):
        _plotting.curves(results, "iterations", save_path)

    def plot_comprehensive_comparison(self, results: Dict[str, SimulationResult], save_path: Optional[str] = None):
        _plotting.comparison(results, save_path)

    # simulation_framework.py:338-382 -- same JSON schema
    def save_results(self, results: Dict[str, SimulationResult], filename: str):
        fields = ("decoder_name", "snr_values") + SimulationResult._LISTS
        blob = {name: {f: getattr(res, f) for f in fields} for name, res in results.items()}
        with open(os.path.join(self.config.results_dir, filename), "w") as fh:
            json.dump(blob, fh, indent=2)

    def load_results(self, filename: str) -> Dict[str, SimulationResult]:
        with open(os.path.join(self.config.results_dir, filename)) as fh:
            blob = json.load(fh)
        out = {}
        for name, data in blob.items():
            res = SimulationResult(data["decoder_name"], data["snr_values"])
            for f in SimulationResult._LISTS:
                setattr(res, f, data[f])
            out[name] = res
        return out


def create_test_decoders(code: LDPCCode) -> Dict[str, Union[Callable, torch.nn.Module]]:
    """The ten-decoder comparison set of simulation_framework.py:384-420 with its canonical parameters."""
    qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
    decoders: Dict[str, Union[Callable, torch.nn.Module]] = {
        'Basic MinSum': BasicMinSumDecoder(code, factor=0.7),
        'N-NMS': NeuralMinSumDecoder(code, max_iterations=10),
        'N-OMS': NeuralOffsetMinSumDecoder(code, max_iterations=10),
    }
    for weight_type in (1, 2, 3, 4):
        decoders[f'N-2D-NMS Type {weight_type}'] = Neural2DMinSumDecoder(
            code, weight_sharing_type=weight_type, max_iterations=10)
    decoders['N-2D-OMS Type 2'] = Neural2DOffsetMinSumDecoder(code, weight_sharing_type=2, max_iterations=10)
    decoders['RCQ MinSum'] = RCQMinSumDecoder(code, bc=3, bv=8, quantizer_params=qp, max_iterations=10)
    decoders['W-RCQ Type 2'] = WeightedRCQDecoder(code, bc=3, bv=8, quantizer_params=qp,
                                                  weight_sharing_type=2, max_iterations=10)
    return decoders

"""N > 1 host logic on CPU: two gloo ranks run the Monte-Carlo loop with a synthetic per-frame error
model (a pure function of the global frame index, like the device Philox noise) and must reproduce the
reference's SEQUENTIAL stop rule (simulation_framework.py:110) exactly, independent of the rank count."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def frame_model(idx: np.ndarray):
    """(bit_errors, iterations) of global frame idx -- deterministic, ~7 % frame errors."""
    h = (idx.astype(np.uint64) * np.uint64(2654435761) + np.uint64(12345)) % np.uint64(1000)
    err = h < 70
    be = np.where(err, 1 + (h % np.uint64(9)), 0).astype(np.int32)
    it = (1 + (h % np.uint64(10))).astype(np.int32)
    return be, it


def sequential_reference(max_frames, max_errors, n):
    """The reference's loop, one frame at a time (simulation_framework.py:110-136)."""
    fe = be = it = nf = 0
    while nf < max_frames and fe < max_errors:
        b, i = frame_model(np.array([nf]))
        if b[0] > 0:
            fe += 1
            be += int(b[0])
        it += int(i[0])
        nf += 1
    return fe / nf, be / (nf * n), it / nf, nf, fe


def _make_sim():
    sys.path.insert(0, ROOT)
    import ldpc_b200  # noqa: F401
    from ldpc_b200.simulation_framework import LDPSimulator, SimulationConfig

    class FakeDeviceSimulator(LDPSimulator):
        def _device(self):
            return torch.device("cpu")

        def _make_round_runner(self, decoder, code, snr_db, device):
            def run(count, frame0, counters, fbe, fit):
                b, i = frame_model(np.arange(frame0, frame0 + count))
                fbe[:count] = torch.from_numpy(b)
                fit[:count] = torch.from_numpy(i)
                counters += torch.tensor([int((b > 0).sum()), int(b.sum()), int(i.sum()), count])
            return run

    return FakeDeviceSimulator, SimulationConfig


class _Code:
    n = 100


def _worker(rank, world, port, cases, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Sim, Cfg = _make_sim()
    res = []
    for batch, max_frames, max_errors in cases:
        sim = Sim(Cfg(batch_frames=batch, save_results=False))
        fer, ber, avg, _, nf, fe = sim.simulate_single_snr(None, _Code, 2.0, max_frames, max_errors)
        res.append((fer, ber, avg, nf, fe))
    if rank == 0:
        out.put(res)
    dist.barrier()
    dist.destroy_process_group()


# the last three are large enough for the adaptive round schedule (pilot round, rounds sized from the error rate)
CASES = [(64, 10000, 25), (37, 500, 1000), (128, 1000, 3), (16, 50, 1), (4096, 30000, 400), (8192, 6000, 5), (2048, 9000, 10 ** 6)]


def test_single_process_matches_sequential_rule():
    Sim, Cfg = _make_sim()
    for batch, max_frames, max_errors in CASES:
        sim = Sim(Cfg(batch_frames=batch, save_results=False))
        fer, ber, avg, _, nf, fe = sim.simulate_single_snr(None, _Code, 2.0, max_frames, max_errors)
        assert (fer, ber, avg, nf, fe) == sequential_reference(max_frames, max_errors, _Code.n)


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ranks_match_sequential_rule(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, CASES, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for (batch, max_frames, max_errors), got in zip(CASES, res):
        assert got == sequential_reference(max_frames, max_errors, _Code.n), (batch, max_frames, max_errors)


def test_adaptive_rounds_decode_fewer_frames_for_the_same_answer():
    """Round sizes follow the error rate seen so far (simulation_framework.next_round_frames); with the exact stop rule
    the reported numbers cannot depend on them."""
    Sim, Cfg = _make_sim()
    from ldpc_b200.simulation_framework import next_round_frames
    assert next_round_frames(32768, 10 ** 6, 200, 0, 0) == 1024                 # pilot: a few times max_errors
    assert next_round_frames(32768, 10 ** 6, 1000, 0, 0) == 4096
    assert next_round_frames(32768, 10 ** 6, 200, 1024, 0) == 32768             # nothing seen yet: a full round
    assert next_round_frames(32768, 10 ** 6, 200, 1024, 1024) <= 128            # every frame fails: pilot was enough
    assert next_round_frames(32768, 500, 200, 1024, 1) == 500                   # never beyond max_frames
    assert next_round_frames(65536, 10 ** 6, 200, 2048, 10, world=2) % 256 == 0   # whole warps on every rank
    assert next_round_frames(32768, 10 ** 6, 200, 0, 0, fer_hint=1.0) == 384     # previous SNR point: every frame failed
    assert next_round_frames(32768, 10 ** 6, 200, 0, 0, fer_hint=1e-4) == 32768
    assert next_round_frames(32768, 10 ** 6, 200, 0, 0, fer_hint=0.0) == 32768
    decoded = {}
    for adaptive in (False, True):
        class Counting(Sim):
            def _make_round_runner(self, *a):
                run = super()._make_round_runner(*a)

                def counted(count, *rest):
                    decoded[adaptive] = decoded.get(adaptive, 0) + count
                    run(count, *rest)
                return counted
        sim = Counting(Cfg(batch_frames=8192, save_results=False, adaptive_rounds=adaptive))
        for snr in (2.0, 2.5):      # the second point starts from the first one's error rate instead of a pilot
            out = sim.simulate_single_snr(None, _Code, snr, 100000, 40)
            assert (out[0], out[1], out[2], out[4], out[5]) == sequential_reference(100000, 40, _Code.n)
    assert decoded[True] < decoded[False] // 4, decoded

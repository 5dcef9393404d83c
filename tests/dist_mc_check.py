"""Multi-GPU Monte-Carlo check, run under torchrun (one rank per GPU, NCCL):
every rank runs the sharded LDPSimulator loop; rank 0 also runs the same sweep points alone on its GPU
(world-size-1 semantics, via a private simulator that ignores torch.distributed) and the results must be
IDENTICAL (Philox noise is keyed by the global frame index; the stop rule is sequential in frame order).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dist_mc_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ldpc_b200 as L
    from ldpc_b200 import simulation_framework as sf

    code = L.codes.dvbs2_shaped(max_iterations=10, scale=20)
    torch.manual_seed(0)
    dec = L.Neural2DMinSumDecoder(code, 2, 10)
    with torch.no_grad():
        dec._beta_table.fill_(0.8)
        dec._alpha_table.fill_(1.0)
    cases = [(1.8, 4000, 40, 256), (2.2, 3000, 1000, 500), (1.5, 900, 7, 128)]
    got = []
    for snr, max_frames, max_errors, batch in cases:
        sim = sf.LDPSimulator(sf.SimulationConfig(batch_frames=batch, seed=3, save_results=False))
        r = sim.simulate_single_snr(dec, code, snr, max_frames, max_errors)
        got.append((r[0], r[1], r[2], r[4], r[5]))
    dist.barrier()
    if rank == 0:
        saved = sf._dist
        sf._dist = lambda: (None, 0, 1)          # single-GPU semantics on rank 0's device
        try:
            for (snr, max_frames, max_errors, batch), g in zip(cases, got):
                sim = sf.LDPSimulator(sf.SimulationConfig(batch_frames=batch * 3, seed=3, save_results=False))
                r = sim.simulate_single_snr(dec, code, snr, max_frames, max_errors)
                want = (r[0], r[1], r[2], r[4], r[5])
                assert g == want, (g, want)
                print(f"snr {snr}: world {dist.get_world_size()} == world 1: FER {g[0]:.4f} frames {g[3]} errors {g[4]}")
        finally:
            sf._dist = saved
        print("DIST_MC_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

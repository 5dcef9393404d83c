"""Is the stall that a starting nvidia-smi causes host-side or GPU-side?  Per-step device times of the headline decode
with and without checkpoints (early_stop off: a whole call is enqueued without waiting for the GPU) while nvidia-smi
processes are started in the background.   python tools/stall_probe.py"""
import subprocess, sys, threading, time
import torch
sys.path.insert(0, ".")
import bench
import ldpc_b200 as L
from ldpc_b200.engine import Engine

code = bench.make_code(L, "dvbs2")
dec = bench.build_decoder(L, code, "n2d2")
b, a = dec._tables()
llr = L.awgn_llr(code.n, 65536, 2.0, seed=1234, llr_sign=-1)
stop = False


def spawner():
    while not stop:
        subprocess.run(["nvidia-smi", "--query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap",
                        "--format=csv,noheader"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        time.sleep(0.4)


for early in (True, False):
    eng = Engine(code.graph, max_iterations=10, early_stop=early, beta=b, beta_index=dec._beta_index, alpha=a,
                 alpha_index=dec._alpha_index, device=0)
    for _ in range(3):
        eng.decode_device(llr)
    torch.cuda.synchronize()
    for noisy in (False, True):
        stop = False
        th = threading.Thread(target=spawner) if noisy else None
        if th:
            th.start()
            time.sleep(0.5)
        steps = 16
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for k in range(steps):
            eng.decode_device(llr)
            ev[k + 1].record()
        torch.cuda.synchronize()
        stop = True
        if th:
            th.join()
        print("early_stop", early, "nvidia-smi being started" if noisy else "quiet", [round(ev[k].elapsed_time(ev[k + 1]), 1) for k in range(steps)], flush=True)
    eng.close()

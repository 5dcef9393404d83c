"""Per-rank, per-step timing of the C1 bench leg ((7,4) Basic f64, one on-chip launch per step) under torchrun:
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/c1_diag.py"""
import json, os, sys, time
import torch
sys.path.insert(0, ".")
import bench

rig = bench.Rig(None)
import ldpc_b200 as L
B = 1 << 22
code = bench.make_code(L, "h74")
dec = bench.build_decoder(L, code, "basic")
eng = dec._engine(rig.local_rank)
eng.reserve(B)
llr = L.awgn_llr(code.n, B, 2.0, seed=1234, frame0=rig.rank * B, llr_sign=-1, device=rig.local_rank).double()
torch.cuda.synchronize()
for _ in range(3):
    out = eng.decode_device(llr)
rig.barrier()
rows = []
for rep in range(3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    host = []
    ev[0].record()
    for k in range(10):
        t = time.perf_counter()
        out = eng.decode_device(llr)
        host.append(round((time.perf_counter() - t) * 1e3, 3))
        ev[k + 1].record()
    torch.cuda.synchronize()
    rows.append({"rank": rig.rank, "rep": rep, "dev_ms": [round(ev[k].elapsed_time(ev[k + 1]), 3) for k in range(10)], "host_ms": host})
    rig.barrier()
prof = eng.profile_read()
for r in range(rig.world):
    if r == rig.rank:
        for row in rows:
            print(json.dumps(row), flush=True)
        print(json.dumps({"rank": rig.rank, "small_decodes": prof["small_decodes"], "launches": prof["launches"]}), flush=True)
    rig.barrier()
if rig.world > 1:
    rig.dist.destroy_process_group()

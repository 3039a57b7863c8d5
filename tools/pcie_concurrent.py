#!/usr/bin/env python
"""Host-side limit of the end-to-end arm: N ranks (one per GPU) copy a pinned 403 MB buffer to their GPU and another one back
at the same time, like pamg_smooth_host does every step.  Prints per-rank and aggregate GB/s, so that the e2e numbers of
bench.py at N = 1, 2, 4, 8 can be read against what the box's PCIe / host-memory path delivers when all GPUs copy at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_concurrent.py

(torch is plumbing here: pinned memory, streams, a gloo barrier.)"""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group(backend="gloo")
    torch.cuda.set_device(local)
    n = 50331648                                           # doubles of one c5 field (403 MB)
    h_in = torch.empty(n, dtype=torch.float64).pin_memory(); h_in.fill_(1.0)
    h_out = torch.empty(n, dtype=torch.float64).pin_memory()
    d_a = torch.empty(n, dtype=torch.float64, device="cuda"); d_b = torch.zeros(n, dtype=torch.float64, device="cuda")
    up, down = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    for mode in ("h2d", "d2h", "both"):
        for rep in range(2):                               # first pass warms up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            up.wait_stream(torch.cuda.current_stream()); down.wait_stream(torch.cuda.current_stream())
            for _ in range(5):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(up):
                        d_a.copy_(h_in, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(down):
                        h_out.copy_(d_b, non_blocking=True)
            torch.cuda.current_stream().wait_stream(up); torch.cuda.current_stream().wait_stream(down)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
        nbytes = 8 * n * (2 if mode == "both" else 1)
        out[mode] = nbytes / (ms * 1e-3) / 1e9
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, out)
    else:
        allr = [out]
    if rank == 0:
        print(json.dumps({"ranks": world, "per_rank_GBps": {k: [round(r[k], 1) for r in allr] for k in out},
                          "aggregate_GBps": {k: round(sum(r[k] for r in allr), 1) for k in out}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

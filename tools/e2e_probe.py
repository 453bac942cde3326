#!/usr/bin/env python
"""What bounds the end-to-end path when several ranks share a host?  (builder tool; torchrun --nproc-per-node N)

Per rank, all ranks concurrently (barrier before every leg): pinned H2D alone; H2D + D2H on two streams; the
bdx_submit_pinned / bdx_fetch_view pipeline at several batch sizes, with and without the host reading the results."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from bdx_b200 import capi  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def agg(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    out = [t.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(out, t)
    return [round(float(v.item()), 1) for v in out]


n, L = 10_000_000, 150
cfg = bench.make_config()
config = capi.Config(cfg)
gen = capi.Stream(config, device=local, max_reads=0, max_bytes=0)
d_seq = torch.empty(n * L, dtype=torch.uint8, device="cuda")
d_off = torch.empty(n + 1, dtype=torch.int32, device="cuda")
gen.synth_device(bench.synth_spec(rank * n), n, d_seq.data_ptr(), d_off.data_ptr())
gen.sync()
h_seq = torch.empty(n * L, dtype=torch.uint8, pin_memory=True)
h_seq.copy_(d_seq)
h_back = torch.empty(200_000_000, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize()
seq_np = h_seq.numpy()
res = {}

# 1. H2D alone, 2. H2D + D2H
chunk = 75_000_000
d_buf = torch.empty(chunk, dtype=torch.uint8, device="cuda")
d_out = torch.empty(10_000_000, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, with_d2h in (("h2d_only_gbs", False), ("h2d_with_d2h_gbs", True)):
    barrier()
    t0 = time.perf_counter()
    for k in range(20):
        with torch.cuda.stream(s1):
            d_buf.copy_(h_seq[k * chunk:(k + 1) * chunk], non_blocking=True)
        if with_d2h:
            with torch.cuda.stream(s2):
                h_back[k * 10_000_000:(k + 1) * 10_000_000].copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    res[name] = agg(20 * chunk / (time.perf_counter() - t0) / 1e9)

# 3. the pipeline
for B, read_results in ((250_000, True), (500_000, True), (500_000, False), (1_000_000, True), (2_500_000, True)):
    st = capi.Stream(config, device=local, max_reads=B, max_bytes=B * L)
    off_np = (np.arange(B + 1, dtype=np.int32) * L)
    h_off = torch.from_numpy(off_np).pin_memory()
    off_p = h_off.numpy()
    nb = n // B

    def run(steps):
        q, m = 0, 0
        for _ in range(steps):
            for k in range(nb):
                st.submit(seq_np[k * B * L:(k + 1) * B * L], off_p, tag=k, pinned=True)
                q += 1
                if q == 4:
                    _, r = st.fetch(copy=False)
                    if read_results:
                        m += int(np.count_nonzero(r["bc1"]))
                    q -= 1
        while q:
            _, r = st.fetch(copy=False)
            q -= 1
        return m

    run(1)
    barrier()
    t0 = time.perf_counter()
    run(3)
    torch.cuda.synchronize()
    res[f"e2e_B{B}_{'read' if read_results else 'noread'}_Mreads"] = agg(3 * n / (time.perf_counter() - t0) / 1e6)
    st.close()
if rank == 0:
    res["cpus"] = os.cpu_count()
    print(json.dumps(res))

"""Throughput of the reference's default 4000-read chunks through several host workers (one stream each,
the reference's worker model): python tools/bench_workers.py on a B200."""
import sys, time, threading
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch, bench
from bdx_b200 import capi
cfg = bench.make_config(); config = capi.Config(cfg)
B, nb = 4000, 250
st0 = capi.Stream(config, device=0, max_reads=0, max_bytes=0)
n = B * nb
d_seq = torch.empty(n * 150, dtype=torch.uint8, device="cuda"); d_off = torch.empty(n + 1, dtype=torch.int32, device="cuda")
st0.synth_device(bench.synth_spec(0), n, d_seq.data_ptr(), d_off.data_ptr()); st0.sync()
h_seq = torch.empty(n * 150, dtype=torch.uint8, pin_memory=True); h_seq.copy_(d_seq)
h_off = torch.empty(B + 1, dtype=torch.int32, pin_memory=True); h_off.copy_(d_off[:B + 1]); torch.cuda.synchronize()
seq_np, off_np = h_seq.numpy(), h_off.numpy()
def work(st, reps):
    q = 0
    for _ in range(reps):
        for k in range(nb):
            st.submit(seq_np[k * B * 150:(k + 1) * B * 150], off_np, tag=k, pinned=True); q += 1
            if q == 4: st.fetch(copy=False); q -= 1
    while q: st.fetch(copy=False); q -= 1
for W in (1, 2, 4, 8, 16):
    sts = [capi.Stream(config, device=0, max_reads=B, max_bytes=B * 150) for _ in range(W)]
    for reps in (1, 3):
        ths = [threading.Thread(target=work, args=(s, reps)) for s in sts]
        t0 = time.perf_counter(); [t.start() for t in ths]; [t.join() for t in ths]; dt = time.perf_counter() - t0
    print(W, "workers x 4000-read chunks:", W * 3 * n / dt / 1e6, "M reads/s", flush=True)
    for s in sts: s.close()

#!/usr/bin/env python
"""Throughput of the device FASTQ block demultiplexer (bdx_demux_block, SURVEY.md 8f-1 + 8f-3).

Workload: config 2's reads (150 bp, 96 barcodes) and config 4's (trim 5' + adapter trim 3') wrapped
into FASTQ records with a fixed-width header and a quality line -- 316 bytes of text per record.
Reports, per block of --reads records: CUDA-event milliseconds of every stage, records/s and GB/s of
FASTQ text with the block resident in HBM (BDX_DEMUX_DEVICE_IO) and from pinned host memory, the HBM
roofline fraction of the non-classification stages (algorithmic bytes = text in + records out +
per-record tables), and a byte-for-byte check of a sample against the host mirror of writer_task.
Not the headline bench (bench.py); one JSON line per configuration.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

READ_LEN = 150
SEED = 0x42444D58
STAGES = ["h2d", "index_records", "pack", "classify", "keys_sort", "offsets_buckets", "copy", "d2h"]


def fastq_text(seq: np.ndarray, n: int) -> np.ndarray:
    """n fixed-width records: '@r' + 10 digits, sequence, '+', quality."""
    hdr = 12
    rec = hdr + 1 + READ_LEN + 1 + 1 + 1 + READ_LEN + 1
    a = np.empty((n, rec), dtype=np.uint8)
    a[:, 0] = ord("@"); a[:, 1] = ord("r")
    idx = np.arange(n, dtype=np.int64)
    for d in range(10):
        a[:, 2 + 9 - d] = (idx // 10 ** d % 10 + 48).astype(np.uint8)
    a[:, hdr] = 10
    a[:, hdr + 1:hdr + 1 + READ_LEN] = seq.reshape(n, READ_LEN)
    o = hdr + 1 + READ_LEN
    a[:, o] = 10; a[:, o + 1] = ord("+"); a[:, o + 2] = 10
    a[:, o + 3:o + 3 + READ_LEN] = (33 + (idx[:, None] + np.arange(READ_LEN)[None, :]) % 41).astype(np.uint8)
    a[:, o + 3 + READ_LEN] = 10
    return a.reshape(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", type=int, default=20_000)
    ap.add_argument("--workers", type=int, default=3)
    args = ap.parse_args()
    import torch
    import bdx_b200 as bdx
    from bdx_b200 import capi
    from bdx_b200.demux import Chunk, Writer
    import bench
    import bench_configs

    torch.cuda.set_device(0)
    n = args.reads
    hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    cfg4, sp4, name4 = bench_configs.configs()["4"]
    jobs = [("config2 reads as FASTQ: 96 barcodes, no trimming", bench.make_config(),
             dict(start_lo=1, start_hi=120, set2_mode=0, end_lo=0, end_hi=0)),
            ("config4 reads as FASTQ: trim 5' barcode + 3' adapter", cfg4, sp4)]
    for name, cfg, sp in jobs:
        config = capi.Config(cfg)
        st = capi.Stream(config, device=0, max_reads=0, max_bytes=0)
        d_seq = torch.empty(n * READ_LEN, dtype=torch.uint8, device="cuda")
        d_off = torch.empty(n + 1, dtype=torch.int32, device="cuda")
        spec = capi.SynthSpec(seed=SEED, first_read=0, read_len=READ_LEN, plant_permille=900, n_permille_x10=50, **sp)
        st.synth_device(spec, n, d_seq.data_ptr(), d_off.data_ptr())
        st.sync()
        text = fastq_text(d_seq.cpu().numpy(), n)
        h_text = torch.from_numpy(text).pin_memory()
        d_text = h_text.cuda()
        torch.cuda.synchronize()
        tb = text.size

        def run(dev):
            src = (d_text.data_ptr(), tb) if dev else h_text.numpy()
            mode = capi.DEMUX_SINGLE | (capi.DEMUX_DEVICE_IO if dev else 0)
            acc = np.zeros(8)
            out = None
            for it in range(args.steps + 2):
                out = st.demux_block(src, None, final_block=1, mode=mode)
                if it >= 2:
                    acc += np.array(st.demux_stage_ms())
            return acc / args.steps, out

        ms_dev, out = run(True)
        ms_host, out = run(False)

        # host workers, one stream each (the reference's worker model, core.jl:454-466): the synchronous
        # calls of different workers overlap H2D, kernels and D2H of consecutive blocks
        import threading
        import time
        W, blocks_per_worker = args.workers, 4
        sub_n = n // 4
        sub_tb = sub_n * (tb // n)
        workers = [capi.Stream(config, device=0, max_reads=0, max_bytes=0) for _ in range(W)]
        h_np = h_text.numpy()

        def work(wi, reps):
            for k in range(reps):
                blk = (wi + k) % 4
                o = workers[wi].demux_block(h_np[blk * sub_tb:(blk + 1) * sub_tb], None, final_block=1, mode=capi.DEMUX_SINGLE)
                assert o.n_records == sub_n

        for reps in (1, blocks_per_worker):       # warm-up, then timed
            ths = [threading.Thread(target=work, args=(wi, reps)) for wi in range(W)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            dt = time.perf_counter() - t0
        mt_rate = W * blocks_per_worker * sub_n / dt
        for w_ in workers:
            w_.close()
        buckets, o1, _, res = st.demux_views(out)
        # sample check against the host mirror of writer_task (same classification results)
        k = min(args.check, n)
        rec = tb // n
        ch = Chunk(1)
        for i in range(k):
            r = text[i * rec:(i + 1) * rec].tobytes().split(b"\n")
            ch.headers.append(r[0]); ch.seqs.append(r[1]); ch.pluses.append(r[2]); ch.quals.append(r[3])
        import tempfile
        ok = True
        with tempfile.TemporaryDirectory() as td:
            w = Writer(td, "p", "", cfg)
            w.write_chunk(ch, res[:k])
            w.close()
            sub = st.demux_block(text[:k * rec], None, final_block=1, mode=capi.DEMUX_SINGLE)
            sb, so1, _, _ = st.demux_views(sub)
            for b in sb:
                fn = "p." + bdx.output_filename(cfg, int(b["status"]), int(b["bc1"]), int(b["bc2"]))
                want = open(os.path.join(td, fn), "rb").read()
                ok = ok and want == so1[b["offset1"]:b["offset1"] + b["length1"]].tobytes()
            ok = ok and len(sb) == len(os.listdir(td))
        out_bytes = int(out.out1_len)
        io_ms = float(ms_dev[1] + ms_dev[2] + ms_dev[4] + ms_dev[5] + ms_dev[6])
        # algorithmic bytes of the non-classification stages: text read twice (count, index), once more for
        # pack + copy; records written once; 32 B record table written + read 3x; sort 3 x 8 B in+out
        alg = 4 * tb + out_bytes + n * (32 * 4 + READ_LEN + 48)
        line = {"config": name, "records": n, "text_bytes": tb, "out_bytes": out_bytes, "buckets": int(out.n_buckets),
                "stage_ms_device_resident": dict(zip(STAGES, [round(float(x), 3) for x in ms_dev])),
                "stage_ms_from_pinned_host": dict(zip(STAGES, [round(float(x), 3) for x in ms_host])),
                "records_per_sec_device_resident": n / (float(ms_dev[1:7].sum()) * 1e-3),
                "text_gbs_device_resident": tb / (float(ms_dev[1:7].sum()) * 1e-3) / 1e9,
                "records_per_sec_from_host": n / (float(ms_host.sum()) * 1e-3),
                "text_gbs_from_host": tb / (float(ms_host.sum()) * 1e-3) / 1e9,
                "host_workers": {"workers": W, "records_per_block": sub_n, "records_per_sec": mt_rate,
                                 "text_gbs": mt_rate * (tb // n) / 1e9},
                "io_stages": {"ms": io_ms, "algorithmic_bytes": alg, "achieved_gbs": alg / (io_ms * 1e-3) / 1e9,
                              "hbm_peak_gbs": hbm_peak, "frac": alg / (io_ms * 1e-3) / 1e9 / hbm_peak},
                "sample_checked_records": k, "sample_byte_exact": bool(ok)}
        print(json.dumps(line), flush=True)
        st.close()
        del d_seq, d_off, d_text, h_text


if __name__ == "__main__":
    main()

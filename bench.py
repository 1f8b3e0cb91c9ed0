#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark of the B200 fingerprint-retrieval hot path.

Headline (BASELINE.json `metric`): search queries/s at the full-scale ~56 M-fingerprint database,
where a "query" is one (test id, sequence length) evaluation -- the unit of the reference's own
"ms/query" (eval/eval_faiss.py:246-258).  One step = the whole evaluation job of the reference CLI:
2,000 query sequences (the ICASSP test ids) x sequence lengths 1 3 5 9 11 19 (12,000 queries),
k_probe 20, against [dummy_db; db] = 56,000,000 + 29,500 unit-norm 128-d rows generated on the device.
With N GPUs the same database is row-sharded (strong scaling) and the per-rank top-k / candidate
scores are combined with an NCCL all-gather and a max all-reduce.

  value      device-resident: queries already in HBM, CUDA events around the kernels + collectives
  e2e        the public host API: queries H2D from pinned memory, predictions D2H, every step
  roofline   the flat scan kernel: rows_local x 256 B (and 2 x rows x 128 x query rows flop) per launch /
             CUDA-event launch time vs the measured HBM copy bandwidth and sustained bf16 rate
             (MEASURED_PEAKS.json); the nearer bound leads, the other is under "other_bound"
  fingerprint  secondary metric of the same path: log-mel + encoder segments/s (tensor roofline)
  cpu_baseline the CPU oracle (numpy/BLAS port of the reference + faiss-flat semantics) on a bounded
             sample, rank 0 at N=1 only

`--impl reference` times that CPU port alone with all host threads (no GPU work at all).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DUMMY_FULL = 56_000_000
N_DB = 29_500
SEQ_LENS = [1, 3, 5, 9, 11, 19]
K_PROBE = 20
FP_SEGS_PER_STEP = 4000          # 32 groups of TS_BATCH_SZ = 125 = one full encoder pass (ENC_CHUNK_MAX)
FP_E2E_SEGS = 8000               # host-API step: what one generate.py call hands over (64 batches)
FLOPS_PER_SEGMENT = 607_199_232  # model/arch.py (conv + div-enc)
SAMPLE_ROWS = 250_000            # CPU baseline: database sample
SAMPLE_IDS = 30                  # CPU baseline: test ids per step


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def _test_ids():
    return np.load(os.path.join(ROOT, "neural-audio-fp_b200", "eval", "test_ids_icassp2021.npy")).astype(np.int64)


def _make_queries(db, seed=12):
    from nafp_b200 import synth
    return synth.synth_fp_queries(db, seed)


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu = gpu
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
            rows = [r.strip().split(", ") for r in open(self.path) if r.strip()]
            rows = [r for r in rows if len(r) >= 8]
            sm = sorted(float(r[0]) for r in rows)
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = float(rows[0][1])
                out["samples"] = len(sm)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for i, nm in enumerate(names):
                    if any(r[4 + i].strip().lower().startswith("active") for r in rows):
                        out["reasons"].append(nm)
        except Exception as e:  # clocks are evidence, not a reason to lose the number
            out["error"] = str(e)
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        return out


# ----------------------------------------------------------------------------------------------
# CPU port (oracle) -- cpu_baseline leg and the --impl reference arm
# ----------------------------------------------------------------------------------------------
def cpu_search_sample(sample_rows, db, query, test_ids, n_ids, steps, warmup, n_full_rows, threads):
    """Time the CPU oracle (reference hot loop eval_faiss.py:204-243 over a flat L2 index) on
    `n_ids` test ids against `sample_rows` + db rows.  Exhaustive search cost is linear in the row
    count, so queries/s at the full database = measured x (sample rows / full rows)."""
    from oracle.flat_index import FlatL2
    from oracle import seq_match

    class Fast(FlatL2):
        def search(self, q, k):
            return FlatL2.search(self, q, k, fast=True)

    idx = Fast(128)
    idx.add(sample_rows)
    idx.add(db)
    recon = idx._data()
    n_dummy = len(sample_rows)
    ids = test_ids[:n_ids]
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        seq_match.evaluate(idx, query, recon, n_dummy, ids, SEQ_LENS, K_PROBE)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    q_per_s_sample = n_ids * len(SEQ_LENS) / sec
    scale = (n_dummy + len(db)) / float(n_full_rows)
    return dict(value=q_per_s_sample * scale, unit="queries/s", cores=threads, kind="port",
                measured_queries_per_s_on_sample=q_per_s_sample, sec_per_step=sec,
                sample=f"{n_ids} of {len(test_ids)} test ids x {len(SEQ_LENS)} lengths against a "
                       f"{n_dummy + len(db):,}-row sample of the {n_full_rows:,}-row database; exhaustive search is "
                       f"linear in rows, value = measured x {scale:.5f}")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from nafp_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n_full = args.db_rows + N_DB
    sample = synth.synth_fp_db(min(SAMPLE_ROWS, args.db_rows), seed=11)
    db = synth.synth_fp_db(N_DB, seed=13)
    query = _make_queries(db)
    test_ids = _test_ids()
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    cb = cpu_search_sample(sample, db, query, test_ids, SAMPLE_IDS, steps, warm, n_full, threads)
    line = {"impl": "reference", "metric": "search_queries_per_s", "value": cb["value"], "unit": "queries/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": cb["sec_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(args, 1), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def _config(args, world):
    return {"workload": f"exact flat-L2 search + sequence rescoring: 2,000 query sequences x lengths {SEQ_LENS}, "
                        f"k_probe {K_PROBE}, vs {args.db_rows + N_DB:,}-row 128-d database "
                        f"(BASELINE configs[3]; configs[1] reported under 'mini_1M')",
            "db_rows": args.db_rows + N_DB, "n_test_ids": 2000, "seq_lens": SEQ_LENS, "k_probe": K_PROBE,
            "parallelism": f"db-row-shard x{world}" if world > 1 else "single GPU",
            "l2_policy": "inputs larger than L2 (bf16 scan copy >= 1.8 GB per GPU is streamed every pass)"}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def build_sharded_index(ctx, torch, sidx, n_dummy, dev):
    """Fill this rank's block of [dummy (seed 11); db (seed 13)] on the device, chunk by chunk."""
    from nafp_b200._lib import check, lib
    import ctypes
    lo, hi = sidx.local_rows_needed()
    chunk = 4_000_000
    buf = torch.empty((min(chunk, hi - lo), 128), dtype=torch.float32, device=dev)
    r = lo
    while r < hi:
        if r < n_dummy:
            n = min(chunk, min(hi, n_dummy) - r)
            check(lib.nafp_synth_fp_rows(ctx.h, 11, r, n, 59, 0.5, ctypes.c_void_p(buf.data_ptr())))
        else:
            n = min(chunk, hi - r)
            check(lib.nafp_synth_fp_rows(ctx.h, 13, r - n_dummy, n, 59, 0.5, ctypes.c_void_p(buf.data_ptr())))
        sidx.add_local_dev(buf.data_ptr(), n)
        r += n
    torch.cuda.synchronize(dev)
    del buf


def time_steps(torch, dev, fn, steps, warmup, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    barrier()
    return e0.elapsed_time(e1) / steps


def run_gpu(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from nafp_b200._lib import Context, check, lib
    from nafp_b200.dist import ShardedFlatIndex, TorchComm
    from nafp_b200.model import weights as W
    from nafp_b200.model.fp import FingerPrinter

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        comm = TorchComm()

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = Context.get(local_rank)
    # all torch work (copies, NCCL, CUDA events) on the stream the library launches its kernels on
    torch.cuda.set_stream(torch.cuda.ExternalStream(ctx.stream, device=dev))
    hbm_peak, tf_burst, tf_sustained, peak_src = _peaks()
    n_dummy = args.db_rows
    n_total = n_dummy + N_DB
    test_ids = _test_ids()
    n_queries = len(test_ids) * len(SEQ_LENS)

    # ---- database shard + queries
    t_build = time.time()
    sidx = ShardedFlatIndex(n_total, rank, world, max_len=max(SEQ_LENS), device=local_rank, comm=comm)
    build_sharded_index(ctx, torch, sidx, n_dummy, dev)
    dbt = torch.empty((N_DB, 128), dtype=torch.float32, device=dev)
    check(lib.nafp_synth_fp_rows(ctx.h, 13, 0, N_DB, 59, 0.5, ctypes.c_void_p(dbt.data_ptr())))
    db_host = dbt.cpu().numpy()
    query_host = _make_queries(db_host)
    t_build = time.time() - t_build

    q_pin = torch.from_numpy(query_host).pin_memory()
    ids_pin = torch.from_numpy(test_ids).pin_memory()
    sl_pin = torch.tensor(SEQ_LENS, dtype=torch.int32).pin_memory()
    q_dev, ids_dev, sl_dev = q_pin.to(dev), ids_pin.to(dev), sl_pin.to(dev)
    result = {}

    def step_resident():
        result["pid"], result["psc"] = sidx.seq_match_dev(q_dev, ids_dev, sl_dev, K_PROBE)

    def step_e2e():
        q = q_pin.to(dev, non_blocking=True)
        ids = ids_pin.to(dev, non_blocking=True)
        sl = sl_pin.to(dev, non_blocking=True)
        pid, psc = sidx.seq_match_dev(q, ids, sl, K_PROBE)
        result["pid_host"] = pid.to("cpu")
        result["psc_host"] = psc.to("cpu")

    # ---- headline: resident
    sampler = ClockSampler(local_rank)
    sampler.start()                       # nvidia-smi needs a few hundred ms to start; warm-up runs under load too
    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize(dev)
    sidx.index.profile_scans(True)
    sidx.index.last_search_stats()
    launches0 = ctx.launches
    ms_res = max_over_ranks(time_steps(torch, dev, step_resident, args.steps, 0, barrier))
    clocks = sampler.stop()
    launches = ctx.launches - launches0
    scan_ms, n_scans = sidx.index.profile_scans(False)
    stats = sidx.index.last_search_stats()
    # ---- e2e
    ms_e2e = max_over_ranks(time_steps(torch, dev, step_e2e, args.steps, max(1, args.warmup // 2), barrier))

    pid = result["pid"].cpu().numpy()
    gt = test_ids + n_dummy
    top1 = [float(100.0 * np.mean(pid[:, si, 0] == gt)) for si in range(len(SEQ_LENS))]

    # ---- roofline of the scan kernel (this rank's rows; all ranks launch the same count).
    # One launch streams the rank's bf16 rows once (rows x 256 B) and multiplies them with up to 256 query
    # rows (2 x rows x 128 x query rows flop).  With 256-row passes the two bounds are within 15 % of each
    # other on B200 (0.72 us of HBM time vs 0.65-0.78 us of tensor time per 128 rows at 1.6-1.3 GHz); the
    # line reports the bound the kernel is closer to, the other one rides along.
    rows_local = sidx.hi - sidx.lo
    scan_avg_ms = scan_ms / max(n_scans, 1)
    alg_bytes = rows_local * 256.0
    q_rows_per_launch = stats["rows"] / max(stats["passes"], 1) if stats.get("passes") else 256.0
    alg_flops = 2.0 * rows_local * 128.0 * q_rows_per_launch
    gbs = alg_bytes / (scan_avg_ms * 1e-3) / 1e9 if n_scans else 0.0
    tfs = alg_flops / (scan_avg_ms * 1e-3) / 1e12 if n_scans else 0.0
    hbm_part = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs)", "algorithmic_bytes_per_launch": alg_bytes}
    tensor_part = {"bound": "tensor", "achieved": tfs, "peak": tf_sustained, "unit": "TFLOP/s", "frac": tfs / tf_sustained,
                   "peak_source": f"{peak_src} (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)",
                   "algorithmic_flops_per_launch": alg_flops}
    first, second = (tensor_part, hbm_part) if tensor_part["frac"] >= hbm_part["frac"] else (hbm_part, tensor_part)
    # DRAM bytes of ONE launch from the committed ncu --set full capture (same kernel, same 56,029,500-row
    # single-GPU shard, 256 query rows): dram__bytes_read.sum + dram__bytes_write.sum
    traffic = 14.650915e9 + 7.794176e6 if rows_local == N_DUMMY_FULL + N_DB else None
    roofline = {"kernel": "flat_scan_kernel", **first, "traffic": traffic,
                "traffic_source": "profiles/r1_prof_scan_r1f_summary.csv" if traffic else None,
                "other_bound": second, "launches": n_scans,
                "avg_launch_ms": scan_avg_ms, "scan_share_of_step": scan_ms / max(ms_res * args.steps, 1e-9),
                "query_rows_per_launch": q_rows_per_launch}

    # ---- secondary: fingerprint generation (every rank runs its own batches, no collective)
    fp = None
    if not args.no_fp:
        m_fp = FingerPrinter(ctx).load(W.init_weights(7))
        x_dev = torch.empty((FP_SEGS_PER_STEP, 8000), dtype=torch.float32, device=dev)
        check(lib.nafp_synth_audio(ctx.h, 5, rank * FP_SEGS_PER_STEP, FP_SEGS_PER_STEP, ctypes.c_void_p(x_dev.data_ptr())))
        emb_dev = torch.empty((FP_SEGS_PER_STEP, 128), dtype=torch.float32, device=dev)
        # e2e: int16 PCM segments from pinned host memory (the / 2**15 of audio_utils.py:243-244 runs on the GPU),
        # fingerprints back to host memory
        # one e2e step = one call of generate.py's size (batches_per_call 64 x TS_BATCH_SZ 125 = two encoder passes: the
        # upload of the second runs under the kernels of the first)
        x_pin = (x_dev.clamp(-1.0, 1.0) * 32767.0).round().to(torch.int16).cpu().repeat(FP_E2E_SEGS // FP_SEGS_PER_STEP, 1).pin_memory()
        emb_host = np.empty((FP_E2E_SEGS, 128), np.float32)

        def fp_resident():
            check(lib.nafp_fingerprint(ctx.h, ctypes.c_void_p(x_dev.data_ptr()), FP_SEGS_PER_STEP, 125,
                                       ctypes.c_void_p(emb_dev.data_ptr())))

        def fp_e2e():
            check(lib.nafp_fingerprint_pcm16_host(ctx.h, ctypes.c_void_p(x_pin.data_ptr()), FP_E2E_SEGS, 125,
                                            emb_host.ctypes.data_as(ctypes.c_void_p)))

        l0 = ctx.launches
        ms_fp = max_over_ranks(time_steps(torch, dev, fp_resident, args.steps, args.warmup, barrier))
        fp_launches = (ctx.launches - l0) // (args.steps + args.warmup)
        ms_fp_e2e = max_over_ranks(time_steps(torch, dev, fp_e2e, args.steps, 1, barrier))
        segs = FP_SEGS_PER_STEP * world
        tfl = segs / (ms_fp * 1e-3) * FLOPS_PER_SEGMENT / 1e12 / world
        fp = {"metric": "fp_segments_per_s", "value": segs / (ms_fp * 1e-3), "unit": "segments/s", "ms_per_step": ms_fp,
              "segments_per_step": segs, "dtype": "fp16 operands, fp32 accumulate",
              "e2e": {"value": FP_E2E_SEGS * world / (ms_fp_e2e * 1e-3), "unit": "segments/s", "segments_per_step": FP_E2E_SEGS * world,
                      "h2d_bytes_per_step": FP_E2E_SEGS * 16000, "d2h_bytes_per_step": FP_E2E_SEGS * 512,
                      "through": "nafp_fingerprint_pcm16_host ((n, 8000) int16 rows from pinned host memory; model/generate.py's "
                                 "track-window entry point uploads half as many samples)"},
              "gpu_launches_per_step": int(fp_launches),
              "roofline": {"kernel": "conv_gemm_kernel (encoder, whole step)", "bound": "tensor", "achieved": tfl,
                           "peak": tf_sustained, "unit": "TFLOP/s", "frac": tfl / tf_sustained, "traffic": None,
                           "peak_source": f"{peak_src} (bf16_tflops_sustained, per GPU)"}}

    # ---- mini scale (BASELINE configs[1]): same job against a 1 M-row database, N = 1 only
    mini = None
    if world == 1 and not args.no_mini and n_dummy > 1_000_000:
        midx = ShardedFlatIndex(1_000_000 + N_DB, 0, 1, max_len=max(SEQ_LENS), device=local_rank)
        build_sharded_index(ctx, torch, midx, 1_000_000, dev)

        def mini_step():
            result["mini"] = midx.seq_match_dev(q_dev, ids_dev, sl_dev, K_PROBE)

        ms_mini = time_steps(torch, dev, mini_step, args.steps, args.warmup, barrier)
        mp = result["mini"][0].cpu().numpy()
        mini = {"db_rows": 1_000_000 + N_DB, "value": n_queries / (ms_mini * 1e-3), "unit": "queries/s",
                "ms_per_step": ms_mini,
                "top1_hit_rate": [float(100.0 * np.mean(mp[:, si, 0] == test_ids + 1_000_000)) for si in range(len(SEQ_LENS))]}
        del midx

    # ---- IVF-PQ at the mini scale: same job through the approximate index the reference
    # builds for index_type "ivfpq" (nlist 256, M 64, 8 bit, nprobe 40), host API incl. H2D / D2H
    def ivfpq_leg(rows_dummy):
        """Index build (k-means on a 1-in-N sample of the dummy rows, encode + decode of every row) and the timed
        evaluation job through the host API; the database is generated chunk by chunk on the device."""
        from nafp_b200.eval.utils.get_index import IVFPQ, Index
        t_ivf = time.time()
        iidx = Index(IVFPQ, 128, nlist=256, pq_m=64, pq_nbits=8, device=local_rank)
        chunk = min(4_000_000, rows_dummy)
        buf = torch.empty((chunk, 128), dtype=torch.float32, device=dev)
        check(lib.nafp_synth_fp_rows(ctx.h, 11, 0, chunk, 59, 0.5, ctypes.c_void_p(buf.data_ptr())))
        # 100,000 training rows from the first chunk (<= 256 per centroid is what faiss samples anyway)
        iidx.train(buf[:: max(1, chunk // 100_000)].contiguous().cpu().numpy())
        t_train = time.time() - t_ivf
        iidx.reserve(rows_dummy + N_DB)
        r = 0
        while r < rows_dummy:
            n = min(chunk, rows_dummy - r)
            if r:
                check(lib.nafp_synth_fp_rows(ctx.h, 11, r, n, 59, 0.5, ctypes.c_void_p(buf.data_ptr())))
            iidx.add_dev(buf.data_ptr(), n)
            r += n
        iidx.add_dev(dbt.data_ptr(), N_DB)
        iidx.nprobe = 40
        torch.cuda.synchronize(dev)
        t_add = time.time() - t_ivf - t_train
        del buf
        sl_host = np.asarray(SEQ_LENS, np.int32)

        def ivf_step():
            result["ivf"] = iidx.seq_match(query_host, test_ids, sl_host, K_PROBE)

        ms_ivf = time_steps(torch, dev, ivf_step, args.steps, args.warmup, barrier)
        ip = result["ivf"][0]
        out = {"index": "IVFPQ nlist 256, M 64, 8 bit, nprobe 40", "db_rows": rows_dummy + N_DB,
               "value": n_queries / (ms_ivf * 1e-3), "unit": "queries/s", "ms_per_step": ms_ivf, "train_s": t_train,
               "add_s": t_add,
               "through": "host API (nafp_seq_match: H2D queries, D2H predictions inside the timed region)",
               "top1_hit_rate": [float(100.0 * np.mean(ip[:, si, 0] == test_ids + rows_dummy)) for si in range(len(SEQ_LENS))]}
        del iidx
        return out

    ivf = None
    if world == 1 and not args.no_mini and not args.no_ivfpq and n_dummy >= 1_000_000:
        ivf = ivfpq_leg(1_000_000)
    # BASELINE configs[4] on one GPU (94 GB for the codes, lists and the resident reconstruction): opt-in, the
    # default run stays short
    ivf_full = None
    if world == 1 and args.ivfpq_full and n_dummy > 1_000_000:
        ivf_full = ivfpq_leg(n_dummy)
    if world > 1 and args.ivfpq_full:
        # BASELINE configs[4]: the IVF-PQ index row-sharded like the flat one (quantizers trained on rank 0 and
        # broadcast, codes / lists / reconstructions local), same all-gather + merge + max all-reduce
        from nafp_b200.eval.utils.get_index import IVFPQ
        t_ivf = time.time()
        sivf = ShardedFlatIndex(n_total, rank, world, max_len=max(SEQ_LENS), device=local_rank, comm=comm, index_type=IVFPQ)
        tr = torch.empty((min(4_000_000, n_dummy), 128), dtype=torch.float32, device=dev)
        check(lib.nafp_synth_fp_rows(ctx.h, 11, 0, tr.shape[0], 59, 0.5, ctypes.c_void_p(tr.data_ptr())))
        sivf.train(tr[:: max(1, tr.shape[0] // 100_000)].contiguous().cpu().numpy())
        del tr
        t_train = time.time() - t_ivf
        build_sharded_index(ctx, torch, sivf, n_dummy, dev)
        t_add = time.time() - t_ivf - t_train

        def sivf_step():
            result["sivf"] = sivf.seq_match_dev(q_dev, ids_dev, sl_dev, K_PROBE)

        ms_sivf = max_over_ranks(time_steps(torch, dev, sivf_step, args.steps, args.warmup, barrier))
        sp = result["sivf"][0].cpu().numpy()
        ivf_full = {"index": "IVFPQ nlist 256, M 64, 8 bit, nprobe 40, row-sharded", "db_rows": n_total,
                    "value": n_queries / (ms_sivf * 1e-3), "unit": "queries/s", "ms_per_step": ms_sivf, "train_s": t_train,
                    "add_s": t_add, "through": "device-resident queries (seq_match_dev), NCCL all-gather + max all-reduce",
                    "top1_hit_rate": [float(100.0 * np.mean(sp[:, si, 0] == test_ids + n_dummy)) for si in range(len(SEQ_LENS))]}
        del sivf

    # ---- CPU baseline (rank 0, N = 1): the oracle port on a bounded sample
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        ns = min(SAMPLE_ROWS, n_dummy)
        samp = torch.empty((ns, 128), dtype=torch.float32, device=dev)
        check(lib.nafp_synth_fp_rows(ctx.h, 11, 0, ns, 59, 0.5, ctypes.c_void_p(samp.data_ptr())))
        sample_rows = samp.cpu().numpy()
        del samp
        cpu = cpu_search_sample(sample_rows, db_host, query_host, test_ids, SAMPLE_IDS, 1, 0, n_total, threads)

    if rank == 0:
        line = {"metric": "search_queries_per_s", "value": n_queries / (ms_res * 1e-3), "unit": "queries/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16 scan, f32 re-rank", "data": "synthetic", "config": _config(args, world),
                "clocks": clocks,
                "e2e": {"value": n_queries / (ms_e2e * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(query_host.nbytes + test_ids.nbytes + 4 * len(SEQ_LENS)),
                        "d2h_bytes_per_step": int(len(test_ids) * len(SEQ_LENS) * 10 * 12)},
                "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu,
                "top1_hit_rate": dict(zip(map(str, SEQ_LENS), top1)),
                "search_stats_per_step": {k: v / args.steps for k, v in stats.items()},
                "fingerprint": fp, "mini_1M": mini, "ivfpq_1M": ivf, "ivfpq_full": ivf_full, "db_build_s": t_build}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nafp", choices=["nafp", "reference"])
    ap.add_argument("--db-rows", type=int, default=N_DUMMY_FULL, help="dummy_db rows (full scale: 56,000,000)")
    ap.add_argument("--no-fp", action="store_true")
    ap.add_argument("--no-mini", action="store_true")
    ap.add_argument("--no-ivfpq", action="store_true")
    ap.add_argument("--ivfpq-full", action="store_true", help="also time IVF-PQ on the full-size database (94 GB more)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())

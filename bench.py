#!/usr/bin/env python
"""
Headline benchmark: LSH kNN queries/s @k=10 over 10M x 256-bit ITQ codes
(BASELINE.json configs[1]: ITQ-256 hashing + LinearHashIndex Hamming top-k over
10M x 512-d fp32 descriptors, batches of 4096 queries, euclidean re-rank).

    python bench.py --gpus N --steps K --warmup W            # this framework, N GPUs
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU path (rank 0 only)

One "step" = one batch of Q queries through the whole query path
(sb_itq_hash_tc -> sb_hamming_scan_tc4 over the unique-code table -> candidate
expansion -> sb_rerank -> sb_rerank_select), i.e. LSHNearestNeighborIndex.nn
(reference smqtk_indexing/impls/nn_index/lsh.py:452-519) for Q queries.

  value   whole-job queries/s with the queries already resident in HBM, through the product's
          default path (the pipeline replayed as one CUDA graph once the batch shape repeats)
  e2e     the same through the public plugin call LSHNearestNeighborIndex.nn_batch
          with pinned HOST query buffers in and host result arrays out
  roofline  the dominant kernel (ham_filter_tc_kernel: tcgen05 +-1 dot products, packed FP4 operands by
          default, FP8 with SB_TC_SCAN_FORMAT=fp8): algorithmic flops 2*Q*U*b per step over its CUDA-event
          duration -- measured in a second, kernel-by-kernel pass of the same K steps --, against a cuBLASLt
          GEMM of the same operand format timed in this very run (2 x the measured bf16 figure and the
          tensor-memory read floor beside it); `rerank` and `single_query_scan` carry the HBM-bound
          kernels' GB/s
  cpu_baseline  the reference's algorithm (oracle/ref_port.py, literal port) on
          this box's host cores, bounded sample
  parity_checked  8 queries of the timed batch re-derived with plain torch (float64 hash, byte-LUT
          popcount top-k, CSR expansion, float64 distances) outside the timed region, at every N

N > 1: one process per GPU (torchrun).  Descriptors are row-sharded; the 320 MB code table is
replicated; the scan is partitioned over the QUERIES (each rank scans the whole table for its slice of
the batch) and every rank finishes its slice alone, reading candidate rows from the owners' HBM over
NVLink (CUDA IPC); one NCCL all-gather assembles the results.  `config.parallelism` states what ran;
`--scan-partition rows` runs the row-partitioned scan (per-rank table slices, all-gather of the local
top-k keys + merge) for comparison.  Total work is fixed (strong scaling).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "LSH kNN queries/s @k=10, 10M×256-bit codes, 1/8 B200; Hamming scan GB/s"
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--bits", type=int, default=256)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--fit-rows", type=int, default=100_000)
    ap.add_argument("--fit-iters", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scan-partition", default="auto", choices=["auto", "rows", "queries"],
                    help="N > 1: how the Hamming scan is split over the ranks")
    ap.add_argument("--rerank", default="auto", choices=["auto", "peer", "allreduce"],
                    help="N > 1: candidate rows read from the owners' HBM (peer) or owner-computes + all-reduce")
    ap.add_argument("--no-graph", action="store_true", help="never replay the pipeline as a CUDA graph")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="BASELINE.json configs: c2 = configs[1] (the headline; default), c3 = configs[2] exact flat L2 "
                         "kNN k=100 over 12.5M x 128-d rows PER GPU (100M on 8), c4 = configs[3] LSH + histogram "
                         "intersection re-rank on 1M x 4096-d, k=50, c5 = configs[4] ItqFunctor.fit + build on 50M x 256-d")
    ap.add_argument("--c5-rows", type=int, default=50_000_000)
    ap.add_argument("--cpu-budget-s", type=float, default=20.0, help="CPU-baseline sample size (seconds of CPU work)")
    ap.add_argument("--ref-budget-s", type=float, default=100.0, help="--impl reference: seconds of CPU work in total")
    return ap.parse_args()


# --------------------------------------------------------------------------- helpers
def fp8_peak():
    """Dense FP8 tensor peak in TFLOP/s: 2 x the measured bf16 cuBLAS figure (burst: the kernel is timed
    alone between L2 flushes), else 2 x the profiling guide's bf16 fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return 2.0 * float(json.load(f)["bf16_tflops"]), "2 x MEASURED_PEAKS.json bf16_tflops (burst)"
    except Exception:
        return 2.0 * 1590.0, "2 x B200_PROFILING.md bf16 fallback (1.59 PFLOP/s)"


def measure_fp8_gemm(dev):
    """cuBLASLt FP8 (E4M3 x E4M3 -> BF16) GEMM, 8192^3, best of 10 with CUDA events: the tensor-pipe
    denominator measured in THIS run on THIS box (MEASURED_PEAKS.json has no FP8 figure).  None when the
    library call is unavailable."""
    import torch
    try:
        n = 8192
        a = torch.randn((n, n), device=dev).to(torch.float8_e4m3fn)
        b = torch.randn((n, n), device=dev).to(torch.float8_e4m3fn).t()          # column-major B as cuBLASLt wants
        one = torch.ones((), device=dev)
        for _ in range(3):
            torch._scaled_mm(a, b, scale_a=one, scale_b=one, out_dtype=torch.bfloat16)
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._scaled_mm(a, b, scale_a=one, scale_b=one, out_dtype=torch.bfloat16)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception as e:                   # pragma: no cover - depends on the torch build
        sys.stderr.write("fp8 GEMM measurement unavailable: %r\n" % (e,))
        return None


def measure_fp4_gemm(dev):
    """cuBLASLt block-scaled FP4 (E2M1 x E2M1, one E4M3 scale per 16 elements -> BF16: the FP4 GEMM this torch
    build exposes; it issues at the same tensor-pipe rate as the kernel's kind::mxf4) GEMM, 8192^3, best of 10
    with CUDA events: the tensor-pipe denominator of the packed-FP4 scan.  None when the library call is
    unavailable (the caller then uses 2 x the measured FP8 figure and says so)."""
    import torch
    try:
        n = 8192
        a = torch.randint(0, 256, (n, n // 2), device=dev, dtype=torch.uint8).view(torch.float4_e2m1fn_x2)
        b = torch.randint(0, 256, (n, n // 2), device=dev, dtype=torch.uint8).view(torch.float4_e2m1fn_x2).t()
        sc = torch.full((n * (n // 16),), 0x38, device=dev, dtype=torch.uint8).view(torch.float8_e4m3fn)   # 1.0 everywhere
        for _ in range(3):
            torch._scaled_mm(a, b, scale_a=sc, scale_b=sc, out_dtype=torch.bfloat16)
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._scaled_mm(a, b, scale_a=sc, scale_b=sc, out_dtype=torch.bfloat16)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception as e:                   # pragma: no cover - depends on the torch build
        sys.stderr.write("fp4 GEMM measurement unavailable: %r\n" % (e,))
        return None


TC4_QUERIES_PER_COLUMN = 3     # hamming_tc4.cu: queries per FP32 accumulator column


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic(name="scan_traffic.json"):
    """DRAM bytes of the scan kernel from the committed ncu capture (profiles/<name>), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region.  NVML from a thread every 5 ms
    (the timed region of a fast path lasts ~0.1 s: `nvidia-smi -lms` would catch one sample);
    falls back to polling nvidia-smi when NVML cannot be loaded."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.thread = None
        self.stop_flag = False
        self.sm, self.mx, self.reasons = [], [], set()
        self.errors, self.last_error, self.sample_now = 0, None, None

    def _make_sampler(self, nv, h):
        """The NVML queries of one sample, bound to a device handle (built in the calling thread so that
        `sample_now` exists before the timed region starts, whatever the polling thread's scheduling)."""
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown",
                                       getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0)),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown",
                                               getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0)),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown",
                                               getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0)),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap",
                                        getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0))}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons",
                              getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None))
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception as e:
            sys.stderr.write("clock sampler: max clock query failed: %r\n" % (e,))
            mx = float("nan")
        lock = threading.Lock()

        def sample():
            with lock:
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    self.mx.append(mx)
                    if get_reasons is not None:
                        r = int(get_reasons(h))
                        for nm, bit in bits.items():
                            if bit and (r & bit):
                                self.reasons.add(nm)
                except Exception as e:
                    self.errors += 1
                    self.last_error = repr(e)

        return sample

    def _nvml_loop(self):
        while not self.stop_flag:
            self.sample_now()
            time.sleep(0.004)

    def sample_between_steps(self):
        """Called by the timed loop right after a step's kernels were launched (the events bracket the
        step, not this call): a guaranteed sample under load even when the polling thread is starved."""
        if self.sample_now is not None:
            self.sample_now()

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            gid = str(self.gpu_index)
            h = nv.nvmlDeviceGetHandleByUUID(gid) if gid.startswith("GPU-") else nv.nvmlDeviceGetHandleByIndex(int(gid))
            self.sample_now = self._make_sampler(nv, h)
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            if self.sample_now is not None:
                self.sample_now()            # at least one sample taken while the last timed kernels drain
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.errors:
                sys.stderr.write("clock sampler: %d failed NVML samples (%s)\n" % (self.errors, self.last_error))
        elif self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in out.splitlines():
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    self.sm.append(float(parts[0]))
                    self.mx.append(float(parts[1]))
                except ValueError:
                    continue
                for nm, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        # the first samples can precede the first kernel: keep the busiest 80 %
        sm_sorted = sorted(self.sm)
        busy = sm_sorted[len(sm_sorted) // 5:] if len(sm_sorted) >= 5 else sm_sorted
        return {"sm_mhz": statistics.median(busy) if busy else None,
                "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.thread is not None else "nvidia-smi"}


# --------------------------------------------------------------------------- CPU legs (oracle = checker / baseline only)
def literal_cpu_leg(args, steps, warmup, budget_s):
    """The reference's query path (oracle/ref_port.py) on host cores, over a bounded
    sample of the workload: Q_s = 1 query per step against the first U_s codes of a
    synthetic table of the same shape; queries/s is scaled linearly from U_s to the
    full table (the scan is a flat loop over the code set, linear.py:235-238)."""
    import numpy as np
    import ref_port as RP

    rng = np.random.RandomState(0)
    D, b, U = args.dim, args.bits, args.rows
    # random orthonormal rotation: model quality does not change the cost of the query path
    rot = np.linalg.qr(rng.randn(D, D))[0][:, :b]
    mean = np.full(D, 0.5)
    qs = rng.rand(max(steps + warmup, 1), D).astype(np.float32)

    # calibration: per-code cost of the literal scan on 20k codes
    def make(n_codes):
        x = rng.rand(n_codes, D).astype(np.float32)
        z = (x.astype(np.float64) - mean) @ rot
        words = np.packbits(z >= 0, axis=1)            # big-endian bytes == integer value
        w = words.shape[1]
        raw = words.tobytes()
        ints = [int.from_bytes(raw[i:i + w], "big") for i in range(0, len(raw), w)]
        idx = RP.LiteralLshIndex(mean, rot, "euclidean")
        idx.adopt_codes(range(n_codes), x, ints)
        return idx

    cal = make(20_000)
    t0 = time.perf_counter()
    cal.nn(qs[0], args.k)
    per_code = (time.perf_counter() - t0) / 20_000
    n_calls = max(steps + warmup, 1)
    u_s = int(min(U, max(50_000, budget_s / (per_code * n_calls))))
    u_s = min(u_s, 4_000_000)                          # host memory: x_s is u_s * D * 4 bytes
    idx = make(u_s) if u_s != 20_000 else cal
    for i in range(warmup):
        idx.nn(qs[i], args.k)
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        idx.nn(qs[warmup + i], args.k)
        times.append(time.perf_counter() - t0)
    t_query_sample = sum(times) / len(times)
    t_query_full = t_query_sample * (U / u_s)
    sample = ("literal port of the reference query path (heapq.nsmallest over a set of Python ints, "
              "bin(a^b).count('1')), 1 thread: %d steps x 1 query x %d of the %d codes; "
              "per-code cost %.2f us extrapolated linearly to the full table" % (steps, u_s, U, 1e6 * t_query_sample / u_s))
    return {"value": 1.0 / t_query_full, "unit": "queries/s", "cores": 1, "kind": "port", "sample": sample,
            "ms_per_step": 1e3 * t_query_sample, "host_cores_available": len(os.sched_getaffinity(0))}


def compiled_cpu_leg(args, budget_s):
    """Extra, stronger CPU line: the C/OpenMP restatement (oracle/lsh_oracle.c) of the
    Hamming top-k on all host cores."""
    import numpy as np
    import c_oracle as C
    C.build()
    rng = np.random.RandomState(1)
    W = (args.bits + 31) // 32
    u_s = min(args.rows, 2_000_000)
    db = rng.randint(0, 2 ** 32, size=(u_s, W), dtype=np.uint64).astype(np.uint32)
    threads = C.max_threads()
    q_s = max(threads, 8)
    q = rng.randint(0, 2 ** 32, size=(q_s, W), dtype=np.uint64).astype(np.uint32)
    C.hamming_topk(db[:10000], q, args.k)
    t0 = time.perf_counter()
    reps = 0
    while True:
        C.hamming_topk(db, q, args.k)
        reps += 1
        el = time.perf_counter() - t0
        if el > min(budget_s, 5.0) or reps >= 20:
            break
    per_query_full = el / (reps * q_s) * (args.rows / u_s)
    return {"value": 1.0 / per_query_full, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": "C/OpenMP Hamming top-k only (no hash, no re-rank): %d x %d queries x %d of %d codes, scaled linearly"
                      % (reps, q_s, u_s, args.rows)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    leg = literal_cpu_leg(args, args.steps, args.warmup, budget_s=args.ref_budget_s)
    line = {
        "impl": "reference", "metric": METRIC, "value": leg["value"], "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": leg["value"], "unit": "queries/s", "cores": leg["cores"], "kind": leg["kind"],
                         "sample": leg["sample"]},
        "e2e": {"value": leg["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python on this path and cannot travel to the GPU box; this is its "
                "literal restatement (oracle/ref_port.py, pinned to the reference's outputs in tests/golden)",
    }
    print(json.dumps(line))


def workload_config(args, n_gpus, parallelism=None):
    return {
        "workload": "%s: ITQ-%d hashing + LinearHashIndex Hamming top-k over %dx%d-d fp32 descriptors, "
                    "batch %d queries, k=%d, %s re-rank" % (
                        "configs[3]" if args.config == "c4" else "configs[1]", args.bits, args.rows, args.dim, args.queries, args.k,
                        "histogram-intersection" if args.config == "c4" else "euclidean"),
        "rows": args.rows, "dim": args.dim, "bits": args.bits, "queries_per_step": args.queries, "k": args.k,
        "parallelism": parallelism or ("x%d" % n_gpus if n_gpus > 1 else "single GPU"),
        "l2": "L2 flushed between timed steps (512 MiB write + read-back); code table %d MB" % (args.rows * args.bits // 8 // 10 ** 6),
    }


# --------------------------------------------------------------------------- parity (plain torch, outside the timed region)
def parity_check(torch, dist, functor, state, x_local, row_lo, q_dev, k, result, n_check=8, method="euclidean"):
    """Re-derive the answers of `n_check` queries of the timed batch WITHOUT any kernel of this repo:
    float64 hash (itq.py:404-408), byte-LUT popcount over the whole unique-code table + torch.topk on
    (distance, row) keys (linear.py:232-240 with the canonical tie order), CSR expansion (lsh.py:490-496),
    float64 euclidean distances on the rank that owns each row (summed over ranks with one all-reduce when
    N > 1), stable sort, first k (lsh.py:513-519).  Raises on a mismatch; returns the record for the JSON line."""
    import numpy as np
    rows_out, d_out = result
    dev = q_dev.device
    Q, D = q_dev.shape
    table, csr_off, csr_rows = state.table, state.csr_off, state.csr_rows
    U, W = table.shape
    mean = torch.from_numpy(np.real(np.asarray(functor.mean_vec)).astype(np.float64)).to(dev)
    rot = torch.from_numpy(np.ascontiguousarray(np.real(np.asarray(functor.rotation)), dtype=np.float64)).to(dev)
    b = rot.shape[1]
    z = (q_dev.double() - mean) @ rot
    # a device bit may differ from the float64 bit only for |z| < 1e-5 |q - m| |r| (tests): take queries away from that
    stable = z.abs().min(dim=1).values > 4e-5 * (q_dev.double() - mean).norm(dim=1)
    picks = []
    for c in range(n_check):                                    # spread over the batch: every rank's slice is covered at N = 8
        for qi in range(c * Q // n_check, (c + 1) * Q // n_check):
            if bool(stable[qi]):
                picks.append(qi)
                break
    lut = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int32, device=dev)
    n_local = x_local.shape[0]
    worst = 0.0
    for qi in picks:
        bits = torch.zeros(W * 32, dtype=torch.int64, device=dev)
        bits[W * 32 - b:] = (z[qi] >= 0).to(torch.int64)        # index 0 = most significant bit (bits.py:17-20)
        words = (bits.view(W, 32) << torch.arange(31, -1, -1, device=dev)).sum(dim=1)
        words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
        d_all = torch.empty(U, dtype=torch.int64, device=dev)
        for s0 in range(0, U, 2_000_000):
            x = (table[s0:s0 + 2_000_000] ^ words[None, :]).contiguous().view(torch.uint8)
            d_all[s0:s0 + 2_000_000] = lut[x.long()].sum(dim=1)
        keys = (d_all << 40) | torch.arange(U, device=dev)
        near = (torch.topk(keys, min(k, U), largest=False, sorted=True).values & ((1 << 40) - 1)).tolist()
        cand = torch.cat([csr_rows[int(csr_off[c]):int(csr_off[c + 1])] for c in near])
        mine = (cand >= row_lo) & (cand < row_lo + n_local)
        dd = torch.zeros(cand.numel(), dtype=torch.float64, device=dev)
        if bool(mine.any()):
            xr = x_local[(cand[mine] - row_lo)].double()
            if method == "hik":                                 # metrics.py:70: 1 - sum(min(a, b))
                dd[mine] = 1.0 - torch.minimum(xr, q_dev[qi].double()[None, :]).sum(dim=1)
            else:
                dd[mine] = ((xr - q_dev[qi].double()[None, :]) ** 2).sum(dim=1).sqrt()
        if dist is not None:
            dist.all_reduce(dd, op=dist.ReduceOp.SUM)
        order = torch.sort(dd, stable=True).indices[:k]
        want_rows, want_d = cand[order], dd[order]
        m = want_rows.numel()
        got_rows, got_d = rows_out[qi][:m], d_out[qi][:m]
        rel = float(((got_d - want_d).abs() / want_d.abs().clamp(min=1e-9)).max()) if m else 0.0
        worst = max(worst, rel)
        gaps_ok = m < 2 or float((want_d[1:] - want_d[:-1]).min()) > 1e-6 * float(want_d.max())
        if rel > 1e-5 or (gaps_ok and not torch.equal(got_rows, want_rows)) or bool((rows_out[qi][m:] != -1).any()):
            raise SystemExit("bench.py: PARITY FAILURE on query %d: got rows %s dists %s, torch reference rows %s dists %s"
                             % (qi, got_rows.tolist(), got_d.tolist(), want_rows.tolist(), want_d.tolist()))
    if len(picks) < n_check // 2:
        raise SystemExit("bench.py: parity check found only %d stable queries" % len(picks))
    return {"parity_checked": True, "queries": picks, "max_rel_dist_err": worst,
            "how": "plain torch: float64 hash, byte-LUT popcount + topk over all %d codes, CSR expansion, float64 "
                   "euclidean on the owning rank%s, stable sort; rows exact, distances rtol 1e-5" % (
                       U, " + one SUM all-reduce" if dist is not None else "")}


# --------------------------------------------------------------------------- secondary BASELINE configs
def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the LSH hot path has no CPU implementation")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    return torch, dist, world, rank, dev


def _timed_steps(torch, dist, world, fn, steps, flush):
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    evs = []
    barrier()
    token = torch.zeros(1, dtype=torch.int32, device=flush.device)
    for _ in range(steps):
        flush.zero_()
        flush.view(torch.int64).sum()
        if world > 1:
            dist.all_reduce(token)                             # ranks start the step together (stream-ordered)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    barrier()
    ms = torch.tensor([sum(a.elapsed_time(b_) for a, b_ in evs)], dtype=torch.float64, device=flush.device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms[0])


def _finish(torch, dist, world, index=None):
    if world > 1:
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        if index is not None and hasattr(index, "close"):
            index.close()
        os._exit(0)


def run_c3(args):
    """BASELINE configs[2]: exact brute-force L2 kNN, k = 100, 12.5M x 128-d fp32 rows PER GPU (100M rows on 8)."""
    torch, dist, world, rank, dev = _dist_setup()
    from smqtk_indexing_b200 import _lib
    rows, D, k, Q = 12_500_000, 128, 100, args.queries
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = torch.rand((rows, D), device=dev, generator=g)
    if world > 1:
        from smqtk_indexing_b200.distributed import ShardedFlatL2Index
        index = ShardedFlatL2Index()
        index.build(x)
        query = lambda qq: index.query(qq, k)
    else:
        from smqtk_descriptors.impls.descriptor_set.memory import MemoryDescriptorSet
        from smqtk_indexing_b200.impls.nn_index.flat import FlatL2NearestNeighborsIndex
        index = FlatL2NearestNeighborsIndex(MemoryDescriptorSet())
        index.build_index_matrix(x)
        query = lambda qq: index.nn_batch(qq, k, return_device=True)
    gq = torch.Generator(device=dev).manual_seed(7)
    q = torch.rand((Q, D), device=dev, generator=gq)
    q_host = q.cpu().pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for _ in range(max(args.warmup, 3)):
        r, d = query(q)
    # parity: float64 brute force over ALL shards for 3 queries (each rank its rows, minima all-gathered)
    worst = 0.0
    for qi in (0, Q // 2, Q - 1):
        acc = []
        for s0 in range(0, rows, 2_000_000):
            xs = x[s0:s0 + 2_000_000].double()
            acc.append(((xs - q[qi].double()[None, :]) ** 2).sum(1).sqrt().topk(k, largest=False).values)
            del xs
        loc = torch.cat(acc).topk(k, largest=False).values
        if world > 1:
            allv = [torch.empty_like(loc) for _ in range(world)]
            dist.all_gather(allv, loc)
            loc = torch.cat(allv).topk(k, largest=False).values
        err = float(((d[qi] - loc).abs() / loc.clamp(min=1e-30)).max())
        worst = max(worst, err)
        if err > 1e-9:
            raise SystemExit("bench.py --config c3: PARITY FAILURE on query %d (max rel err %.3g)" % (qi, err))
    _lib.profile_fetch()
    _lib.profile_enable(True)
    launches0 = _lib.launch_count()
    ms = _timed_steps(torch, dist, world, lambda: query(q), args.steps, flush)
    _lib.profile_enable(False)
    launches = _lib.launch_count() - launches0
    per = {}
    for name, t in _lib.profile_fetch():
        per[name] = per.get(name, 0.0) + t / args.steps

    def host_step():
        rr, dd = query(q_host.to(dev, non_blocking=True))
        return rr.cpu(), dd.cpu()
    host_step()
    ms_e2e = _timed_steps(torch, dist, world, host_step, args.steps, flush)
    if rank == 0:
        peak, peak_src = hbm_peak()
        filt = per.get("l2_filter_tma_kernel", 0.0)
        flops = 2.0 * Q * rows * D
        line = {
            "metric": "exact flat L2 kNN queries/s @k=100, 12.5M x 128-d fp32 rows per GPU (BASELINE configs[2]: 100M rows on 8 B200)",
            "value": Q * args.steps / (ms * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32 filter (rigorous margin) + fp32 error-free re-rank", "data": "synthetic",
            "config": {"workload": "configs[2]: exact brute-force L2 kNN over %d x 128-d fp32 rows (%d per GPU x %d GPUs), "
                                   "batch %d queries, k=100" % (rows * world, rows, world, Q),
                       "rows_total": rows * world, "rows_per_gpu": rows, "dim": D, "k": k, "queries_per_step": Q,
                       "parallelism": "rows sharded x%d, local exact top-k, one all-gather of (distance, row), merge" % world
                                      if world > 1 else "single GPU",
                       "l2": "L2 flushed between timed steps; table %.1f GB per GPU" % (rows * D * 4 / 1e9)},
            "e2e": {"value": Q * args.steps / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": Q * D * 4,
                    "d2h_bytes_per_step": Q * k * 16, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "parity_checked": True, "parity": {"queries": [0, Q // 2, Q - 1], "max_rel_dist_err": worst,
                                               "how": "float64 torch brute force over all shards, k-th distance profile, rtol 1e-9"},
            "roofline": {"kernel": "l2_filter_tma_kernel", "bound": "tensor", "achieved": flops / (filt * 1e-3) / 1e12 if filt else None,
                         "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None, "kernel_ms": filt,
                         "algorithmic_flops_per_step": flops,
                         "note": "single-pass TF32 filter: 2*Q*N*D flop per rank and step; no TF32 GEMM peak is measured on "
                                 "this pool (MEASURED_PEAKS.json has bf16 only; TF32 is nominally half of it) -- see "
                                 "profiles/ for the ncu tensor-pipe and DRAM figures of this kernel"},
            "kernel_ms_per_step": per,
        }
        print(json.dumps(line))
    _finish(torch, dist, world)


def run_c5(args):
    """BASELINE configs[4]: ItqFunctor.fit + index build throughput on 50M x 256-d descriptors (row-sharded over N GPUs)."""
    torch, dist, world, rank, dev = _dist_setup()
    from smqtk_indexing_b200 import engine
    from smqtk_indexing_b200.impls.lsh_functor.itq import ItqFunctor
    n_total, D, b = args.c5_rows, 256, args.bits
    n = n_total // world
    X = torch.empty((n, D), device=dev)
    g = torch.Generator(device=dev).manual_seed(50 + rank)
    for s0 in range(0, n, 4_000_000):
        X[s0:s0 + 4_000_000].uniform_(generator=g)
    torch.cuda.synchronize()

    def step():
        f = ItqFunctor(bit_length=b, itq_iterations=args.fit_iters, random_seed=0)
        t0 = time.perf_counter()
        f.fit_matrix(X, want_codes=False, group=dist.group.WORLD if world > 1 else None)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        m = engine.DeviceLshIndex()
        m.set_rows(X, f.get_hash_packed(X))            # hash + sb_unique_codes of the local rows
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        return f, m, t1 - t0, t2 - t1
    step()                                              # warm-up
    fits, builds = [], []
    for _ in range(max(1, min(args.steps, 3))):
        f, m, tf, tb = step()
        fits.append(tf)
        builds.append(tb)
    import numpy as np
    r = np.asarray(f.rotation)
    ortho = float(np.abs(r.T @ r - np.eye(b)).max())
    t = torch.tensor([min(fits), min(builds)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        fit_s, build_s = float(t[0]), float(t[1])
        line = {"metric": "ItqFunctor.fit + build_index descriptors/s on %dM x 256-d (BASELINE configs[4])" % (n_total // 10 ** 6),
                "value": n_total / (fit_s + build_s), "unit": "descriptors/s", "n_gpus": world, "steps": len(fits), "warmup": 1,
                "ms_per_step": 1e3 * (fit_s + build_s), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "fp64 covariance + tensor-core streaming ITQ iterations (3xTF32 sign step, fixed-point TF32 Gram)",
                "data": "synthetic",
                "config": {"workload": "configs[4]: ItqFunctor.fit (%d iterations, %d bits) + hash + unique table + CSR over %d x 256-d "
                                       "fp32 descriptors" % (args.fit_iters, b, n_total), "rows": n_total, "rows_per_gpu": n,
                           "parallelism": "rows sharded x%d: all-reduce of [D], [D,D], [b,D] partial sums in fit; every rank "
                                          "indexes its own rows" % world if world > 1 else "single GPU"},
                "fit_s": fit_s, "build_s": build_s, "fit_descriptors_per_s": n_total / fit_s, "build_descriptors_per_s": n_total / build_s,
                "model_orthonormality_err": ortho, "parity_checked": ortho < 1e-8,
                "e2e": {"value": n_total / (fit_s + build_s), "unit": "descriptors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "training data is generated on the device (51 GB would not stream through PCIe per step); no host leg"}}
        print(json.dumps(line))
    _finish(torch, dist, world)


# --------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from smqtk_indexing_b200 import _lib, engine
    from smqtk_indexing_b200.impls.lsh_functor.itq import ItqFunctor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the LSH hot path has no CPU implementation")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    if args.config == "c4":      # BASELINE configs[3]: 1M x 4096-d L1-normalised histograms, hik re-rank, k = 50
        args.rows, args.dim, args.bits, args.k = 1_000_000, 4096, 256, 50
        args.queries = min(args.queries, 1024)
        args.fit_rows = min(args.fit_rows, 20_000)
        args.no_cpu_baseline = True      # the literal port's CPU leg is sized for configs[1]
    METHOD = "hik" if args.config == "c4" else "euclidean"
    U, D, b, Q, k = args.rows, args.dim, args.bits, args.queries, args.k
    bounds = [U * r // world for r in range(world + 1)]
    n_local = bounds[rank + 1] - bounds[rank]

    # ---- synthetic descriptors (shard generated on its GPU), model, index ----
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    x_local = torch.rand((n_local, D), generator=g, device=dev, dtype=torch.float32)
    if METHOD == "hik":
        for s0 in range(0, n_local, 100_000):
            x_local[s0:s0 + 100_000] /= x_local[s0:s0 + 100_000].sum(dim=1, keepdim=True)
    functor = ItqFunctor(bit_length=b, itq_iterations=args.fit_iters, random_seed=0)
    t_fit0 = time.perf_counter()
    if world == 1:
        functor.fit_matrix(x_local[:min(args.fit_rows, n_local)], want_codes=False)
    else:
        # row-sharded fit: every rank passes 1/N of the training rows, the partial sums are all-reduced
        # (fit.py: [D], [D, D] and [b, D] values per step), every rank ends with the same model
        functor.fit_matrix(x_local[:min((args.fit_rows + world - 1) // world, n_local)], want_codes=False,
                           group=dist.group.WORLD)
    torch.cuda.synchronize()
    fit_s = time.perf_counter() - t_fit0

    t_build0 = time.perf_counter()
    if world == 1:
        from smqtk_dataprovider.impls.key_value_store.memory import MemoryKeyValueStore
        from smqtk_descriptors.impls.descriptor_set.memory import MemoryDescriptorSet
        from smqtk_indexing_b200.impls.hash_index.linear import LinearHashIndex
        from smqtk_indexing_b200.impls.nn_index.lsh import LSHNearestNeighborIndex
        index = LSHNearestNeighborIndex(functor, MemoryDescriptorSet(), MemoryKeyValueStore(), LinearHashIndex(),
                                        distance_method=METHOD)
        index.build_index_matrix(x_local)
        n_codes = index._mirror.num_codes
        scan_rows = n_codes

        use_graph = [not args.no_graph]

        def query_dev(qd):
            return index.nn_batch(qd, k, return_device=True, graph=use_graph[0])

        def query_host(qh):
            return index.nn_batch(qh, k, graph=use_graph[0])
        parallelism = "single GPU"
        state = index._mirror
        peers = None
    else:
        from smqtk_indexing_b200.distributed import ShardedLshIndex
        index = ShardedLshIndex(functor, METHOD, scan_partition=args.scan_partition, rerank=args.rerank,
                                graph=not args.no_graph)
        index.build(x_local)
        n_codes = index.num_codes
        by_q = index._by_queries(Q)
        # rows of the code table each rank scans, and for how many of the batch's queries
        scan_rows = n_codes if by_q else index.scan_hi - index.scan_lo
        use_graph = [not args.no_graph]

        def query_dev(qd):
            index.graph = use_graph[0]
            return index.query(qd, k)

        def query_host(qh):
            index.graph = use_graph[0]
            return index.query_host(qh, k)
        parallelism = ("x%d: descriptors row-sharded; %s; re-rank %s" % (
            world,
            "scan split over the QUERIES (code table replicated, each rank scans all %d codes for %d of the %d queries)"
            % (n_codes, (Q + world - 1) // world, Q) if by_q else
            "scan split over table ROWS (each rank scans %d of %d codes for all queries; all-gather of local top-k keys + merge)"
            % (index.scan_hi - index.scan_lo, n_codes),
            "per query slice with candidate rows read from the owners' HBM over NVLink (CUDA IPC), one all-gather of results"
            if index.peers is not None else
            "owner-computes + SUM all-reduce of candidate distances (peer access unavailable: %s)" % index.peer_error))
        state = index
        peers = index.peers
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t_build0

    # ---- queries: identical on every rank, pinned on the host ----
    gq = torch.Generator().manual_seed(1)
    q_host = torch.rand((Q, D), generator=gq, dtype=torch.float32)
    if METHOD == "hik":
        q_host /= q_host.sum(dim=1, keepdim=True)
    q_host = q_host.pin_memory()
    q_dev = q_host.to(dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():
        """Evict L2 between timed steps: write a 512 MiB buffer (the contract's flush),
        then read it back so the lines left in L2 are CLEAN -- otherwise the first
        ~126 MB the next kernel reads pay for the write-back of the flush itself."""
        flush.zero_()
        flush.view(torch.int64).sum()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_token = torch.zeros(1, dtype=torch.int32, device=dev)

    def timed(fn, arg, steps, on_step=None):
        """EXACTLY `steps` steps; per-step CUDA events (the L2 flush between steps is
        outside the events); returns the summed milliseconds.  `on_step` runs after each step has
        been launched (the clock sampler's synchronous sample: the GPU is busy at that point)."""
        evs = []
        barrier()
        for _ in range(steps):
            flush_l2()
            if world > 1:
                # line the ranks up ON THE DEVICE before the step's first event: a step contains a collective, so a
                # rank whose host is late (rank 0 samples the clocks after every step) would otherwise be waited for
                # inside the other ranks' events.  Stream-ordered, no host synchronisation.
                dist.all_reduce(step_token)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(arg)
            e1.record()
            evs.append((e0, e1))
            if on_step is not None:
                on_step()
        barrier()
        return sum(a.elapsed_time(b_) for a, b_ in evs)

    for _ in range(max(args.warmup, 3)):
        query_dev(q_dev)                       # (the 2nd batch of a shape captures the CUDA graph)
    barrier()

    # ---- parity, outside the timed region, at every N: 8 queries of the batch re-derived with plain torch ----
    parity = parity_check(torch, dist if world > 1 else None, functor, state, x_local, bounds[rank], q_dev, k,
                          query_dev(q_dev), method=METHOD)
    barrier()

    gpu_id = "GPU-%s" % torch.cuda.get_device_properties(dev).uuid if hasattr(
        torch.cuda.get_device_properties(dev), "uuid") else str(local_rank)
    sampler = ClockSampler(gpu_id)
    if rank == 0:
        sampler.start()
    # timed region A (the reported value): the product's default path
    ms_dev = timed(query_dev, q_dev, args.steps, on_step=(sampler.sample_between_steps if rank == 0 else None))
    clocks = sampler.stop() if rank == 0 else None
    # timed region B: the same K steps launched kernel by kernel with an event pair around every launch
    # (a graph replay has no host-side launches to bracket): per-kernel durations, launch count
    use_graph[0] = False
    query_dev(q_dev)
    launches0 = _lib.launch_count()
    _lib.profile_fetch()
    _lib.profile_enable(True)
    ms_eager = timed(query_dev, q_dev, args.steps)
    _lib.profile_enable(False)
    launches = _lib.launch_count() - launches0
    prof = _lib.profile_fetch()
    use_graph[0] = not args.no_graph

    # ---- end to end: pinned host queries in, host results out, through the public call ----
    for _ in range(2):
        query_host(q_host)
    ms_e2e = timed(query_host, q_host, args.steps)

    # ---- single-query scan (the reference's per-call shape; HBM-bound) ----
    from smqtk_indexing_b200 import device as devops
    table = index._mirror.table if world == 1 else index.table
    one_q = functor.get_hash_packed(q_dev[:1])
    for _ in range(3):
        devops.hamming_scan_keys(table, one_q, k)
    _lib.profile_fetch()
    _lib.profile_enable(True)
    for _ in range(20):
        flush_l2()
        devops.hamming_scan_keys(table, one_q, k)
    _lib.profile_enable(False)
    prof1 = [ms for name, ms in _lib.profile_fetch() if name == "hamming_scan_kernel"]

    # ---- max over ranks ----
    t = torch.tensor([ms_dev, ms_e2e, ms_eager], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_eager = float(t[0]), float(t[1]), float(t[2])

    fp8_meas = measure_fp8_gemm(dev) if rank == 0 else None
    from smqtk_indexing_b200 import device as _devmod
    scan_fmt = _devmod.TC_SCAN_FORMAT
    fp4_meas = measure_fp4_gemm(dev) if (rank == 0 and scan_fmt == "fp4") else None
    if rank == 0:
        W = (b + 31) // 32
        per_kernel = {}
        for name, ms in prof:
            per_kernel.setdefault(name, []).append(ms)
        peak, peak_src = hbm_peak()
        ms_step = ms_dev / args.steps
        ms_step_eager = ms_eager / args.steps
        q_per_rank = (Q + world - 1) // world if (world > 1 and index._by_queries(Q)) else Q
        tc = per_kernel.get("ham_filter_tc_kernel", [])
        if tc:
            # batched path: +-1 dot products on tcgen05; several launches per step (chunks).  Default: packed FP4
            # operands, TC4_QUERIES_PER_COLUMN queries per FP32 accumulator column (hamming_tc4.cu); SB_TC_SCAN_FORMAT=fp8: E4M3
            # operands, FP16 accumulators (hamming_tc.cu)
            scan_ms = sum(tc) / args.steps
            alg_flops = 2.0 * q_per_rank * scan_rows * (32 * W)           # this rank's share of 2*Q*U*b
            achieved = alg_flops / (scan_ms * 1e-3) / 1e12
            proxy, proxy_src = fp8_peak()
            if scan_fmt == "fp4":
                nominal = 9000.0
                if fp4_meas:
                    tpeak, tpeak_src = fp4_meas, "cuBLASLt block-scaled FP4 (E2M1, 1x16 E4M3 scales) GEMM 8192^3 (torch._scaled_mm), best of 10, timed in this run"
                elif fp8_meas:
                    tpeak, tpeak_src = 2.0 * fp8_meas, ("2 x the cuBLASLt FP8 E4M3 GEMM 8192^3 timed in this run (no MXFP4 GEMM in "
                                                        "this torch build; kind::mxf4 issues at twice the f8f6f4 rate)")
                else:
                    tpeak, tpeak_src = 2.0 * proxy, "2 x (" + proxy_src + ")"
            else:
                nominal = 4500.0
                tpeak, tpeak_src = (fp8_meas, "cuBLASLt FP8 E4M3 GEMM 8192^3 (torch._scaled_mm), best of 10, timed in this run") \
                    if fp8_meas else (proxy, proxy_src)
            traffic = recorded_traffic("scan_tc_traffic.json")
            # tensor-memory read floor: every (row, query) accumulator leaves TMEM through tcgen05.ld at 64 B/clk/SM
            # (B300_MICROARCH.md "LDTM throughput"); bytes per pair = 4 / TC4_QUERIES_PER_COLUMN (FP4) or 2 (FP8)
            tmem_bpp = (4.0 / TC4_QUERIES_PER_COLUMN) if scan_fmt == "fp4" else 2.0
            sm_hz = (clocks.get("sm_mhz") or 1800.0) * 1e6
            tmem_floor_ms = float(q_per_rank) * scan_rows * tmem_bpp / (64.0 * 148 * sm_hz) * 1e3
            roofline = {
                "kernel": "ham_filter_tc_kernel (%s operands)" % scan_fmt, "bound": "tensor",
                "achieved": achieved, "peak": tpeak, "unit": "TFLOP/s", "frac": achieved / tpeak,
                "traffic": traffic.get("dram_bytes_per_step") if traffic else None,
                "traffic_source": "profiles/scan_tc_traffic.json (ncu dram__bytes_read+write of one batch, N=1 shape)"
                                  if traffic else None,
                "peak_source": tpeak_src,
                "frac_of_2x_measured_bf16": achieved / proxy, "proxy_peak": proxy,
                "nominal_peak": nominal, "frac_of_nominal": achieved / nominal,
                "tmem_read_floor_ms": tmem_floor_ms, "frac_of_tmem_read_floor": tmem_floor_ms / scan_ms,
                "algorithmic_flops_per_step": alg_flops, "launches_per_step": len(tc) / args.steps,
                "kernel_ms": scan_ms, "kernel_share_of_step": scan_ms / ms_step_eager,
                "note": "per rank.  algorithmic flops = 2*Q*U*b: one multiply-add per (query, code, bit) of this rank's share; "
                        "the kernel issues b+64 (fp4) / b+32 (fp8) per pair (the threshold rides in one extra K step).  Durations "
                        "come from the kernel-by-kernel pass (region B, %.3f ms/step); the reported value is the graph-replayed "
                        "pass (region A).  A +-1 operand GEMM draws less power than cuBLAS's random-data GEMM and holds a "
                        "higher clock, so frac can exceed 1 against a measured GEMM peak; frac_of_nominal uses the datasheet "
                        "figure.  The second ceiling is the tensor-memory read path (tmem_read_floor_ms at the sampled SM "
                        "clock).  Equivalent algorithmic bytes (Q*U*b/8 per step): %.1f TB/s." % (
                            ms_step_eager, float(q_per_rank) * scan_rows * W * 4 / (scan_ms * 1e-3) / 1e12),
            }
            int_pipe = None
        else:
            scan = per_kernel.get("hamming_scan_kernel", [])
            scan_ms = sum(scan) / args.steps if scan else float("nan")
            alg_bytes = float(q_per_rank) * scan_rows * W * 4
            achieved = alg_bytes / (scan_ms * 1e-3) / 1e9
            traffic = recorded_traffic()
            roofline = {
                "kernel": "hamming_scan_kernel<%d>" % W, "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms": scan_ms, "kernel_share_of_step": scan_ms / ms_step_eager if scan else None,
                "note": "algorithmic bytes = Q*U*(b/8): every query against every code; each loaded code word "
                        "is reused from registers for the CTA's whole query tile, so frac > 1 is expected and the "
                        "binding resources are the ALU (LOP3) and XU (POPC) pipes -- see int_pipe and profiles/",
            }
            int_pipe = {"pairs_per_s": float(q_per_rank) * scan_rows / (scan_ms * 1e-3), "popc_per_pair": 4, "lop3_per_pair": 16}
        # stage 3 (HBM-bound gather): candidates x D x 4 bytes over the re-rank kernel's duration
        rr = per_kernel.get("rerank_kernel", [])
        rr_ms = sum(rr) / args.steps if rr else None
        rerank_line = None
        if rr_ms:
            rr_bytes = float(q_per_rank) * k * max(state.max_rows_per_code, 1) * D * 4
            rerank_line = {"kernel": "rerank_kernel<%s>" % METHOD, "bound": "hbm", "kernel_ms": rr_ms,
                           "candidate_slots_per_step": int(q_per_rank * k * max(state.max_rows_per_code, 1)),
                           "algorithmic_bytes_per_step": rr_bytes, "achieved_gbs": rr_bytes / (rr_ms * 1e-3) / 1e9,
                           "frac_of_peak": rr_bytes / (rr_ms * 1e-3) / 1e9 / peak, "peak": peak, "peak_source": peak_src,
                           "note": "one warp per candidate row (random row gathers%s); %d slots x %d B in %.1f us -- "
                                   "latency-bound at the C2 size, see profiles/r2_rerank_c4_full.md for the C4 shape" % (
                                       ", peer rows over NVLink" if peers is not None else "",
                                       int(q_per_rank * k * max(state.max_rows_per_code, 1)), D * 4, rr_ms * 1e3)}
        line = {
            "metric": METRIC if args.config == "c2" else
            "LSH kNN queries/s @k=50, 1M x 4096-d L1-normalised histograms, ITQ-256, histogram-intersection re-rank (BASELINE configs[3])",
            "value": Q * args.steps / (ms_dev * 1e-3), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": ("%s (+-1 operands, exact integer distances)" % scan_fmt) if tc else "u32", "data": "synthetic",
            "config": workload_config(args, world, parallelism),
            "clocks": clocks,
            "e2e": {"value": Q * args.steps / (ms_e2e * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": (q_per_rank if (world > 1 and peers is not None) else Q) * D * 4,
                    "d2h_bytes_per_step": Q * k * 16,
                    "ms_per_step": ms_e2e / args.steps},
            "e2e_note": ("per rank: with the scan split over the queries a rank copies only its slice of the batch "
                         "to the device (h2d_bytes_per_step) and reads the whole all-gathered result back"
                         if world > 1 and peers is not None else "whole batch in, whole result out"),
            "gpu_launches": int(launches),
            "gpu_launches_note": "kernels of this repo per %d steps, counted in the kernel-by-kernel pass (region B); the "
                                 "graph-replayed pass (region A) runs the same kernels from one cudaGraphLaunch per step%s"
                                 % (args.steps, "" if not args.no_graph else " [graphs disabled: A == B]"),
            "graph_replay": not args.no_graph,
            "eager": {"ms_per_step": ms_step_eager, "value": Q * args.steps / (ms_eager * 1e-3)},
            "parity_checked": bool(parity.get("parity_checked")), "parity": parity,
            "roofline": roofline,
            "rerank": rerank_line,
            "int_pipe": int_pipe,
            "fp8_gemm_tflops_measured": fp8_meas, "fp4_gemm_tflops_measured": fp4_meas,
            "single_query_scan": {
                "what": "sb_hamming_scan with Q=1 (LinearHashIndex.nn call shape): HBM-bound, whole table",
                "kernel_ms": statistics.median(prof1) if prof1 else None,
                "achieved_gbs": (table.shape[0] * W * 4 / (statistics.median(prof1) * 1e-3) / 1e9) if prof1 else None,
                "frac_of_peak": (table.shape[0] * W * 4 / (statistics.median(prof1) * 1e-3) / 1e9 / peak) if prof1 else None,
            },
            "kernel_ms_per_step": {n_: sum(v) / args.steps for n_, v in per_kernel.items()},
            "index": {"unique_codes": int(n_codes), "rows_local": int(n_local), "scan_rows_local": int(scan_rows),
                      "queries_scanned_per_rank": int(q_per_rank), "build_s": build_s, "fit_s": fit_s,
                      "peer_rerank": peers is not None},
        }
        if world == 1 and not args.no_cpu_baseline:
            leg = literal_cpu_leg(args, steps=3, warmup=1, budget_s=args.cpu_budget_s)
            line["cpu_baseline"] = {k_: leg[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
            try:
                line["cpu_baseline_compiled"] = compiled_cpu_leg(args, budget_s=5.0)
            except Exception as e:           # the compiled checker is optional
                line["cpu_baseline_compiled"] = {"unavailable": str(e)[:120]}
        print(json.dumps(line))
    if world > 1:
        # orderly teardown, then a hard exit: captured graphs that hold NCCL work and CUDA-IPC mappings of
        # exited peers can stall interpreter shutdown; nothing is left to flush but stdout
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        index.close()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.config in ("c3", "c5") and not (args.gpus > 1 and "RANK" not in os.environ):
        (run_c3 if args.config == "c3" else run_c5)(args)
        return
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_b200(args)


if __name__ == "__main__":
    main()

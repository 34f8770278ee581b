#!/usr/bin/env python
"""bench.py -- image-pair matches/sec of the all-pairs 2-NN matching path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg5|...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU BFMatcher path on the host cores

A "step" is one pass of the hot path over the whole workload: every image pair q<t of the
synthetic descriptor set (findBestPair's loop, /root/reference/src/Sfm.cpp:511-515) goes through
2-NN + ratio test (getMatching, src/Sfm.cpp:590-608).  One "pair-match" = one such pair, its match
list delivered to rank 0's HOST memory.

Default workload at every N: BASELINE.json configs[2] -- 200 images x 10 000 x 486-bit (AKAZE-shape)
descriptors, all 19 900 pairs, STRONG scaling: the same job on 1, 2, 4, 8 GPUs (it fits one GPU, and it is
the configuration the metric "at 1/2/4/8 B200" is quoted on).  configs[1] (cfg2, 50 x 5 000) and the float
shape (cfg4s) ride along in the N=1 line as `configs1_cfg2` and `float` blocks.

  value  : pair-matches/s with the descriptors already resident in HBM on every rank when the timed region
           starts; the region covers matching on all ranks AND the gather of every match list to rank 0's host
           memory (the library's pipelined device->host path on every rank; at N>1 into shared page-locked tables
           rank 0 maps, --gather shared, with the NCCL forms of the gather measured beside it as gather_alt).
           Timed on the device (CUDA events around each step), max over ranks.
  e2e    : the same metric through the public API with HOST buffers in: H2D of the descriptors, (NCCL
           broadcast), matching, (NCCL gather to rank 0), D2H of the match table.  Wall clock, max over ranks.
  resident_device_only: the round-1 definition of `value` (match lists left in HBM), for continuity.
  roofline: the 2-NN kernel that actually ran against its bound: the tensor pipe (default engines: `tensor_kind` says which --
           kind::mxf4 for binary descriptors below 512 bit, with `vs_i8_pipe` beside it for continuity with the kind::i8
           kernel of round 1; kind::i8 at 512 bit; kind::f16 for floats) or the POPC pipe / FP32 lanes (--binary-engine
           popc, --float-mode exact).  Peaks: measured tcgen05 issue rates (profiles/tcgen05_peaks_r02.json, from
           tools/tc_bench.cu and tools/mxf4_bench.cu) when present, else MEASURED_PEAKS.json scaled.
  alt_engine: the other Hamming engine on the same shard in the same run.
  verified: `--verify` (default 2) random pairs of the end-to-end table, byte-compared with the CPU oracle.
  cpu_baseline: the reference's own CPU path (OpenCV BFMatcher via cv2, else the C oracle) on a bounded sample
           of the same pairs, timed on this box's host cores (N=1, rank 0).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, images at N=1, descriptors per image, weak-scale images with N?)
    "cfg2": ("binary", 50, 5000, False),    # configs[1]: 50 x 5k x 486 bit, 1 GPU
    "cfg2w": ("binary", 50, 5000, True),    # the same per-GPU work at every N (round 1's default)
    "cfg3": ("binary", 200, 10000, False),  # configs[2]: 200 x 10k, sharded over 2/4/8 -- the default
    "cfg4": ("float", 300, 8000, False),    # configs[3]: 300 x 8k x 128 f32
    "cfg5": ("binary", 1000, 20000, False), # configs[4]: headline, ~500k pairs
    "cfg5s": ("binary", 40, 20000, True),   # cfg5's pair shape (20k x 20k) on a 40-image subset
    "orb": ("orb", 100, 2000, True),        # ORB shape: 256 bit
    "float2": ("float", 40, 4000, True),
    "cfg4s": ("float", 60, 8000, True),     # cfg4's pair shape (8k x 8k x 128 f32) on a 60-image subset
    "cfg4x": ("floatx", 60, 8000, True),    # same shape, NON-integer values: fp16/TF32 ranking + exact refinement path
    # configs[0]: the reference's own fixture (data/temple, 10 images, 45 pairs) through the committed cv2 descriptors
    "temple_sift": ("golden:temple_sift", 10, 850, False),
    "temple_akaze": ("golden:temple_akaze", 10, 700, False),
}
DEFAULT_WORKLOAD = "cfg3"


def images_for(n1: int, gpus: int, weak: bool) -> int:
    if not weak or gpus == 1:
        return n1
    target = gpus * n1 * (n1 - 1) // 2
    n = int(math.ceil((1 + math.sqrt(1 + 8 * target)) / 2))
    while n * (n - 1) // 2 < target:
        n += 1
    return n


def make_descriptors(kind: str, n_images: int, n_desc: int, seed: int = 0):
    from sfm_danpipeline_b200 import synth
    if kind.startswith("golden:"):  # descriptors cv2 extracted from /root/reference/data/temple (tests/golden/make_golden.py)
        z = np.load(os.path.join(ROOT, "tests", "golden", kind.split(":")[1] + ".npz"))
        offs = np.concatenate([[0], np.cumsum(z["rows"])])
        is_f = "sift" in kind
        d = z["desc"].astype(np.float32 if is_f else np.uint8)
        return [np.ascontiguousarray(d[offs[i]:offs[i + 1]]) for i in range(len(z["rows"]))], (1 if is_f else 0)
    if kind in ("binary", "orb") and n_images * n_desc >= 1_000_000:
        # large sets (cfg3, cfg5): same generator, same seeds, images dealt to worker processes
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        bits = synth.AKAZE_BITS if kind == "binary" else synth.ORB_BITS
        with ProcessPoolExecutor(min(32, os.cpu_count() or 1), mp_context=mp.get_context("spawn")) as ex:
            return list(ex.map(synth.binary_image_task, [(i, n_desc, bits, seed) for i in range(n_images)], chunksize=4)), 0
    if kind == "binary":
        return synth.binary_images(n_images, n_desc, synth.AKAZE_BITS, seed), 0
    if kind == "orb":
        return synth.binary_images(n_images, n_desc, synth.ORB_BITS, seed), 0
    return synth.float_images(n_images, n_desc, 128, seed, integer=(kind != "floatx")), 1


def workload_string(name, kind, n_images, n_desc, cross):
    shape = ("486-bit AKAZE" if kind in ("binary", "golden:temple_akaze") else ("256-bit ORB" if kind == "orb" else "128-d f32 SIFT")) \
        + ("-shape" if not kind.startswith("golden:") else " (real, about that many rows per image)")
    n_pairs = n_images * (n_images - 1) // 2
    return f"{name}: {n_images} images x {n_desc} x {shape} descriptors, all {n_pairs} pairs q<t, ratio 0.8, cross_check={bool(cross)}"


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (pynvml)."""

    def __init__(self, index: int, period: float = 0.1):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = get(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return None
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz or 0),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------ CPU arm
def cpu_pair_fn(norm: int):
    """(callable(Q,T)->n_matches, kind, cores, description).  cv2 is the very OpenCV code the
    reference links (cv::BFMatcher::knnMatch -> cv::batchDistance); the C oracle is the port."""
    import oracle
    cores = os.cpu_count() or 1
    if oracle.have_cv2():
        import cv2
        cv2.setNumThreads(cores)
        ratio = np.float32(0.8)

        def run(Q, T):
            d, _i = oracle.knn2_cv2(Q, T, norm)
            d = d.astype(np.float32)
            return int((d[:, 0] <= ratio * d[:, 1]).sum())

        return run, "reference", cv2.getNumThreads(), f"cv2 {cv2.__version__} batchDistance(K=2)+ratio (OpenCV BFMatcher code path)"
    oracle.build()

    def run(Q, T):
        return len(oracle.match_pair(Q, T, norm, 0.8, False, threads=cores))

    return run, "port", cores, "oracle/bf_oracle.c knn2 + ratio, python thread pool over query rows"


def time_cpu_sample(descs, norm, budget_s: float, max_pairs: int, seed: int = 0):
    from sfm_danpipeline_b200 import synth
    run, kind, cores, how = cpu_pair_fn(norm)
    pairs = synth.all_pairs(len(descs))
    rng = np.random.default_rng(seed)
    order = rng.permutation(len(pairs))
    q, t = pairs[order[0]]
    run(descs[q], descs[t])  # warm-up pair
    done, t0 = 0, time.perf_counter()
    while done < max_pairs and time.perf_counter() - t0 < budget_s:  # cycle through the sample until the budget is spent
        q, t = pairs[order[(1 + done) % len(order)]]
        run(descs[q], descs[t])
        done += 1
    dt = time.perf_counter() - t0
    return done / dt, kind, cores, f"{done} of {len(pairs)} pairs ({how}), {dt:.1f} s", dt


def run_reference_arm(args, name, kind, n_images, n_desc):
    """The reference's own CPU implementation of the path on this box's host cores, same workload string as our arm.
    Under torchrun rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the pairs of the workload all have the same shape: a subset of the images is enough to draw the bounded sample from
    descs, norm = make_descriptors(kind, min(n_images, 24), n_desc, args.seed)
    per_step, step_s = [], []
    how = cores = kind_s = None
    for s in range(args.warmup + args.steps):
        v, kind_s, cores, how, dt = time_cpu_sample(descs, norm, args.cpu_seconds / max(args.steps, 1), 100000, seed=s)
        if s >= args.warmup:
            per_step.append(v)
            step_s.append(dt)
    value = float(np.mean(per_step))
    n_pairs = n_images * (n_images - 1) // 2
    line = {"impl": "reference", "metric": "image-pair matches/sec (all-pairs 2-NN + ratio test)", "value": value,
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean(step_s)), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8" if norm == 0 else "f32", "data": "synthetic",
            "config": {"workload": workload_string(name, kind, n_images, n_desc, False),
                       "sample": "each step = a bounded random sample of the workload's pairs (all pairs have the same shape), "
                                 f"drawn from the first {min(n_images, 24)} images; the whole workload would take {n_pairs / value:.0f} s"},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind_s, "sample": how},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def verify_pairs(get, pairs, descs, norm, n, cross):
    """Bit-exactness spot check of the end-to-end table at full size: n random pairs vs the CPU oracle."""
    import oracle
    rng = np.random.default_rng(123)
    ok = 0
    integer_valued = norm == 0 or all(bool((d == np.floor(d)).all()) for d in descs[:4])
    worst = 0.0
    for k in rng.choice(len(pairs), min(n, len(pairs)), replace=False):
        q, t = int(pairs[k][0]), int(pairs[k][1])
        exp = oracle.match_pair_cv2(descs[q], descs[t], norm, 0.8, cross) if oracle.have_cv2() else \
            oracle.match_pair(descs[q], descs[t], norm, 0.8, cross, threads=os.cpu_count() or 1)
        got = np.asarray(get(q, t))
        if integer_valued:  # Hamming, and L2 on integer-valued (SIFT-like) data: every fp32 partial sum is exact
            assert got.tobytes() == exp.tobytes(), f"pair ({q},{t}) differs from the oracle"
        else:
            # arbitrary floats: north_star's bar -- distances within 1e-4 relative; a match may appear or vanish only
            # where the ratio test is decided inside that tolerance (fp32 summation order differs from OpenCV's)
            both = np.intersect1d(got["queryIdx"], exp["queryIdx"])
            g = got[np.isin(got["queryIdx"], both)]
            e = exp[np.isin(exp["queryIdx"], both)]
            rel = np.abs(g["distance"] - e["distance"]) / np.maximum(e["distance"], 1e-30)
            same = g["trainIdx"] == e["trainIdx"]
            assert (rel[same] <= 1e-4).all(), f"pair ({q},{t}): distance outside 1e-4 relative"
            flips = (len(got) - len(both)) + (len(exp) - len(both)) + int((~same).sum())
            assert flips <= max(2, len(exp) // 1000), f"pair ({q},{t}): {flips} decisions differ"
            worst = max(worst, float(rel[same].max()) if same.any() else 0.0)
        ok += 1
    return {"pairs_checked": ok, "against": "cv2 BFMatcher path" if oracle.have_cv2() else "C oracle",
            "result": "bit-identical" if integer_valued else f"within 1e-4 relative (worst {worst:.2e})"}


# ------------------------------------------------------------------------------ peaks
def load_peaks():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tc = {}
    try:  # measured tcgen05.mma issue rates of this pool (tools/pipe_bench tcgen05 -> profiles/tcgen05_peaks_r02.json)
        tc = json.load(open(os.path.join(ROOT, "profiles", "tcgen05_peaks_r02.json")))
    except Exception:
        pass
    traffic = {}
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu captures
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))
    except Exception:
        pass
    return peaks, tc, traffic


def roofline_block(stats_path, norm, engine, knn_ms, knn_work, knn_launches, step_ms_sum, n_sm, peaks, tc, tensor_kind=0):
    """The dominant kernel against the pipe that bounds it.  knn_work = algorithmic POPC32 ops (Hamming) / FLOPs (L2)."""
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    bf16 = float(peaks.get("bf16_tflops", 1590.0))
    bf16_src = "MEASURED_PEAKS.json bf16_tflops (cuBLAS burst)" if "bf16_tflops" in peaks else "fallback 1590 TFLOP/s (B200_PROFILING.md)"
    knn_s = max(knn_ms, 1e-9) * 1e-3
    if engine == "tensor":
        macs = knn_work / 16.0 * 512.0  # 512 u8 MACs per 16 algorithmic POPC32 (one 486-bit distance)
        ach = 2.0 * macs / knn_s / 1e12
        i8_peak = float(tc["i8_tops"]) if "i8_tops" in tc else 2.0 * bf16
        if tensor_kind == 2:  # TM_F4P: bits as E2M1 nibbles on the FP4 pipe
            peak = float(tc.get("mxf4_tops", 2.0 * i8_peak))
            src = ("of measured: tcgen05.mma kind::mxf4.block_scale issue-only microbenchmark on this pool (tools/mxf4_bench.cu, profiles/tcgen05_peaks_r02.json)"
                   if "mxf4_tops" in tc else "of 2 x the kind::i8 peak (kind::mxf4 contracts 64 elements in the 64 cycles kind::i8 needs for 32)")
            how = "tcgen05 kind::mxf4 on bits unpacked to E2M1 nibbles, unit block scales"
        elif "i8_tops" in tc:
            peak, src = i8_peak, "of measured: tcgen05.mma kind::i8 issue-only microbenchmark on this pool (profiles/tcgen05_peaks_r02.json)"
            how = "tcgen05 kind::i8 on bits unpacked to bytes"
        else:
            peak, src = i8_peak, f"of measured (inferred): 2 x {bf16_src}; int8 dense is nominally twice bf16"
            how = "tcgen05 kind::i8 on bits unpacked to bytes"
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TOP/s", "tensor_kind": {1: "i8", 2: "mxf4"}.get(tensor_kind, "i8"),
                "peak_source": src + "; ops = 2*Nq*Nt*512 per pair, " + how,
                "vs_i8_pipe": {"peak": i8_peak, "frac": ach / i8_peak, "note": "the same work against the kind::i8 rate round 1 and the first half of round 2 ran at"},
                "nominal": {"peak": 9000.0 if tensor_kind == 2 else 4500.0, "frac": ach / (9000.0 if tensor_kind == 2 else 4500.0),
                            "note": "B200 dense nominal: fp4 9 POP/s, int8/fp8 4.5 POP/s"},
                "popc_equivalent": {"achieved_GPOPC32": knn_work / knn_s / 1e9,
                                    "x_nominal_popc_roofline": knn_work / knn_s / 1e9 / (n_sm * 16 * sm_max_mhz * 1e6 / 1e9)}}
    elif norm == 0:
        peak = n_sm * 16 * sm_max_mhz * 1e6 / 1e9  # GPOPC32/s: 16 POPC/clk/SM x SMs x max SM clock
        ach = knn_work / knn_s / 1e9
        roof = {"bound": "popc", "achieved": ach, "peak": peak, "unit": "GPOPC32/s",
                "peak_source": f"nominal 16 POPC/clk/SM x {n_sm} SMs x {sm_max_mhz:.0f} MHz (measured 16.0/clk/SM, profiles/pipe_bench_r01.txt)",
                "executed_popc_frac": ach * 9.0 / 16.0 / peak,
                "note": "carry-save compression executes 9 POPC per 16 algorithmic ones (486-bit rows): `frac` counts algorithmic work and can "
                        "exceed 1, `executed_popc_frac` is the share of the XU pipe's issue slots actually used"}
    elif stats_path in (2, 3):
        ach = knn_work / knn_s / 1e12
        if "f16_tflops" in tc:
            peak, src = float(tc["f16_tflops"]), "of measured: tcgen05.mma kind::f16 issue-only microbenchmark on this pool (profiles/tcgen05_peaks_r02.json)"
        else:
            peak, src = bf16, f"of measured: {bf16_src}"
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "peak_source": src + "; the float path contracts an fp16 copy with kind::f16 (integer-valued SIFT data: exact; arbitrary values: two "
                                     "ranking passes + exact refinement, counted once); algorithmic FLOPs = 2*Nq*Nt*128 per pair"}
    else:
        peak = n_sm * 128 * 2 * sm_max_mhz * 1e6 / 1e12  # fp32 FMA lanes: exact mode runs on CUDA cores
        roof = {"bound": "fp32", "achieved": knn_work / knn_s / 1e12, "peak": peak, "unit": "TFLOP/s",
                "peak_source": f"128 FFMA lanes/SM x {n_sm} SMs x {sm_max_mhz:.0f} MHz; algorithmic FLOPs = 2*Nq*Nt*128"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["kernel_ms_per_launch"] = knn_ms / max(knn_launches, 1)
    roof["kernel_share_of_step"] = knn_ms / max(step_ms_sum, 1e-9)
    roof["traffic"] = None
    return roof


# ------------------------------------------------------------------------------ one workload
class Env:
    pass


def measure(env, args, name, kind, n_images, n_desc, *, steps, warmup, e2e_steps, e2e_warmup, cross=False, float_mode=0, binary_engine=0,
            alt=False, verify=0, clocks=False, device_only_iters=3):
    """Runs one workload on env.world ranks; returns the measurements on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from sfm_danpipeline_b200 import Matcher
    from sfm_danpipeline_b200 import distributed as D
    world, rank, local, dev = env.world, env.rank, env.local, env.dev
    norm = 0 if kind in ("binary", "orb", "golden:temple_akaze") else 1
    descs = make_descriptors(kind, n_images, n_desc, args.seed)[0] if rank == 0 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    m = Matcher(norm, 0.8, cross, device=local, float_mode=float_mode, binary_engine=binary_engine)
    if world > 1:
        D.broadcast_descriptors(m, descs, 0)
    else:
        m.set_descriptors(descs)
    rows = list(m.rows)
    pairs = D.all_pairs(n_images)
    shards = D.shard_pairs(pairs, rows, world)
    mine = pairs[shards[rank]]
    torch.cuda.synchronize()
    acc = {"knn_ms": 0.0, "knn_work": 0.0, "knn_launches": 0, "matches": 0}

    def note_stats():
        st = m.stats()
        acc["knn_ms"] += st["last_knn_ms"]
        acc["knn_work"] += st["last_knn_work"]
        acc["knn_launches"] += st["last_knn_launches"]

    def match_fn(chunk, slot):
        c, mm, k = D.match_shard(m, chunk, rows, slot=slot)
        note_stats()
        acc["matches"] += k
        return c, mm

    def resident_step(gather=None):
        """Descriptors resident -> every match list in rank 0's host memory."""
        gather = gather or args.gather
        if world > 1 and gather == "shared":
            t = D.match_and_share(m, pairs, shards, 0, rows=rows)
            note_stats()
            if not getattr(m, "_shm_unavailable", False):
                acc["matches"] += m.shared_table_info()[1]
            return t
        if world > 1 and gather == "nccl-once":
            c, mm, k = D.match_shard(m, mine, rows)
            note_stats()
            acc["matches"] += k
            return D.gather_results(pairs, shards, c, mm, 0)
        if world > 1:
            return D.match_and_gather(match_fn, pairs, shards, rows, 0)
        m.match_all_pairs()
        note_stats()
        t = m.result_table(copy=False)
        acc["matches"] += len(t[3])
        return t

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_steps(k, fn):
        total = 0.0
        for _ in range(k):
            env.flush.zero_()
            barrier()
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total

    for _ in range(warmup):
        resident_step()
    barrier()
    for k in acc:
        acc[k] = 0
    launches0 = m.stats()["kernel_launches"]
    sampler = ClockSampler(local) if clocks else None
    if sampler:
        sampler.__enter__()
    t0 = time.perf_counter()
    step_ms_sum = timed_steps(steps, resident_step)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    if sampler:
        sampler.__exit__()
    launches = m.stats()["kernel_launches"] - launches0
    res_acc = dict(acc)
    gather_alt = None
    if world > 1 and alt:  # the other forms of the gather step on the same resident descriptors, for the record
        gather_alt = {}
        for mode in ("shared", "nccl", "nccl-once"):
            if mode == args.gather:
                continue
            resident_step(mode)
            t_alt = timed_steps(2, lambda: resident_step(mode)) / 2
            tt = torch.tensor([t_alt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            gather_alt[mode] = {"value": len(pairs) / (tt.item() * 1e-3), "ms_per_step": tt.item()}

    # ---- round-1 definition (match lists left in HBM): continuity + the kernel's share of a pure device step
    dev_ms = []
    kacc = {"knn_ms": 0.0, "knn_work": 0.0, "knn_launches": 0, "step_ms": 0.0}
    for it in range(1 + device_only_iters):
        env.flush.zero_()
        barrier()
        D.match_shard(m, mine, rows)
        if it:
            st = m.stats()
            dev_ms.append(st["last_match_ms"])
            kacc["knn_ms"] += st["last_knn_ms"]
            kacc["knn_work"] += st["last_knn_work"]
            kacc["knn_launches"] += st["last_knn_launches"]
            kacc["step_ms"] += st["last_match_ms"]
    dev_only = float(np.mean(dev_ms)) if dev_ms else 0.0

    stat = torch.tensor([step_ms_sum, float(launches), res_acc["knn_ms"], res_acc["knn_work"], float(res_acc["knn_launches"]),
                         float(res_acc["matches"]), dev_only], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stat.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stat.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        step_ms_max, launches_sum, dev_only_max = mx[0].item(), int(sm[1].item()), mx[6].item()
        total_matches = int(sm[5].item())
    else:
        step_ms_max, launches_sum, dev_only_max, total_matches = step_ms_sum, int(launches), dev_only, int(res_acc["matches"])
    ms_per_step = step_ms_max / steps

    # ---- end to end: host buffers in, host table out
    def e2e_step():
        s0 = m.stats()
        if world > 1:
            table, _ = D.match_all_pairs_distributed(m, descs, 0, gather=args.gather)
        else:
            ta = time.perf_counter()
            m.set_descriptors(descs)
            tb = time.perf_counter()
            m.match_all_pairs()
            tc_ = time.perf_counter()
            table = m.result_table(copy=False)
            assert int(table[1].sum()) == len(table[3])  # the host table is complete
            if os.environ.get("SFMM_BENCH_TRACE") == "1":
                print(f"[e2e] set_descriptors {1e3 * (tb - ta):.2f} ms  match_all_pairs {1e3 * (tc_ - tb):.2f} ms  "
                      f"table {1e3 * (time.perf_counter() - tc_):.2f} ms  device {m.stats()['last_match_ms']:.2f} ms", file=sys.stderr)
        s1 = m.stats()
        return table, s1["h2d_bytes"] - s0["h2d_bytes"], s1["d2h_bytes"] - s0["d2h_bytes"]

    table = None
    for _ in range(e2e_warmup):
        table = None
        table, _a, _b = e2e_step()
    e2e_total = 0.0
    h2d = d2h = 0
    for _ in range(e2e_steps):
        table = None
        env.flush.zero_()
        barrier()
        t0 = time.perf_counter()
        table, h2d, d2h = e2e_step()
        torch.cuda.synchronize()
        e2e_total += (time.perf_counter() - t0) * 1e3
    e2e_ms = e2e_total / max(e2e_steps, 1)
    if world > 1:
        tmax = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_ms = tmax.item()
        if rank == 0:
            d2h = int((table.nbytes if hasattr(table, "parts") else table.matches.nbytes) + table.counts.nbytes)

    out = None
    if rank == 0:
        engine = ("tensor" if m.stats()["float_path"] == 2 else "popc") if norm == 0 else "float"
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        # kernel time: CUDA events around the 2-NN launches of the device-only iterations (one stream, launches back to back; in the
        # host-table pipeline two chunk slots are in flight and an event pair would also cover the wait for the other slot's kernel)
        roof = roofline_block(m.stats()["float_path"], norm, engine, kacc["knn_ms"], kacc["knn_work"], kacc["knn_launches"],
                              kacc["step_ms"], n_sm, env.peaks, env.tc, tensor_kind=int(m.stats().get("tensor_kind", 0)))
        roof["kernel_share_of_step"] = min(1.0, kacc["knn_ms"] / device_only_iters / max(ms_per_step, 1e-9))
        roof["kernel_share_of_device_only_step"] = kacc["knn_ms"] / max(kacc["step_ms"], 1e-9)
        key = f"{name}/{engine}/{'cross' if cross else 'plain'}"
        if key in env.traffic and n_images == WORKLOADS[name][1]:
            # measured by ncu per pair (one launch of the same kernel on the same workload), scaled to this run's pairs per launch
            roof["traffic"] = env.traffic[key]["bytes_per_pair"] * len(mine) * device_only_iters / max(kacc["knn_launches"], 1)
            roof["traffic_source"] = env.traffic[key]["source"]
        else:
            roof["traffic_source"] = "no ncu --set full capture of this exact workload is committed: null"
        row_bytes = m.cols * (1 if norm == 0 else 4)
        alg_bytes = (float(sum((rows[q] + rows[t]) * row_bytes for q, t in mine)) + 16.0 * res_acc["matches"] / max(steps, 1)) * device_only_iters
        roof["algorithmic_bytes_per_launch"] = alg_bytes / max(kacc["knn_launches"], 1)
        roof["hbm"] = {"algorithmic_GBps": alg_bytes / max(kacc["knn_ms"] * 1e-3, 1e-12) / 1e9, "peak_GBps": env.peaks.get("hbm_gbs"),
                       "note": "compute-bound path: HBM is reported, not the binding roof"}
        out = {"value": len(pairs) / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "wall_ms_per_step": wall_ms / steps,
               "pairs": int(len(pairs)), "pairs_per_gpu": int(len(mine)), "matches_per_step": total_matches // max(steps, 1),
               "engine": engine, "norm": norm, "workload": workload_string(name, kind, n_images, n_desc, cross),
               "e2e": {"value": len(pairs) / (e2e_ms * 1e-3), "unit": "pairs/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d),
                       "d2h_bytes_per_step": int(d2h)},
               "resident_device_only": {"value": len(pairs) / (dev_only_max * 1e-3) if dev_only_max else None, "ms_per_step": dev_only_max,
                                        "note": "round-1 definition of `value`: match lists left in HBM (sfmm_match_pairs_device), library CUDA events"},
               "gpu_launches": launches_sum, "roofline": roof, "clocks": sampler.summary() if sampler else None,
               "gather_chunks": D.gather_chunks(pairs, shards, rows) if world > 1 else None, "gather_alt": gather_alt,
               "gather_used": ("nccl (shared-memory tables did not fit in /dev/shm)" if getattr(m, "_shm_unavailable", False) else args.gather)}
        if verify:
            get = table.getMatching if world > 1 else (lambda q, t: m.getMatching(q, t))
            out["verified"] = verify_pairs(get, pairs, descs, norm, verify, cross)
    if alt and norm == 0 and m.cols <= 64 and not cross:
        # same workload, same run, through the other Hamming engine (bit-identical tables): the north-star XOR+POPC kernel
        # is reported beside the tensor engine AUTO picks, with its own POPC-pipe roofline (rank 0's shard at N>1)
        eng_now = "tensor" if m.stats()["float_path"] == 2 else "popc"
        other = "popc" if eng_now == "tensor" else "tensor"
        table = None
        m.close()
        m = None
        if rank == 0:
            m2 = Matcher(0, 0.8, False, device=local, binary_engine=1 if other == "popc" else 2)
            try:
                m2.set_descriptors(descs)
                ms, knn, work, nl = [], 0.0, 0.0, 0
                for it in range(1 + 2):
                    env.flush.zero_()
                    torch.cuda.synchronize()
                    _c, _m, n = D.match_shard(m2, mine, rows)
                    if it >= 1:
                        st = m2.stats()
                        ms.append(st["last_match_ms"])
                        knn += st["last_knn_ms"]
                        work += st["last_knn_work"]
                        nl += st["last_knn_launches"]
                per = float(np.mean(ms))
                n_sm = torch.cuda.get_device_properties(local).multi_processor_count
                r2 = roofline_block(m2.stats()["float_path"], 0, other, knn, work, nl, per * len(ms), n_sm, env.peaks, env.tc, tensor_kind=int(m2.stats().get("tensor_kind", 0)))
                out["alt_engine"] = {"binary_engine": "tensor (tcgen05 on unpacked bits)" if other == "tensor" else "popc (XOR + carry-save + POPC, packed bits)",
                                     "pairs_per_s_per_gpu": len(mine) / (per * 1e-3), "ms_per_step": per, "matches_per_step_this_rank": int(n),
                                     "definition": "resident_device_only (match lists left in HBM)", "roofline": r2,
                                     "note": "same workload, same run, identical match tables; selectable with SfmmConfig.binary_engine"}
            finally:
                m2.close()
        if world > 1:
            dist.barrier()
    if m is not None:
        table = None
        m.close()
    return out, descs, norm


def orb_block(local):
    """The next row of the path: ORB extraction on the GPU vs cv::ORB on the host cores, on the reference's
    own fixture (the ten 640 x 480 data/temple images, committed as a golden), with the parity check in the same run."""
    from sfm_danpipeline_b200 import OrbExtractor
    z = np.load(os.path.join(ROOT, "tests", "golden", "temple_orb_features.npz"))
    imgs = [np.ascontiguousarray(x) for x in z["images"]]
    offs = np.concatenate([[0], np.cumsum(z["counts"])])
    out = {"workload": "ORB(500, 1.2, 8, 31, 0, 2, HARRIS, 31, 20) on the 10 data/temple images (640 x 480), src/Sfm.cpp:358-384"}
    with OrbExtractor(local) as orb:
        for im in imgs:
            orb.detectAndCompute(im)  # warm-up: buffers, taps
        ok = True
        dev_ms = 0.0
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            for i, im in enumerate(imgs):
                kp, d = orb.detectAndCompute(im)
                dev_ms += orb.stats()["last_ms"]
        wall = time.perf_counter() - t0
        for i, im in enumerate(imgs):  # parity: keypoint set + bit-exact descriptors vs the cv2 golden
            kp, d = orb.detectAndCompute(im)
            ref = {(int(k[5]), float(k[0]), float(k[1])): (float(k[3]), float(k[4]), bytes(dd)) for k, dd in zip(z["keypoints"][offs[i]:offs[i + 1]], z["descriptors"][offs[i]:offs[i + 1]])}
            got = {(int(k["octave"]), float(k["x"]), float(k["y"])): (float(k["angle"]), float(k["response"]), bytes(dd)) for k, dd in zip(kp, d)}
            ok &= got == ref
        out.update(images_per_s=reps * len(imgs) / wall, ms_per_image_e2e=1e3 * wall / (reps * len(imgs)), ms_per_image_device=dev_ms / (reps * len(imgs)),
                   kernel_launches_per_image=orb.stats()["kernel_launches"] // ((reps + 2) * len(imgs)),
                   verified="keypoint sets equal, angles / responses / descriptors bit-identical to the cv2 golden" if ok else "MISMATCH")
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count() or 1)
        ref = cv2.ORB_create(500, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 20)
        ref.detectAndCompute(imgs[0], None)
        t0 = time.perf_counter()
        for _ in range(3):
            for im in imgs:
                ref.detectAndCompute(im, None)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 3 * len(imgs) / dt, "unit": "images/s", "cores": cv2.getNumThreads(), "kind": "reference",
                               "sample": f"cv2 {cv2.__version__} ORB.detectAndCompute, 30 images, {dt:.2f} s"}
    except Exception:
        pass
    return out


# ------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--images", type=int, default=0)
    ap.add_argument("--desc", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cross-check", action="store_true")
    ap.add_argument("--float-mode", default="auto", choices=["auto", "exact", "tensor"])
    ap.add_argument("--binary-engine", default="auto", choices=["auto", "popc", "tensor"],
                    help="auto = the library default (tensor engine for <= 512-bit descriptors); popc = XOR+CSA+POPC kernel "
                         "(the north-star design); tensor = tcgen05 engine (FP4 pipe below 512 bit, kind::i8 at 512).  The other engine is reported as alt_engine")
    ap.add_argument("--no-alt-engine", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs1_cfg2 and float blocks of the N=1 line")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--e2e-warmup", type=int, default=2,
                    help="untimed end-to-end steps: the pipeline slots size their device/pinned buffers over the first two")
    ap.add_argument("--gather", default="shared", choices=["shared", "nccl", "nccl-once"],
                    help="N>1: how the match lists reach rank 0's host memory -- shared: every GPU copies its records over its own PCIe link "
                         "into a shared page-locked table rank 0 maps; nccl: chunk-wise NCCL gather to rank 0's GPU + D2H there, overlapped "
                         "with matching; nccl-once: one NCCL gather at the end (round 1).  The other two are reported as gather_alt")
    ap.add_argument("--device-only-iters", type=int, default=3, help="iterations of the device-only (round-1 definition) measurement behind the roofline")
    ap.add_argument("--verify", type=int, default=2, help="check this many random pairs of the e2e table against the CPU oracle (0 = off)")
    args = ap.parse_args()

    kind, n1, n_desc, weak = WORKLOADS[args.workload]
    n_desc = args.desc or n_desc
    n_images = args.images or images_for(n1, args.gpus, weak)
    if args.impl == "reference":
        return run_reference_arm(args, args.workload, kind, n_images, n_desc)

    import torch
    import torch.distributed as dist

    env = Env()
    env.world = int(os.environ.get("WORLD_SIZE", "1"))
    env.rank = int(os.environ.get("RANK", "0"))
    env.local = int(os.environ.get("LOCAL_RANK", "0"))
    if env.world != args.gpus and env.world == 1 and args.gpus > 1:
        raise SystemExit("--gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
    torch.cuda.set_device(env.local)
    if env.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", env.local))
    env.dev = torch.device("cuda", env.local)
    env.flush = torch.empty(256 << 20, dtype=torch.uint8, device=env.dev)  # > 126 MB L2
    env.peaks, env.tc, env.traffic = load_peaks()

    fm = {"auto": 0, "exact": 1, "tensor": 2}[args.float_mode]
    be = {"auto": 0, "popc": 1, "tensor": 2}[args.binary_engine]
    e2e_steps = args.e2e_steps or max(2, min(args.steps, 3))
    res, descs, norm = measure(env, args, args.workload, kind, n_images, n_desc, steps=args.steps, warmup=args.warmup, e2e_steps=e2e_steps,
                               e2e_warmup=args.e2e_warmup, cross=args.cross_check, float_mode=fm, binary_engine=be,
                               alt=not args.no_alt_engine, verify=args.verify, clocks=True, device_only_iters=max(1, args.device_only_iters))
    if env.rank == 0:
        line = {
            "metric": "image-pair matches/sec (all-pairs 2-NN + ratio test)", "value": res["value"], "unit": "pairs/s",
            "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
            "dtype": "u8" if norm == 0 else "f32", "data": "synthetic" if not kind.startswith("golden:") else "cv2 descriptors of the reference's data/temple fixture",
            "config": {"workload": res["workload"], "engine": res["engine"], "pairs": res["pairs"], "pairs_per_gpu": res["pairs_per_gpu"],
                       "matches_per_step": res["matches_per_step"],
                       "timed_region": "descriptors resident in HBM on every rank -> all match lists in rank 0's host memory "
                                       "(matching + gather + device->host), CUDA events per step, max over ranks",
                       "l2": "flushed between steps (256 MiB memset outside the timed events); the operand set is also larger than L2",
                       "parallelism": f"pairs sharded over {env.world} rank(s) by cost-sorted snake deal"
                                      + (f"; NCCL broadcast of the descriptors (e2e only); gather={res['gather_used']}" if env.world > 1 else "")},
            "wall_ms_per_step": res["wall_ms_per_step"],
            "e2e": res["e2e"], "resident_device_only": res["resident_device_only"],
            "gpu_launches": res["gpu_launches"], "roofline": res["roofline"], "clocks": res["clocks"],
        }
        for k in ("alt_engine", "verified", "gather_alt"):
            if k in res:
                line[k] = res[k]
        if env.world == 1 and not args.no_cpu_baseline:
            v, kind_s, cores, how, _dt = time_cpu_sample(descs, norm, args.cpu_seconds, 100000)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind_s, "sample": how}
    descs = None
    if env.world == 1 and not args.no_extra and args.workload == DEFAULT_WORKLOAD and not args.cross_check:
        # BASELINE.json configs[1] and the float (SIFT) shape, measured the same way in the same run
        k2, n2, d2, _w = WORKLOADS["cfg2"]
        r2, _d, _n = measure(env, args, "cfg2", k2, n2, d2, steps=max(3, min(args.steps, 10)), warmup=3, e2e_steps=3, e2e_warmup=2, verify=1)
        line["configs1_cfg2"] = {k: r2[k] for k in ("workload", "engine", "value", "ms_per_step", "e2e", "resident_device_only", "roofline", "verified")}
        k4, n4, d4, _w = WORKLOADS["cfg4s"]
        r4, _d, _n = measure(env, args, "cfg4s", k4, n4, d4, steps=3, warmup=3, e2e_steps=2, e2e_warmup=2, verify=1)
        line["float"] = {k: r4[k] for k in ("workload", "engine", "value", "ms_per_step", "e2e", "resident_device_only", "roofline", "verified")}
        line["orb_extraction"] = orb_block(env.local)
    if env.rank == 0:
        print(json.dumps(line), flush=True)
    if env.world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

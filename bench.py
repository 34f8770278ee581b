#!/usr/bin/env python
"""bench.py -- image-pair matches/sec of the all-pairs 2-NN matching path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg5|...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU BFMatcher path on the host cores

A "step" is one pass of the hot path over the whole workload: every image pair q<t of the
synthetic descriptor set (findBestPair's loop, /root/reference/src/Sfm.cpp:511-515) goes through
2-NN + ratio test (getMatching, src/Sfm.cpp:590-608).  One "pair-match" = one such pair.

  value  : pair-matches/s with descriptors already resident in HBM, results left in HBM
           (sfmm_match_pairs_device), device-timed with CUDA events, max over ranks.
  e2e    : the same metric through the public API with HOST buffers: H2D of the descriptors,
           (NCCL broadcast), matching, (NCCL gather to rank 0), D2H of the match table.
  roofline: the 2-NN kernel that actually ran against its bound: the tensor pipe (default engines, `of measured` =
           MEASURED_PEAKS.json bf16 scaled to the operand type) or the POPC pipe / FP32 lanes (--binary-engine popc,
           --float-mode exact).  alt_engine: the other Hamming engine on the same shard in the same run.
  cpu_baseline: the reference's own CPU path (OpenCV BFMatcher via cv2, else the C oracle) on a
           bounded sample of the same pairs, timed on this box's host cores (N=1, rank 0).

Default workload at N GPUs: AKAZE-shape 486-bit descriptors, 5000 per image (BASELINE.json
configs[1]); the image count grows with N so that every GPU keeps ~1225 pairs (weak scaling).
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
    "cfg2": ("binary", 50, 5000, True),     # configs[1]: 50 x 5k x 486 bit, 1 GPU
    "cfg3": ("binary", 200, 10000, False),  # configs[2]: 200 x 10k, sharded over 2/4/8
    "cfg4": ("float", 300, 8000, False),    # configs[3]: 300 x 8k x 128 f32
    "cfg5": ("binary", 1000, 20000, False), # configs[4]: headline, ~500k pairs
    "cfg5s": ("binary", 40, 20000, True),   # cfg5's pair shape (20k x 20k) on a 40-image subset
    "orb": ("orb", 100, 2000, True),        # ORB shape: 256 bit
    "float2": ("float", 40, 4000, True),
    "cfg4s": ("float", 60, 8000, True),     # cfg4's pair shape (8k x 8k x 128 f32) on a 60-image subset
    "cfg4x": ("floatx", 60, 8000, True),    # same shape, NON-integer values: TF32 ranking + exact refinement path
    # configs[0]: the reference's own fixture (data/temple, 10 images, 45 pairs) through the committed cv2 descriptors
    "temple_sift": ("golden:temple_sift", 10, 850, False),
    "temple_akaze": ("golden:temple_akaze", 10, 700, False),
}


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
    if kind in ("binary", "orb") and n_images * n_desc >= 4_000_000:
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
    return done / dt, kind, cores, f"{done} of {len(pairs)} pairs ({how}), {dt:.1f} s"


def run_reference_arm(args, kind, n_images, n_desc):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    descs, norm = make_descriptors(kind, min(n_images, 24), n_desc, args.seed)
    per_step = []
    how = cores = kind_s = None
    for s in range(args.warmup + args.steps):
        v, kind_s, cores, how = time_cpu_sample(descs, norm, args.cpu_seconds / max(args.steps, 1), 100000, seed=s)
        if s >= args.warmup:
            per_step.append(v)
    value = float(np.mean(per_step))
    line = {"impl": "reference", "metric": "image-pair matches/sec (all-pairs 2-NN + ratio test)", "value": value,
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8" if norm == 0 else "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {kind} {n_desc} descriptors/image; each step = a bounded random sample of the pairs"},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind_s, "sample": how},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def verify_pairs(table, descs, norm, n, cross, world):
    """Bit-exactness spot check of the end-to-end table at full size: n random pairs vs the CPU oracle."""
    import oracle
    rng = np.random.default_rng(123)
    if world > 1:
        pairs, get = table.pairs, table.getMatching
    else:
        pairs = table[0]
        pos = {(int(q), int(t)): i for i, (q, t) in enumerate(pairs)}
        get = lambda q, t: table[3][table[2][pos[(q, t)]]: table[2][pos[(q, t)]] + table[1][pos[(q, t)]]]  # noqa: E731
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
    res = {"pairs_checked": ok, "against": "cv2 BFMatcher path" if oracle.have_cv2() else "C oracle",
           "result": "bit-identical" if integer_valued else f"within 1e-4 relative (worst {worst:.2e})"}
    return res


def alt_engine_line(args, local, world, rank, descs, mine, rows, n_pairs, flush, D, engine, n_sm, sm_max_mhz, peaks):
    """Resident throughput of the OTHER Hamming engine on the same shard (rank 0's shard at N>1), with its own roofline."""
    import torch
    from sfm_danpipeline_b200 import BINARY_POPC, BINARY_TENSOR, Matcher
    m2 = Matcher(0, 0.8, False, device=local, binary_engine=BINARY_TENSOR if engine == "tensor" else BINARY_POPC)
    try:
        m2.set_descriptors(descs)
        ms, knn, work = [], 0.0, 0.0
        for it in range(2 + 3):
            flush.zero_()
            torch.cuda.synchronize()
            _c, _m, n = D.match_shard(m2, mine, rows)
            if it >= 2:
                st = m2.stats()
                ms.append(st["last_match_ms"])
                knn += st["last_knn_ms"]
                work += st["last_knn_work"]
        per = float(np.mean(ms))
        out = {"binary_engine": "tensor (tcgen05 kind::i8 on unpacked bits)" if engine == "tensor" else "popc (XOR + carry-save + POPC, packed bits)",
               "pairs_per_s_per_gpu": len(mine) / (per * 1e-3), "ms_per_step": per, "matches_per_step_this_rank": int(n),
               "note": "same workload, same run, identical match tables; selectable with SfmmConfig.binary_engine"}
        if engine == "popc":
            peak = n_sm * 16 * sm_max_mhz * 1e6 / 1e9
            out["roofline"] = {"bound": "popc", "achieved": work / (knn * 1e-3) / 1e9, "peak": peak, "unit": "GPOPC32/s",
                               "frac": work / (knn * 1e-3) / 1e9 / peak,
                               "peak_source": f"nominal 16 POPC/clk/SM x {n_sm} SMs x {sm_max_mhz:.0f} MHz"}
        else:
            i8 = 2.0 * float(peaks.get("bf16_tflops", 1590.0))
            ach = 2.0 * (work / 16.0 * 512.0) / (knn * 1e-3) / 1e12
            out["roofline"] = {"bound": "tensor", "achieved": ach, "peak": i8, "unit": "TOP/s", "frac": ach / i8,
                               "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops"}
        return out
    finally:
        m2.close()


# ------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--images", type=int, default=0)
    ap.add_argument("--desc", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cross-check", action="store_true")
    ap.add_argument("--float-mode", default="auto", choices=["auto", "exact", "tensor"])
    ap.add_argument("--binary-engine", default="auto", choices=["auto", "popc", "tensor"],
                    help="auto = the library default (tensor engine for <= 512-bit descriptors); popc = XOR+CSA+POPC kernel "
                         "(the north-star design); tensor = tcgen05 kind::i8 engine.  The other engine is reported as alt_engine")
    ap.add_argument("--no-alt-engine", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--e2e-warmup", type=int, default=2,
                    help="untimed end-to-end steps: the two pipeline slots size their device/pinned buffers over the first two")
    ap.add_argument("--verify", type=int, default=0, help="check this many random pairs of the e2e table against the CPU oracle")
    args = ap.parse_args()

    kind, n1, n_desc, weak = WORKLOADS[args.workload]
    n_desc = args.desc or n_desc
    n_images = args.images or images_for(n1, args.gpus, weak)
    if args.impl == "reference":
        return run_reference_arm(args, kind, n_images, n_desc)

    import torch
    import torch.distributed as dist
    from sfm_danpipeline_b200 import FLOAT_AUTO, Matcher
    from sfm_danpipeline_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # synthetic descriptors exist on rank 0 only; the other ranks receive them by NCCL broadcast
    norm = 0 if kind in ("binary", "orb", "golden:temple_akaze") else 1
    descs = make_descriptors(kind, n_images, n_desc, args.seed)[0] if rank == 0 else None
    dev = torch.device("cuda", local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    m = Matcher(norm, 0.8, args.cross_check, device=local, float_mode={"auto": 0, "exact": 1, "tensor": 2}[args.float_mode],
                binary_engine={"auto": 0, "popc": 1, "tensor": 2}[args.binary_engine])
    # ---- resident arm: descriptors in HBM before the timed region -------------------------
    if world > 1:
        D.broadcast_descriptors(m, descs, 0)
    else:
        m.set_descriptors(descs)
    rows = list(m.rows)
    pairs = D.all_pairs(n_images)
    shards = D.shard_pairs(pairs, rows, world)
    mine = pairs[shards[rank]]
    torch.cuda.synchronize()

    def resident_step():
        flush.zero_()
        torch.cuda.synchronize()
        _c, _m, n = D.match_shard(m, mine, rows)
        return n, m.stats()

    for _ in range(args.warmup):
        resident_step()
    barrier()
    launches0 = m.stats()["kernel_launches"]
    dev_ms = knn_ms = knn_work = 0.0
    knn_launches = 0
    n_matches = 0
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            n_matches, st = resident_step()
            dev_ms += st["last_match_ms"]
            knn_ms += st["last_knn_ms"]
            knn_work += st["last_knn_work"]
            knn_launches += st["last_knn_launches"]
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
    launches = m.stats()["kernel_launches"] - launches0
    stat = torch.tensor([dev_ms, float(launches), knn_ms, knn_work, float(knn_launches), float(n_matches)],
                        dtype=torch.float64, device=dev)
    if world > 1:
        mx = stat.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stat.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms_max, launches_sum, total_matches = mx[0].item(), int(sm[1].item()), int(sm[5].item())
    else:
        dev_ms_max, launches_sum, total_matches = dev_ms, int(launches), int(n_matches)
    ms_per_step = dev_ms_max / args.steps
    value = len(pairs) / (ms_per_step * 1e-3)

    # ---- end-to-end arm: host buffers in, host table out ----------------------------------
    def e2e_step():
        s0 = m.stats()
        if world > 1:
            table, _ = D.match_all_pairs_distributed(m, descs, 0)
        else:
            ta = time.perf_counter()
            m.set_descriptors(descs)
            tb = time.perf_counter()
            m.match_all_pairs()
            tc = time.perf_counter()
            table = m.result_table(copy=False)
            assert int(table[1].sum()) == len(table[3])  # the host table is complete
            if os.environ.get("SFMM_BENCH_TRACE") == "1":
                print(f"[e2e] set_descriptors {1e3 * (tb - ta):.2f} ms  match_all_pairs {1e3 * (tc - tb):.2f} ms  "
                      f"table {1e3 * (time.perf_counter() - tc):.2f} ms  device {m.stats()['last_match_ms']:.2f} ms", file=sys.stderr)
        s1 = m.stats()
        return table, s1["h2d_bytes"] - s0["h2d_bytes"], s1["d2h_bytes"] - s0["d2h_bytes"]

    for _ in range(args.e2e_warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = args.e2e_steps or max(2, min(args.steps, 3))
    for _ in range(e2e_steps):
        flush.zero_()
        table, h2d, d2h = e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        tmax = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_ms = tmax.item()
        if rank == 0:
            d2h = int(table.matches.nbytes + table.counts.nbytes)
    e2e_value = len(pairs) / (e2e_ms * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        knn_s = knn_ms * 1e-3
        # the engine the library actually ran (AUTO resolves to the tensor engine for <= 512-bit descriptors)
        engine = ("tensor" if m.stats()["float_path"] == 2 else "popc") if norm == 0 else "float"
        if engine == "tensor":
            i8 = 2.0 * float(peaks.get("bf16_tflops", 1590.0))  # int8 dense = twice the measured bf16 cuBLAS rate
            macs = knn_work / 16.0 * 512.0  # 512 u8 MACs per 16 algorithmic POPC32 (one 486-bit distance)
            roof = {"bound": "tensor", "achieved": 2.0 * macs / knn_s / 1e12, "peak": i8, "unit": "TOP/s",
                    "peak_source": "of measured: 2 x MEASURED_PEAKS.json bf16_tflops (int8 dense is nominally twice bf16; the cuBLAS bf16 burst "
                                   "figure is ~74 % of nominal, so a kernel on {0,1} bytes can read above 1.0); ops = 2*Nq*Nt*512 per pair, "
                                   "tcgen05 kind::i8 on bits unpacked to bytes",
                    "nominal": {"peak": 4500.0, "frac": 2.0 * macs / knn_s / 1e12 / 4500.0, "note": "B200 dense int8/fp8 nominal 4.5 POP/s"},
                    "popc_equivalent": {"achieved_GPOPC32": knn_work / knn_s / 1e9,
                                        "x_nominal_popc_roofline": knn_work / knn_s / 1e9 / (n_sm * 16 * sm_max_mhz * 1e6 / 1e9)},
                    "traffic": None}
        elif norm == 0:
            peak = n_sm * 16 * sm_max_mhz * 1e6 / 1e9  # GPOPC32/s: 16 POPC/clk/SM x SMs x max SM clock
            roof = {"bound": "popc", "achieved": knn_work / knn_s / 1e9, "peak": peak, "unit": "GPOPC32/s",
                    "peak_source": f"nominal 16 POPC/clk/SM x {n_sm} SMs x {sm_max_mhz:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz); "
                                   "measured issue rates in profiles/pipe_bench_r01.txt",
                    "traffic": None}
        elif m.stats()["float_path"] in (2, 3):
            tf32 = 0.5 * float(peaks.get("bf16_tflops", 1590.0))  # TF32 dense = half the measured bf16 cuBLAS rate
            roof = {"bound": "tensor", "achieved": knn_work / knn_s / 1e12, "peak": tf32, "unit": "TFLOP/s",
                    "peak_source": "of measured: 0.5 x MEASURED_PEAKS.json bf16_tflops = the TF32 rate the float path is specified in "
                                   "(north_star: fp32-accurate TF32).  Integer-valued (SIFT) data is contracted from an exact fp16 copy "
                                   "with kind::f16 at twice that rate, so this fraction can exceed 1.0; algorithmic FLOPs = 2*Nq*Nt*128 per pair",
                    "f16_frac": knn_work / knn_s / 1e12 / float(peaks.get("bf16_tflops", 1590.0)),
                    "traffic": None}
        else:
            peak = n_sm * 128 * 2 * sm_max_mhz * 1e6 / 1e12  # fp32 FMA lanes: exact mode runs on CUDA cores
            roof = {"bound": "fp32", "achieved": knn_work / knn_s / 1e12, "peak": peak, "unit": "TFLOP/s",
                    "peak_source": f"128 FFMA lanes/SM x {n_sm} SMs x {sm_max_mhz:.0f} MHz; algorithmic FLOPs = 2*Nq*Nt*128",
                    "traffic": None}
        # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
        # `ncu --set full` captures of exactly this workload (profiles/ncu_*_r01*.txt); null when not captured
        ncu_traffic = {("cfg2", "popc", 1): 18.847e6 + 46.489e6,    # profiles/ncu_binary_r01c_tq2.txt
                       ("cfg2", "tensor", 1): 2276.5e6 + 86.7e6,    # profiles/ncu_tensor_ts_i8p_r01.txt
                       ("cfg5s", "tensor", 1): 4100.7e6 + 117.8e6}  # profiles/ncu_tensor_ts_i8p_cfg5s_r01.txt (one of the step's two launches)
        if n_images == WORKLOADS[args.workload][1] and not args.cross_check:
            roof["traffic"] = ncu_traffic.get((args.workload, engine, world))
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["kernel_ms_per_launch"] = knn_ms / max(knn_launches, 1)
        roof["kernel_share_of_step"] = knn_ms / max(dev_ms, 1e-9)
        row_bytes = m.cols * (1 if norm == 0 else 4)
        alg_bytes = float(sum((rows[q] + rows[t]) * row_bytes for q, t in mine)) * args.steps + 16.0 * n_matches * args.steps
        roof["hbm"] = {"algorithmic_GBps": alg_bytes / knn_s / 1e9, "peak_GBps": peaks.get("hbm_gbs"),
                       "note": "compute-bound path: HBM is reported, not the binding roof"}
        desc_shape = ("486-bit AKAZE" if kind in ("binary", "golden:temple_akaze") else ("256-bit ORB" if kind == "orb" else "128-d f32 SIFT")) \
            + ("-shape" if not kind.startswith("golden:") else " (real, about that many rows per image)")
        line = {
            "metric": "image-pair matches/sec (all-pairs 2-NN + ratio test)", "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
            "dtype": "u8" if norm == 0 else "f32", "data": "synthetic" if not kind.startswith("golden:") else "cv2 descriptors of the reference's data/temple fixture",
            "config": {"workload": f"{args.workload}: {n_images} images x {n_desc} x " + desc_shape +
                                   f" descriptors, all {len(pairs)} pairs q<t, ratio 0.8, cross_check={bool(args.cross_check)}"
                                   + (f", binary_engine={engine}" if norm == 0 else ""),
                       "pairs": int(len(pairs)), "pairs_per_gpu": int(len(mine)), "matches_per_step": total_matches,
                       "l2": "flushed between steps (256 MiB memset outside the timed events)",
                       "parallelism": f"pairs sharded over {world} rank(s); NCCL broadcast + gather only in e2e"},
            "wall_ms_per_step": wall_ms / args.steps,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": launches_sum,
            "roofline": roof,
            "clocks": clk.summary(),
        }
        if norm == 0 and m.cols <= 64 and not args.cross_check and not args.no_alt_engine:
            # same workload, same run, through the other Hamming engine (bit-identical tables): the north-star
            # XOR+POPC kernel is reported beside the tensor engine AUTO picks, with its own POPC-pipe roofline
            line["alt_engine"] = alt_engine_line(args, local, world, rank, descs, mine, rows, len(pairs), flush, D,
                                                 "popc" if engine == "tensor" else "tensor", n_sm, sm_max_mhz, peaks)
        if args.verify:
            line["verified"] = verify_pairs(table, descs, norm, args.verify, args.cross_check, world)
        if world == 1 and not args.no_cpu_baseline:
            v, kind_s, cores, how = time_cpu_sample(descs, norm, args.cpu_seconds, 100000)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind_s, "sample": how}
        print(json.dumps(line), flush=True)
    m.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the trie-mass hot path (BASELINE.json metric: trie weight_sum/max distributions/sec at 128k vocab).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = batch_weight_sum + batch_weight_max over one batch of 64 synthetic Dirichlet rows on the
128,256-token synthetic byte vocabulary (BASELINE.json configs[1]); a "distribution" is one row put through
both reductions.  Under torchrun every rank runs the same per-GPU workload on its own rows (weak scaling,
no collective on the data path); `value` is the whole-job rate, timed on the device, max over ranks.

  value      device-resident: inputs in HBM, CUDA events around exactly K steps (CUDA-graph replays of
             one step per buffer set), rotating over buffer sets larger than L2
  e2e        the same step through the public, reference-shaped API: pinned HOST rows in, numpy arrays
             out (ParallelTokenCharacterTrie.batch_weight_sum_max), H2D and D2H inside the timed region
  roofline   dominant kernel (tile_kernel): algorithmic bytes (4V + 4N per distribution) / its CUDA-event time
  cpu_baseline / --impl reference
             the reference's CPU algorithm (oracle C restatement of the numba loops, OpenMP over rows)
             timed on this box's host cores
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "trie weight_sum+weight_max distributions/sec at 128k vocab"
UNIT = "distributions/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vocab", type=int, default=128256)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--alpha", type=float, default=1.0)
    ap.add_argument("--sets", type=int, default=4, help="rotating buffer sets (each 33 MB in + 176 MB out)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--min-ms", type=float, default=50.0, help="the K timed steps are repeated until this much device time has passed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sampler", action="store_true", help="skip the secondary sampler-kernel measurement")
    ap.add_argument("--allgather", action="store_true",
                    help="N > 1 only: also time the optional NCCL all-gather of the node masses (BASELINE config 5's exchange); "
                         "reported separately, never part of the timed step")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(args):
    return (f"batch_weight_sum + batch_weight_max, synthetic byte vocab V={args.vocab}, batch {args.batch} "
            f"Dirichlet({args.alpha:g}) fp32 rows per GPU")


# ---- the reference's CPU algorithm (oracle) -----------------------------------------------------------------
def cpu_rate(args, layout, idx_to_leaf, seconds_target=12.0):
    """Times oracle weight_sum + weight_max (numba loops restated in C) with all host threads on a bounded sample."""
    import oracle
    from genlm_backend_b200.synthetic import dirichlet_rows

    o = oracle.OracleLayout(idx_to_leaf, layout["child_ptr"], layout["child_idx"])
    threads = oracle.max_threads()
    probe = dirichlet_rows(2, args.vocab, alpha=args.alpha, seed=99)
    o.weight_sum(probe, threads=1)  # warm
    t0 = time.perf_counter()
    o.weight_sum(probe, threads=1)
    o.weight_max(probe, threads=1)
    per_row = (time.perf_counter() - t0) / 2
    rows = int(max(threads * 2, min(threads * 64, seconds_target * threads / max(per_row, 1e-6))))
    rows = max(rows, args.batch)
    rows = min(rows, 4096)
    ws = dirichlet_rows(rows, args.vocab, alpha=args.alpha, seed=100)
    best = float("inf")
    for _ in range(2):
        t0 = time.perf_counter()
        o.weight_sum(ws, threads=threads)
        o.weight_max(ws, threads=threads)
        best = min(best, time.perf_counter() - t0)
    return {
        "value": rows / best, "unit": UNIT, "cores": int(o.last_threads), "kind": "port",
        "sample": f"{rows} rows x (sum+max), V={args.vocab}, oracle/trie_oracle.c (numba loops of base.py:346-393 in C, "
                  f"fp64), OpenMP over rows, best of 2, host has {os.cpu_count()} logical cpus",
        "single_thread_rows_per_s": 1.0 / per_row,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from genlm_backend_b200 import TokenCharacterTrie
    from genlm_backend_b200.synthetic import synth_vocab

    trie = TokenCharacterTrie(synth_vocab(args.vocab))  # host builder only: no GPU work in this arm
    t0 = time.perf_counter()
    cb = cpu_rate(args, trie._layout, trie.idx_to_leaf, seconds_target=8.0)
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.batch / cb["value"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "vocab": args.vocab, "batch": args.batch,
                   "note": "reference CPU algorithm (oracle port of the numba path) on host cores; "
                           "a step is a bounded sample of the workload"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ---- clocks ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.loaded = False
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                if self.loaded:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    for bit, name in names.items():
                        if bits & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- secondary measurement: the SMC row op (BASELINE.json configs[3], per-GPU share) ---------------------------------
def sampler_metrics(dev, vocab, peak, rows=512, iters=20):
    """Fused masked logsumexp + categorical draw over `rows` x `vocab` fp32 log-probabilities (one particle per row) for
    the mask kinds an SMC step uses.  Device time from a CUDA graph of 8 launches rotating over two logit buffers that
    are larger than L2 together."""
    import torch

    from genlm_backend_b200 import smc

    gen = torch.Generator(device=dev).manual_seed(0)
    sets = [torch.log_softmax(torch.randn(rows, vocab, device=dev, generator=gen), dim=-1) for _ in range(2)]
    masks = [torch.rand(rows, vocab, device=dev, generator=gen) < 0.5 for _ in range(2)]
    out = {}

    def pack_bits(keep):  # bool [B, V] -> int32 [B, ceil(V/32)] keep-bitmask (bit i of word w = element 32 w + i)
        pad = (-keep.shape[-1]) % 32
        k = torch.nn.functional.pad(keep, (0, pad)).view(keep.shape[0], -1, 32).to(torch.int64)
        w = (k << torch.arange(32, device=keep.device, dtype=torch.int64)).sum(-1)
        return torch.where(w >= 2**31, w - 2**32, w).to(torch.int32)

    per_graph = 8
    cases = (("no_mask", lambda k: None, 0), ("shared_f32_mask", lambda k: masks[0][0].float().log(), 0),
             ("per_row_bit_mask", lambda k: pack_bits(masks[k]), 4 * ((vocab + 31) // 32)), ("per_row_bool_mask", lambda k: masks[k], vocab))
    for name, mk, mask_bytes in cases:
        ms_ = [mk(0), mk(1)]
        for i in range(3):
            smc.masked_logsumexp_sample(sets[i % 2], ms_[i % 2], seed=i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for k in range(per_graph):
                smc.masked_logsumexp_sample(sets[k % 2], ms_[k % 2], seed=k)
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        sec = a.elapsed_time(b) / (per_graph * iters) / 1e3
        bytes_per_row = 4 * vocab + mask_bytes + 8
        out[name] = {"rows_per_s": rows / sec, "us_per_launch": sec * 1e6, "achieved_GBps": rows * bytes_per_row / sec / 1e9,
                     "frac_of_hbm_peak": rows * bytes_per_row / sec / 1e9 / peak}
    return {"kernel": "lse_sample_kernel<float>", "rows": rows, "vocab": vocab, "bound": "hbm",
            "bytes_per_row": "4V (+ the bytes of a per-row mask: V/8 for bits, V for bools) + 8", **out}


# ---- our arm -------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from genlm_backend_b200 import ParallelTokenCharacterTrie, _lib
    from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its banner ("NCCL version ...") to stdout when the first communicator is created; stdout carries
        # exactly one JSON line, so file descriptor 1 points at stderr until that has happened
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    V, B, K, W = args.vocab, args.batch, args.steps, max(args.warmup, 3)
    trie = ParallelTokenCharacterTrie(synth_vocab(V), devices=[local_rank])
    eng = trie._engine
    N = len(trie)
    eng.ensure_device(local_rank)
    info = eng.plan_info()

    base = dirichlet_rows(B, V, alpha=args.alpha, seed=1 + rank)
    nsets = max(1, args.sets)
    ws_sets = [torch.tensor(np.roll(base, k, axis=0)).to(dev) for k in range(nsets)]
    # output slabs as the engine allocates them: [B, N] views of rows padded to whole 128-byte lines
    sum_sets = [eng.alloc_out(B, torch.float32, dev) for _ in range(nsets)]
    max_sets = [eng.alloc_out(B, torch.float32, dev) for _ in range(nsets)]
    set_bytes = B * V * 4 + 2 * B * N * 4
    l2_bytes = torch.cuda.get_device_properties(dev).L2_cache_size

    def step(k, phases=0, ops=("sum", "max")):
        eng.reduce(ws_sets[k], ops, out_sum=sum_sets[k], out_max=max_sets[k], phases=phases)

    # kernels per step: permute kernel + tile kernel (both reductions in one launch) + span kernel
    launches_per_step = 2 + (1 if info["n_span"] > 0 else 0)

    # warm-up (also sets kernel attributes, allocates the scratch) ------------------------------------------------
    for i in range(W):
        step(i % nsets)
    torch.cuda.synchronize()

    # The timed loop replays ONE CUDA graph that holds exactly K consecutive steps of a stream of batches rotating over
    # the buffer sets, as a serving loop issues them.
    def capture(n):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n):
                step(i % nsets)
        return g

    graph = None if args.no_graph else capture(K)

    def run_steps():
        """Enqueue exactly K steps, rotating over the buffer sets."""
        if graph is None:
            for i in range(K):
                step(i % nsets)
        else:
            graph.replay()

    run_steps()
    torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()

    # timed region: the K steps, repeated until at least --min-ms of device time has passed ---------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    run_steps()
    ev1.record()
    torch.cuda.synchronize()
    reps = max(1, int(np.ceil(args.min_ms / max(ev0.elapsed_time(ev1), 1e-3))))
    if world > 1:
        t = torch.tensor([reps], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        reps = int(t.item())
    barrier()
    clocks.loaded = True
    ev0.record()
    for _ in range(reps):
        run_steps()
    ev1.record()
    torch.cuda.synchronize()
    clocks.loaded = False
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1)) / reps
    value = world * B * K / (ms_total / 1e3)

    # per-kernel timing for the roofline (one op, one phase at a time), same buffers ----------------------------------
    def time_phase(phases, ops, iters):
        """Average device time of one launch group, replayed from a CUDA graph that holds 16 launches rotating over the
        buffer sets (so host launch overhead never limits the rate of short kernels)."""
        for i in range(3):
            step(i % nsets, phases, ops)
        torch.cuda.synchronize()
        per_graph = 4 * nsets  # launches per graph: the per-replay launch bubble (~1 us) is spread over 16 launches
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for k in range(per_graph):
                step(k % nsets, phases, ops)
        g.replay()
        torch.cuda.synchronize()
        reps = max(2, iters // per_graph)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / (reps * per_graph)

    clocks.loaded = True
    iters = max(50, min(K, 400))
    ms_tile = time_phase(_lib.GT_FLAG_PHASE_TILE, ("sum",), iters)
    ms_tile_max = time_phase(_lib.GT_FLAG_PHASE_TILE, ("max",), iters)
    ms_tile_both = time_phase(_lib.GT_FLAG_PHASE_TILE, ("sum", "max"), iters)
    ms_span_both = time_phase(_lib.GT_FLAG_PHASE_SPAN, ("sum", "max"), iters)
    ms_permute = time_phase(_lib.GT_FLAG_PHASE_PERMUTE, ("sum",), iters)
    ms_sum_op = time_phase(0, ("sum",), iters)
    clocks.loaded = False

    # end to end through the public API: pinned host rows in, numpy out -----------------------------------------------
    E = args.e2e_steps or min(K, 10)
    host_sets = [torch.tensor(np.roll(base, k, axis=0)).pin_memory() for k in range(2)]
    sums = maxes = None
    for i in range(4):  # warm-up with the timed loop's result-retention pattern: fills the pinned-buffer pool
        sums, maxes = trie.batch_weight_sum_max(host_sets[i % 2])
    barrier()
    clocks.loaded = True
    t0 = time.perf_counter()
    for i in range(E):
        sums, maxes = trie.batch_weight_sum_max(host_sets[i % 2])
        _ = float(sums[0, N - 1]) + float(maxes[B - 1, N - 1])  # results are on the host
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    # the same step when the caller reads K nodes per row instead of the whole slab (gt_gather_nodes): the [B, N] results
    # stay on the GPU, H2D of the rows and D2H of 2 x [B, K] values are inside the timed region
    Ks = 257  # e.g. the 256 byte children + end-of-token child of the node a particle stands on
    ids = torch.tensor(np.random.default_rng(5).integers(0, N, size=(B, Ks)), dtype=torch.int32, device=dev)
    for i in range(3):
        trie.batch_weight_sum_max_at(host_sets[i % 2], ids)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(E):
        s_at, m_at = trie.batch_weight_sum_max_at(host_sets[i % 2], ids)
        _ = float(s_at[0, 0]) + float(m_at[B - 1, Ks - 1])
    torch.cuda.synchronize()
    sparse_s = max_over_ranks(time.perf_counter() - t0)
    clocks.loaded = False
    barrier()
    clocks.stop()

    # optional exchange of BASELINE config 5: all-gather of the [B, N] node masses over NVLink (not on the hot path) -------
    allgather = None
    if world > 1 and args.allgather:
        from genlm_backend_b200.sharding import all_gather_rows

        local = sum_sets[0].contiguous()  # this rank's [B, N] block (the slab's row padding is not sent)
        for _ in range(3):
            full = all_gather_rows(local, world * B)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        a.record()
        for _ in range(reps):
            full = all_gather_rows(local, world * B)
        b.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(a.elapsed_time(b) / reps)
        ok = bool(torch.equal(full[rank * B:(rank + 1) * B], local))
        allgather = {"ms": ms, "rows_per_rank": B, "bytes_received_per_gpu": (world - 1) * B * N * 4,
                     "GBps_per_gpu_in": (world - 1) * B * N * 4 / (ms * 1e-3) / 1e9, "own_block_intact": ok,
                     "api": "genlm_backend_b200.sharding.all_gather_rows -> torch.distributed.all_gather_into_tensor (NCCL)",
                     "note": "one reduction's [B, N] fp32 node masses per rank; reported separately, not part of a step"}
        del full

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    peak, peak_src = peaks()
    bytes_one = B * (4 * V + 4 * N)    # one reduction: rows in, one node array out
    bytes_both = B * (4 * V + 8 * N)   # both reductions from one launch (the step's tile_kernel launch)
    achieved = bytes_both / (ms_tile_both / 1e3) / 1e9
    path_achieved = bytes_both * K / (ms_total / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and V == 128256 and B == 64:
        with open(tpath) as f:
            traffic = json.load(f)["traffic_bytes_per_launch"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": workload_name(args), "vocab": V, "nodes": N, "batch_per_gpu": B, "global_batch": B * world,
            "parallelism": f"rows sharded, {world} independent GPU(s), no collective",
            "l2_policy": f"rotating {nsets} buffer sets of {set_bytes / 1e6:.0f} MB ({nsets * set_bytes / 1e6:.0f} MB total) "
                         f"vs L2 {l2_bytes / 1e6:.0f} MB",
            "launch": (f"one CUDA graph of the {K} steps, replayed {reps}x ({ms_total * reps:.1f} ms timed)" if graph is not None else f"direct launches, {reps} repetitions"),
            "timed_repeats": reps,
            "tile_leaves": info["tile_leaves"], "n_span": info["n_span"],
        },
        "e2e": {
            "value": world * B * E / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * V * 4,
            "d2h_bytes_per_step": 2 * B * N * 4, "steps": E,
            "api": "ParallelTokenCharacterTrie.batch_weight_sum_max(pinned host tensor) -> numpy",
        },
        "e2e_sparse_readout": {
            "value": world * B * E / sparse_s, "unit": UNIT, "nodes_read_per_row": Ks, "h2d_bytes_per_step": B * V * 4,
            "d2h_bytes_per_step": 2 * B * Ks * 4, "steps": E,
            "api": "ParallelTokenCharacterTrie.batch_weight_sum_max_at(pinned host tensor, node_ids) -> numpy [B, K] x 2",
            "note": "not the headline: same kernels, results read through gt_gather_nodes instead of copying the [B, N] slabs",
        },
        "gpu_launches": launches_per_step * K * reps,
        "roofline": {
            "bound": "hbm", "kernel": "tile_kernel<float,4> (both reductions of the step in one launch)", "achieved": achieved,
            "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "bytes_per_launch": bytes_both, "ms_per_launch": ms_tile_both,
            "one_reduction": {"bytes_per_launch": bytes_one, "ms_per_launch": ms_tile,
                              "achieved": bytes_one / (ms_tile / 1e3) / 1e9, "frac": bytes_one / (ms_tile / 1e3) / 1e9 / peak},
            "note": "algorithmic bytes = (4V + 8N) per distribution x batch; kernel timed alone with CUDA events from a CUDA graph "
                    "rotating over the buffer sets; traffic = dram read+write of one launch from the ncu capture under profiles/",
        },
        "path_roofline": {
            "achieved": path_achieved, "peak": peak, "unit": "GB/s", "frac": path_achieved / peak,
            "note": "whole step (permute + tile kernel for both reductions + span kernel) against (4V + 8N) bytes per distribution",
        },
        "kernel_ms": {"permute": ms_permute, "tile_sum": ms_tile, "tile_max": ms_tile_max, "tile_both": ms_tile_both,
                      "span_both": ms_span_both, "sum_op_all_phases": ms_sum_op},
        "clocks": clocks.summary(),
    }
    if allgather is not None:
        line["allgather"] = allgather
    if world == 1 and not args.no_sampler:
        line["sampler"] = sampler_metrics(dev, V, peak)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_rate(args, trie._layout, trie.idx_to_leaf)
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

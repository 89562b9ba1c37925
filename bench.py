#!/usr/bin/env python
"""bench.py — HMM cell·gene steps/s (forward-backward + Viterbi) on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload expr|allele]

A "step" is one pass of the hot path over one batch of synthetic input: every (cell, chromosome)
chain of a [genes x cells] matrix gets its Viterbi path and its state posteriors.  One process per
GPU (torchrun for N > 1), cells sharded across ranks, no data-path collective (the
per-cell chains are independent) -> weak scaling, value = steps of all ranks / max-over-ranks time.

Keys of the JSON line (rank 0):
  value      device-resident throughput (inputs already in HBM), CUDA events, max over ranks
  e2e        the same metric through the C-ABI host entry point (hb_hmm_norm_host /
             hb_hmm_binom_host) with pinned HOST buffers: H2D, layout change, kernels, D2H all timed
  roofline   algorithmic bytes per launch / mean kernel time vs the measured HBM copy bandwidth
  cpu_baseline  the CPU oracle (oracle/hb_oracle.c, "port": the reference is R and cannot run here)
             on a bounded sample of the same workload, all host threads
`--impl reference` times only that CPU arm and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "HMM cell-gene steps/s (fwd-bwd+Viterbi)"
UNIT = "steps/s"
DEV_SIGMA = (2.0, 3.1)
DEV = 1.015658

WORKLOADS = {
    # BASELINE.json configs[3] (north_star target), expression half: 10k cells x 20k genes per GPU
    "expr": dict(name="C4-expr: 10k cells x 20k genes per GPU, 22 chromosome segments, per-cell chains, "
                      "3-state Gaussian, fp64, Viterbi + posteriors",
                 cells=10000, positions=20000, nseg=22, S=3, bytes_per_step=33),
    # allele half: 10k cells x 200k het SNPs per GPU
    # resident input: one packed word (n << 16) | x per position -> 4 + 1 + 16 algorithmic bytes / step
    "allele": dict(name="C4-allele: 10k cells x 200k SNPs per GPU, 22 chromosome segments, per-cell "
                        "chains, 2-state binomial, packed 16-bit counts, Viterbi + posteriors",
                   cells=10000, positions=200000, nseg=22, S=2, bytes_per_step=21),
    # BASELINE.json configs[4]: 200k cells x 20k genes cell-sharded over 8 GPUs = 25k cells per GPU (weak scaling:
    # --gpus N runs N x 25k cells).  The TIMED step is the per-cell chains on the rank's shard PLUS one recursion
    # level of the split-and-retest driver on the sharded cells with every exchange of SURVEY 8e on NCCL
    # (group sizes, exact pooled-mean relay, retest all-reduce, posterior all-gather): value = per-cell steps of
    # all ranks / that time.  The level is dominated by the exact 80-bit pooled mean, so this
    # line reads far below the chain kernel's; `roofline` is the chain kernel's own, timed inside the same steps.
    "c5": dict(name="C5: 200k cells x 20k genes over 8 GPUs (25k cells per GPU), per-cell chains (Viterbi + "
                    "posteriors) + one sharded recursion level with its NCCL exchanges inside the timed step",
               cells=25000, positions=20000, nseg=22, S=3, bytes_per_step=33),
}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash():
    """sha256 (16 hex) over the CUDA sources the chain kernels are built from: profiles/traffic.json records
    the hash of the build its ncu capture was taken on, so a stale traffic figure cannot be printed beside a
    changed kernel."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "honeybadger_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        # device code only: the kernels (*.cuh, hb_kernels.*); hb_api.cu / hb_host_math.h are host-side
        if f.endswith(".cuh") or f in ("hb_kernels.cu", "hb_kernels.h"):
            h.update(f.encode())
            h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()[:16]


_SASS_TXT = None
DOMINANT_KERNEL = {"expr": "norm3_fast_kernel", "allele": "binom2_fast_kernel"}


def kernel_sass_hash(workload):
    """sha256 (16 hex) over the MACHINE CODE of the workload's dominant kernel in the built library (every
    instantiation of it; `cuobjdump -sass`, the listing of a function starts at its own offset 0, so other
    kernels in the library do not move it).  The identity that matters for a traffic figure: same SASS, same
    DRAM traffic on the same workload.  None when cuobjdump is not there."""
    import hashlib
    global _SASS_TXT
    so = os.path.join(ROOT, "honeybadger_b200", "libhb_b200.so")
    if _SASS_TXT is None:
        try:
            _SASS_TXT = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, timeout=180).stdout
        except Exception:
            _SASS_TXT = ""
    txt = _SASS_TXT
    name, h, keep, n = DOMINANT_KERNEL[workload], hashlib.sha256(), False, 0
    for ln in txt.splitlines():
        if "Function :" in ln:
            keep = name in ln
            n += keep
        if keep:
            h.update(ln.strip().encode())
    return h.hexdigest()[:16] if n else None


def ncu_traffic(workload):
    """dram bytes per launch from the committed ncu capture — only if it was taken on THIS kernel: the SASS of
    the dominant kernel hashes to what profiles/traffic.json recorded ("sass"), or, without cuobjdump, the device
    sources do ("src").  Otherwise null (stale capture)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        e = json.load(open(p)).get(workload)
        if not isinstance(e, dict):
            return None
        sass = kernel_sass_hash(workload)
        if (sass is not None and e.get("sass") == sass) or e.get("src") == kernel_source_hash():
            return e.get("bytes")
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active") and t0 - 0.05 <= ts <= t1 + 0.15:
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- synthetic inputs
def device_inputs(kind, cells, positions, seed, dev):
    """Gene-major synthetic matrices generated on the device (SURVEY.md §8d shapes)."""
    import torch
    from honeybadger_b200 import hmm, synth
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    seg = synth.segments(positions)
    clone = torch.multinomial(torch.tensor([0.5, 0.3, 0.2], device=dev), cells, True, generator=g)
    if kind == "expr":
        state = torch.zeros((positions, 3), dtype=torch.float64, device=dev)
        state[seg[6]:seg[7], 0] = 1
        state[seg[9]:seg[10], 0] = -1
        state[seg[12]:seg[13], 1] = -1
        state[seg[13]:(seg[13] + seg[14]) // 2, 1] = -1
        sig = torch.empty(cells, dtype=torch.float64, device=dev).uniform_(*DEV_SIGMA, generator=g)
        # resident layout: gene-major rows padded to 1 KB (hmm.alloc_rows)
        x = hmm.alloc_rows(positions, cells, torch.float64, dev)
        x.copy_(torch.randn((positions, cells), dtype=torch.float64, device=dev, generator=g))
        x.mul_(sig).add_(state[:, clone] * DEV)
        return dict(x=x, seg=seg)
    p = torch.full((positions, 3), 0.45, dtype=torch.float32, device=dev)
    p[seg[9]:seg[10], 0] = 0.1
    p[seg[12]:seg[13], 1] = 0.1
    p[seg[13]:(seg[13] + seg[14]) // 2, 1] = 0.1
    n = hmm.alloc_rows(positions, cells, torch.int32, dev)
    r = hmm.alloc_rows(positions, cells, torch.int32, dev)
    rows = max(1, (1 << 26) // cells)
    for a in range(0, positions, rows):  # chunked: keeps generator temporaries small
        b = min(positions, a + rows)
        shp = (b - a, cells)
        cov = torch.rand(shp, device=dev, generator=g) < 0.108
        # 1 + NegBin(size 0.3, mean 3) as a gamma-Poisson mixture, clipped at 1245
        nb = torch.poisson(torch._standard_gamma(torch.full(shp, 0.3, device=dev)) * 10.0, generator=g)
        nn = torch.where(cov, (1 + nb).clamp(max=1245), torch.zeros_like(nb))
        pp = p[a:b][:, clone]
        rr = torch.binomial(nn, pp, generator=g)
        n[a:b] = nn.to(torch.int32)
        r[a:b] = rr.to(torch.int32)
    return dict(r=r, n=n, seg=seg)


def host_sample(kind, cells, positions, seed):
    """numpy inputs of the same distribution for the CPU arm (gene-major)."""
    from honeybadger_b200 import synth
    if kind == "expr":
        x, seg, _ = synth.expression(positions, cells, seed=seed)
        return dict(x=x, seg=seg)
    r, n, seg, _ = synth.allele(positions, cells, seed=seed)
    return dict(r=r, n=n, seg=seg)


# ----------------------------------------------------------------------------- CPU arm
def cpu_rate(kind, wl, target_s=12.0, threads=0):
    """steps/s of the CPU oracle (all threads) on a bounded sample: full-length chains, few cells."""
    from oracle import oracle as O
    O.build()
    nt = threads or (os.cpu_count() or 1)
    positions = wl["positions"]
    cells = max(nt * 4, 16)
    Pi3, Pi2 = O.sym_Pi(3, 1e-6), O.sym_Pi(2, 1e-6)

    def run(ncell):
        d = host_sample(kind, ncell, positions, seed=7)
        t = time.perf_counter()
        if kind == "expr":
            sd = d["x"].std(axis=0, ddof=1)
            O.fbv_norm_batch(d["x"], d["seg"], [-DEV, 0.0, DEV], sd, Pi3, [0.0, 1.0, 0.0], nthreads=nt)
        else:
            O.fbv_binom_batch(d["r"], d["n"], d["seg"], [0.1, 0.45], Pi2, [0.0, 1.0], nthreads=nt)
        return time.perf_counter() - t

    t_small = run(cells)
    scale = max(1.0, min(64.0, target_s / max(t_small, 1e-3)))
    cells2 = int(cells * scale)
    t = run(cells2) if cells2 > cells else t_small
    ncell = cells2 if cells2 > cells else cells
    return dict(value=ncell * positions / t, unit=UNIT, cores=nt, kind="port",
                sample=f"{ncell} cells x {positions} positions x {wl['nseg']} segments, full-length "
                       f"chains, oracle/hb_oracle.c (CPU restatement of HiddenMarkov/nmath, not R), "
                       f"{t:.2f} s wall"), ncell * positions, t


def run_reference(args, wl, kind):
    """--impl reference: the CPU arm alone (rank 0), same line shape."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps_done, times = 0, []
    info = None
    for i in range(args.warmup + args.steps):
        info, nsteps, t = cpu_rate(kind, wl, target_s=4.0)
        if i >= args.warmup:
            steps_done += nsteps
            times.append(t)
    tot = sum(times)
    val = steps_done / tot
    info["value"] = val
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(1, len(times)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": wl["name"], "bounded_sample": info["sample"]},
            "cpu_baseline": info,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of
    the e2e leg are first-touched on the GPU's NUMA node (matters from 2 ranks up)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    # default K: 30 launches (~60 ms) — enough repetitions for a stable mean and still inside the
    # GPU's boost window; beyond ~25 launches this pool's B200s hit sw_power_cap and the same kernel
    # runs 8-10 % slower (kernel_ms_quartiles in the output shows the drift for any K)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="expr", choices=list(WORKLOADS))
    ap.add_argument("--cells", type=int, default=0, help="override cells per GPU (tests)")
    ap.add_argument("--positions", type=int, default=0, help="override genes/SNPs (tests)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-modes", action="store_true",
                    help="skip the viterbi_only / fb_only / sustained legs reported under \"modes\"")
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded recursion-level leg")
    ap.add_argument("--no-flow", action="store_true",
                    help="skip the device-resident end-to-end leg (upload once, calcGexpCnvBoundaries(init=TRUE))")
    ap.add_argument("--no-allele", action="store_true",
                    help="skip the extra device-resident allele measurement reported under \"allele\"")
    args = ap.parse_args()
    kind = args.workload
    c5 = kind == "c5"
    wl = dict(WORKLOADS[kind])
    if c5:
        kind = "expr"   # same inputs, kernels and legs as the expression workload; the step adds the level
    if args.cells:
        wl["cells"] = args.cells
    if args.positions:
        wl["positions"] = args.positions
    if args.cells or args.positions:
        wl["name"] += f" [overridden: {wl['cells']} x {wl['positions']}]"
    if c5 and args.steps > 10:
        args.steps = 10   # a step holds a whole recursion level (0.06 - 0.3 s)
    if args.impl == "reference":
        return run_reference(args, wl, kind)

    import torch
    import torch.distributed as dist
    from honeybadger_b200 import _lib, hmm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    _lib.check(_lib.load().hb_set_device(local))
    # the ranks of one box share its cores: divide them for the host pipeline's copy-out threads
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 4)
    _lib.check(_lib.load().hb_set_host_threads(max(2, min(16, (ncpu - 2) // max(1, world)))))

    cells, positions = wl["cells"], wl["positions"]
    d = device_inputs(kind, cells, positions, seed=100 + rank, dev=dev)
    seg = d["seg"]
    Pi3, Pi2 = hmm.sym_Pi(3, 1e-6), hmm.sym_Pi(2, 1e-6)
    mean, delta3, prob, delta2 = [-DEV, 0.0, DEV], [0.0, 1.0, 0.0], [0.1, 0.45], [0.0, 1.0]
    # Resident layout of the timed leg: warp-tiled [tile of 32 cells][position][lane]
    # (hmm.tile_rows).  The gene-major copy is kept for the e2e leg's host buffers.
    if kind == "expr":
        sd = hmm.chain_sd_dev(d["x"])
        xt = hmm.tile_rows(d["x"])
        out = hmm.hmm_norm_tiled_dev(xt, cells, sd, mean, Pi3, delta3, seg_off=seg)

        chain_ev = []

        def step():
            if c5:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            hmm.hmm_norm_tiled_dev(xt, cells, sd, mean, Pi3, delta3, seg_off=seg, out=out)
            if c5:
                e1.record()
                chain_ev.append((e0, e1))
                hd.sharded_gexp_level(d["x"], c5_labels, cells * world, DEV)
        if c5:
            from honeybadger_b200 import dist as hd
            c5_labels = level_labels(cells, rank, world, dev, seed=100 + rank)
        # sd_prep_kernel + norm3_fast_kernel (+ for c5 the level's ~40 small launches: transposes, cell lists, the
        # two accumulation passes, sd, emissions, the long-chain phases, three retests)
        launches_per_step = 42 if c5 else 2
        in_bytes = d["x"].numel() * 8
    else:
        xn = hmm.pack_counts(hmm.tile_rows(d["r"]), hmm.tile_rows(d["n"]))
        out = hmm.hmm_binom_tiled_dev(xn, None, cells, prob, Pi2, delta2, seg_off=seg)

        def step():
            hmm.hmm_binom_tiled_dev(xn, None, cells, prob, Pi2, delta2, seg_off=seg, out=out)
        launches_per_step = 2  # binom2_fast_kernel + the gated general kernel (exits at once)
        in_bytes = d["r"].numel() * 4
    hmm.status_dev()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.time()
    barrier()
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    barrier()
    t_wall1 = time.time()
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop(t_wall0, t_wall1)
    hmm.status_dev()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    steps_per_pass = cells * positions
    value = world * steps_per_pass * args.steps / (total_ms * 1e-3)
    peak, peak_src = measured_peak()

    # ---------------- parity of the timed buffers against the oracle (rank 0; part of the CPU leg)
    parity = None
    if rank == 0 and not args.no_cpu:
        try:
            parity = parity_block(kind, d, out, cells, seg, sd if kind == "expr" else None, hmm)
        except Exception as exc:
            parity = {"error": repr(exc)[:200]}

    # ---------------- the same kernel in its other modes, each against its own roofline, and a sustained run
    modes = None
    if not args.no_modes and not c5:
        try:
            modes = {}
            nl = max(5, args.steps // 3)
            for name, kw, bps in (("viterbi_only", dict(viterbi=True, gamma=False), 8 + 1 if kind == "expr" else 4 + 1),
                                  ("fb_only", dict(viterbi=False, gamma=True),
                                   8 + 24 if kind == "expr" else 4 + 16)):
                if kind == "expr":
                    o_m = hmm.hmm_norm_tiled_dev(xt, cells, sd, mean, Pi3, delta3, seg_off=seg, **kw)
                    fn = lambda: hmm.hmm_norm_tiled_dev(xt, cells, sd, mean, Pi3, delta3, seg_off=seg, out=o_m, **kw)  # noqa: E731
                else:
                    o_m = hmm.hmm_binom_tiled_dev(xn, None, cells, prob, Pi2, delta2, seg_off=seg, **kw)
                    fn = lambda: hmm.hmm_binom_tiled_dev(xn, None, cells, prob, Pi2, delta2, seg_off=seg, out=o_m, **kw)  # noqa: E731
                ms = time_launches(fn, nl, torch)
                ach = steps_per_pass * bps / (ms * 1e-3) / 1e9
                modes[name] = {"ms": ms, "value": steps_per_pass / (ms * 1e-3), "algorithmic_bytes_per_step": bps,
                               "roofline_frac": ach / peak, "launches": nl}
                del o_m
            if kind == "expr":
                # The chain-resident design (csrc/hb_norm3_resident.cuh): posteriors with x read once and no
                # checkpoints (32 B / step of traffic), on a chain-major copy of the same matrix; and the
                # two-kernel form of the whole step — Viterbi-only (thread per chain, tiled copy) on one stream
                # BESIDE the resident posteriors on another: 43 B / step of traffic instead of 49.
                x_cm = d["x"].t().contiguous()
                g_cm = torch.empty((3, cells, positions), dtype=torch.float64, device=dev)
                fr = lambda: hmm.forwardback_norm_cm_dev(x_cm, sd, mean, Pi3, delta3, seg_off=seg, out=g_cm)  # noqa: E731
                fr()
                gate = hmm.fbr_gate()
                ms = time_launches(fr, nl, torch)
                chk = rng_cols = np.random.default_rng(3).choice(cells, 64, replace=False)
                a_ = g_cm[:, torch.as_tensor(chk, device=dev), :].permute(0, 2, 1).cpu().numpy()
                b_ = tiled_cols(out["gamma"], chk)
                modes["fb_only_resident"] = {
                    "ms": ms, "value": steps_per_pass / (ms * 1e-3), "algorithmic_bytes_per_step": 32,
                    "roofline_frac": steps_per_pass * 32 / (ms * 1e-3) / 1e9 / peak, "launches": nl, "gate": gate,
                    "max_abs_diff_vs_thread_per_chain": float(np.abs(a_ - b_).max()),
                    "layout": "chain-major x[cell][gene] (R's column-major), one block of 1-4 warps per chain"}
                o_v = hmm.hmm_norm_tiled_dev(xt, cells, sd, mean, Pi3, delta3, seg_off=seg, viterbi=True, gamma=False)
                s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()

                def both():
                    cur = torch.cuda.current_stream()
                    s_a.wait_stream(cur)
                    s_b.wait_stream(cur)
                    with torch.cuda.stream(s_a):
                        hmm.hmm_norm_tiled_dev(xt, cells, sd, mean, Pi3, delta3, seg_off=seg, viterbi=True,
                                               gamma=False, out=o_v)
                    with torch.cuda.stream(s_b):
                        hmm.forwardback_norm_cm_dev(x_cm, sd, mean, Pi3, delta3, seg_off=seg, out=g_cm)
                    cur.wait_stream(s_a)
                    cur.wait_stream(s_b)
                ms = time_launches(both, nl, torch)
                same_y = bool(all(torch.equal(p, q) for p, q in zip(
                    [o_v["y"][:-1], o_v["y"][-1:, :, :cells % 32 or 32]],
                    [out["y"][:-1], out["y"][-1:, :, :cells % 32 or 32]])))
                modes["viterbi_beside_resident_fb"] = {
                    "ms": ms, "value": steps_per_pass / (ms * 1e-3), "algorithmic_bytes_per_step": 33,
                    "roofline_frac": steps_per_pass * 33 / (ms * 1e-3) / 1e9 / peak, "launches": nl,
                    "same_states_as_fused_kernel": same_y,
                    "note": "two streams: norm3_fast_kernel<VIT> (tiled copy) || norm3_resident_kernel (chain-major "
                            "copy); needs the matrix resident in both layouts"}
                del x_cm, g_cm, o_v
            # sustained: >= 2 s of back-to-back launches of the headline kernel, clocks sampled throughout
            ns = max(args.steps, int(2000.0 / (total_ms / args.steps)) + 1)
            smp = ClockSampler(local)
            smp.start()
            time.sleep(0.2)
            tw0 = time.time()
            ms = time_launches(step, ns, torch)
            tw1 = time.time()
            ach = steps_per_pass * wl["bytes_per_step"] / (ms * 1e-3) / 1e9
            modes["sustained"] = {"ms": ms, "launches": ns, "value": steps_per_pass / (ms * 1e-3),
                                  "roofline_frac": ach / peak, "clocks": smp.stop(tw0, tw1)}
            hmm.status_dev()
        except Exception as exc:
            modes = {"error": repr(exc)[:200]}

    # ---------------- one recursion level of the split-and-retest driver on the SHARDED resident matrix: the
    # exchanges of SURVEY.md 8e (group sizes, 80-bit relay of the pooled means, retest all-reduce, posterior
    # all-gather) run here, on NCCL when N > 1, and the result is compared with one rank holding every cell
    sharded = None
    if kind == "expr" and not args.no_sharded and not c5:
        try:
            sharded = run_sharded_level(d["x"], cells, rank, world, dev, dist if world > 1 else None)
        except Exception as exc:
            sharded = {"error": repr(exc)[:300]}

    # ---------------- device-resident end to end (upload once -> boundaries out), rank 0 of a 1-GPU run
    flow = None
    if kind == "expr" and world == 1 and not c5 and not args.no_flow and not (args.cells or args.positions):
        try:
            flow = run_resident_flow(cells, positions)
        except Exception as exc:
            flow = {"error": repr(exc)[:300]}

    # ---------------- end to end through the host-buffer C ABI (pinned host memory)
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(kind, wl, d, args, world, dev, hmm, dist if world > 1 else None)
        except Exception as exc:  # report, never fake
            e2e = {"value": None, "unit": UNIT, "error": repr(exc)[:200]}

    # ---------------- the other half of C4 (10k cells x 200k SNPs, 2-state binomial), device-resident
    # only, a few steps: reported under "allele" beside the headline expression workload
    allele = None
    if kind == "expr" and not c5 and not args.no_allele and not (args.cells or args.positions):
        try:
            del out, xt, step
            d.pop("x", None)
            torch.cuda.empty_cache()
            allele = run_allele_extra(cells, rank, dev, hmm, world, dist if world > 1 else None)
        except Exception as exc:
            allele = {"value": None, "error": repr(exc)[:200]}

    if c5:  # the roofline is the chain kernel's: its own events inside the timed steps
        per_launch_ms = [a.elapsed_time(b) for a, b in chain_ev[-args.steps:]]
    mean_ms = float(np.mean(per_launch_ms))
    achieved = steps_per_pass * wl["bytes_per_step"] / (mean_ms * 1e-3) / 1e9
    q = np.percentile(per_launch_ms, [0, 25, 50, 75, 100])
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "kernel_ms_quartiles": [round(float(v), 4) for v in q],
            "traffic": ncu_traffic(kind), "peak_source": peak_src,
            "algorithmic_bytes_per_step": wl["bytes_per_step"], "kernel_ms": mean_ms,
            "kernel": "norm3_fast_kernel<VIT,FB>" if kind == "expr" else "binom2_fast_kernel<VIT,FB>"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:  # the CPU arm gets every host core again
            os.sched_setaffinity(0, range(os.cpu_count() or 1))
        except Exception:
            pass
        cpu, _, _ = cpu_rate(kind, wl)
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": wl["name"], "cells_per_gpu": cells, "positions": positions,
                           "segments": wl["nseg"], "states": wl["S"],
                           "l2": f"inputs {in_bytes / 1e6:.0f} MB + outputs per pass exceed the 126 MB L2",
                           "layout": "warp-tiled [tile of 32 cells][position][lane] (hb_tile_dev), fp64 / int32"},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
                "gpu_launches": launches_per_step * args.steps, "parity": parity, "modes": modes,
                "sharded_level": sharded, "resident_flow": flow,
                "kernel_source_hash": kernel_source_hash(), "kernel_sass_hash": kernel_sass_hash(kind)}
        if allele is not None:
            line["allele"] = allele
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def level_labels(cells, rank, world, dev, seed):
    """Group labels of a 5-height cut for this rank's cells (GLOBAL ids), the way the benchmark sizes get them
    (supplied by the generator — nested cuts of the clone tree, the clones of device_inputs)."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    clone = torch.multinomial(torch.tensor([0.5, 0.3, 0.2], device=dev), cells, True, generator=g).cpu().numpy()
    idx = np.arange(cells) + rank * cells
    return [np.zeros(cells, np.int64), (clone == 0).astype(np.int64), clone.astype(np.int64),
            clone * 2 + (idx % 2) * (clone == 0), clone * 3 + (idx % 3) * (clone == 0)]


def run_sharded_level(x, cells, rank, world, dev, dist, check_cells=1024, check_genes=4000):
    """(1) timed: one level (labels given) on every rank's full shard — pooled means by the exact relay, K
    pooled chains, vote / runs, retest with its all-reduce, posterior all-gather.  (2) checked: the same level
    on a reduced slice (first `check_cells` cells of every rank x first `check_genes` genes) against rank 0
    running it alone on the gathered slice — pooled means / states / runs / cell sets must be identical."""
    import torch
    from honeybadger_b200 import dist as hd
    n_total = cells * world
    labels = level_labels(cells, rank, world, dev, seed=100 + rank)  # the clones device_inputs planted
    hd.sharded_gexp_level(x, labels, n_total, DEV)                   # warm-up (workspace growth, first launches)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    lvl = hd.sharded_gexp_level(x, labels, n_total, DEV)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = {"ms": float(t.item()) * 1e3, "cells_total": n_total, "groups": lvl["K"],
           "runs": len(lvl.get("runs", [])), "g1_cells": int(lvl["g1"].size) if "g1" in lvl else 0,
           "collectives": lvl["collectives"], "backend": "nccl" if dist is not None else "none (1 rank)"}
    # ---- equality with a single rank on a reduced slice
    cc, gg = min(check_cells, cells), min(check_genes, x.shape[0])
    xs = x[:gg, :cc].contiguous()
    ls = [l[:cc] for l in labels]
    got = hd.sharded_gexp_level(xs, ls, cc * world, DEV)
    if dist is not None:
        parts = [torch.empty_like(xs) for _ in range(world)]
        dist.all_gather(parts, xs)
        lab_t = torch.as_tensor(np.stack(ls), device=dev)
        labs = [torch.empty_like(lab_t) for _ in range(world)]
        dist.all_gather(labs, lab_t)
        same = True
        if rank == 0:
            import torch.distributed as tdist
            full = torch.cat(parts, dim=1).contiguous()
            lab_full = torch.cat(labs, dim=1).cpu().numpy()
            # rank 0 alone: temporarily outside the group's collectives (world-1 code path of the same function)
            ref = _single_rank_level(hd, full, [lab_full[h] for h in range(lab_full.shape[0])], cc * world)
            same = bool(np.array_equal(got["pooled"], ref["pooled"]) and np.array_equal(got["y"], ref["y"])
                        and len(got["runs"]) == len(ref["runs"])
                        and all(np.array_equal(p, q) for p, q in zip(got["runs"], ref["runs"]))
                        and ("g1" not in ref or (np.array_equal(got["g1"], ref["g1"])
                                                 and np.array_equal(got["g2"], ref["g2"])
                                                 and float(np.abs(got["prob"] - ref["prob"]).max()) < 1e-12)))
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.broadcast(flag, src=0)
        res["equals_single_gpu"] = bool(flag.item())
        res["checked_on"] = f"{cc * world} cells x {gg} genes (first {cc} cells of every rank)"
        res["grouping"] = run_sharded_grouping(x, cells, rank, world, dev, dist, hd)
    return res


def run_sharded_grouping(x, cells, rank, world, dev, dist, hd, total=20000):
    """The grouping step of a level on sharded cells (`dist.sharded_ward_groups`): runmean local, smoothed cells
    all-gathered, `dist` split by blocks of the symmetric matrix over the ranks, Ward + cutree on rank 0, labels
    broadcast.  STRONG scaling by construction (a fixed population of `total` cells, the first total / world of
    every rank): the pair count is quadratic in the cells and the linkage is sequential, so this is the part of a
    level that does not scale with the box — timed with its sequential remainder named."""
    import torch
    from honeybadger_b200 import _lib, hmm
    per = min(cells, total // world)
    xs = hmm.to_rows(x[:, :per].contiguous())
    sizes = [per] * world
    hd.sharded_ward_groups(xs, sizes, [1, 2, 3, 4, 5])   # warm-up
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    labs = hd.sharded_ward_groups(xs, sizes, [1, 2, 3, 4, 5])
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # this rank's share of the pair work alone (its blocks of the matrix), for the split of the total
    lib = _lib.load()
    G = xs.shape[0]
    S = hmm.alloc_rows(G, per, torch.float64, dev)
    _lib.check(lib.hb_runmean_dev(xs.data_ptr(), xs.stride(0), G, per, 101, S.data_ptr(), S.stride(0), None))
    plans = [(q, hd._block_rows(rank, q, world, sizes)) for q in range(world)]
    nblk = sum(1 for _, p in plans if p is not None)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for q, plan in plans:
        if plan is None:
            continue
        if q == rank:
            blk = torch.empty((per, per), dtype=torch.float64, device=dev)
            _lib.check(lib.hb_dist_sq_dev(S.data_ptr(), S.stride(0), G, per, blk.data_ptr(), None))
        else:
            kind, a, b = plan
            na, nb = (b - a, per) if kind == "rows" else (per, b - a)
            blk = torch.empty((na, nb), dtype=torch.float64, device=dev)
            pa = S.data_ptr() + (a * 8 if kind == "rows" else 0)
            pb = S.data_ptr() + (a * 8 if kind == "cols" else 0)
            _lib.check(lib.hb_dist_sq_rect_dev(pa, S.stride(0), na, pb, S.stride(0), nb, G, blk.data_ptr(), blk.stride(0),
                                               None))
    torch.cuda.synchronize()
    td = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device=dev)
    dist.all_reduce(td, op=dist.ReduceOp.MAX)
    return {"ms": float(t.item()) * 1e3, "cells_total": per * world, "genes": int(G), "scaling": "strong",
            "ms_dist_blocks_max_over_ranks": float(td.item()) * 1e3, "blocks_this_rank": nblk,
            "groups_at_k5": int(len(np.unique(labs[-1]))),
            "exchanges": ["all_gather(smoothed cells)", "send/recv(row panels of the distance matrix -> rank 0)",
                          "broadcast(labels)"],
            "sequential_on_rank0": "Ward.D2 linkage + cutree"}


def _single_rank_level(hd, x_full, labels_full, n_cells):
    """sharded_gexp_level's world = 1 path while a process group exists: patch the world size it sees."""
    import torch.distributed as tdist
    real_ws, real_init = tdist.get_world_size, tdist.is_initialized
    tdist.get_world_size, tdist.is_initialized = (lambda *a, **k: 1), (lambda: False)
    try:
        return hd.sharded_gexp_level(x_full, labels_full, n_cells, DEV)
    finally:
        tdist.get_world_size, tdist.is_initialized = real_ws, real_init


def run_resident_flow(cells, positions):
    """Device-resident end to end: BASELINE config 4's expression half the way the product is meant to be
    used — the host matrix (R layout) is uploaded ONCE, then HoneyBADGER$calcGexpCnvBoundaries(init=TRUE)
    (R/HoneyBADGER.R:1091-1316: grouping, pooled rounds, vote / runs, retest, split, recursion) runs on the
    resident matrix through the RefClass mirror; only index sets, memberships, K state paths per level and the
    per-region posteriors cross PCIe afterwards.  Wall time includes the upload."""
    import pandas as pd
    import torch
    from honeybadger_b200 import honeybadger as hb, synth
    rng = np.random.default_rng(4)
    G, N = positions, cells
    seg = synth.segments(G)
    clone = rng.choice(3, size=N, p=[0.5, 0.3, 0.2])
    x = rng.standard_normal((G, N))
    x *= rng.uniform(*DEV_SIGMA, size=N)
    x[seg[6]:seg[7], clone == 0] += DEV
    x[seg[9]:seg[10], clone == 0] -= DEV
    x[seg[12]:seg[13], clone == 1] -= DEV
    names = [f"g{i}" for i in range(G)]
    chrom = np.empty(G, dtype=object)
    for c in range(22):
        chrom[seg[c]:seg[c + 1]] = f"chr{c + 1}"
    pos = np.concatenate([np.arange(seg[c + 1] - seg[c]) * 1000.0 for c in range(22)])
    genes = hb.GRanges(chrom, pos, pos + 500, names)
    frame = pd.DataFrame(np.asfortranarray(x), index=names, columns=[f"c{i}" for i in range(N)])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    h = hb.HoneyBADGER("c4")
    h.genes = genes
    h.gexp_norm = frame
    h.setGexpDev(dev=DEV)
    h._resident("gexp")                      # the one upload
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    h.calcGexpCnvBoundaries(init=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    regs = []
    for reg in h.bound_genes_final:
        nm = reg[0]
        cells_in = (h.pred_genes_r.loc[nm[0]] > 0).values
        regs.append({"chr": int(np.searchsorted(seg, int(nm[0][1:]), side="right")), "genes": len(nm),
                     "cells": int(cells_in.sum()), "clone_mix": np.bincount(clone[cells_in], minlength=3).tolist()})
    return {"api": "HoneyBADGER.calcGexpCnvBoundaries(init=True) on the resident matrix (RefClass mirror)",
            "cells": N, "genes": G, "upload_s": t1 - t0, "flow_s": t2 - t1, "total_s": t2 - t0,
            "h2d_bytes": int(G) * int(N) * 8, "regions": regs,
            "steps_per_s_pooled_equivalent": float(G) * N / (t2 - t0)}


def tiled_cols(t, cols):
    """columns of a warp-tiled [ntile, T, 32] / [planes, ntile, T, 32] tensor as numpy [T, n] / [planes, T, n]"""
    import torch
    if t.dim() == 4:
        return np.stack([tiled_cols(p, cols) for p in t])
    c = torch.as_tensor(np.asarray(cols), device=t.device)
    return t[c // 32, :, c % 32].t().cpu().numpy()


def parity_block(kind, d, out, cells, seg, sd, hmm, ncols=96):
    """Chains sampled from the TIMED buffers (plus, for the Gaussian kernel, every chain it re-ran exactly)
    against the CPU oracle: states must be identical, posteriors within 1e-9.  Part of the cpu_baseline leg —
    the oracle is the checker, never the thing measured."""
    import torch
    from oracle import oracle as O
    O.build()
    rng = np.random.default_rng(11)
    cols = rng.choice(cells, size=min(ncols, cells), replace=False)
    info = {}
    if kind == "expr":
        fl = hmm.norm3_flagged_chains(cells)
        info["flagged_chains"] = int(hmm.norm3_rerun_chains())
        cols = np.unique(np.concatenate([cols, fl[:64, 1]])).astype(np.int64)
        ct = torch.as_tensor(cols, device=d["x"].device)
        yo, go, _, rc = O.fbv_norm_batch(d["x"][:, ct].cpu().numpy(), seg, [-DEV, 0.0, DEV], sd[ct].cpu().numpy(),
                                         O.sym_Pi(3, 1e-6), [0.0, 1.0, 0.0])
        info["kernel"] = "norm3_fast_kernel" if hmm.last_kernels()["norm3_fast"] == 1 else "norm3_kernel"
    else:
        cols = np.unique(cols).astype(np.int64)
        ct = torch.as_tensor(cols, device=d["r"].device)
        yo, go, _, rc = O.fbv_binom_batch(d["r"][:, ct].cpu().numpy(), d["n"][:, ct].cpu().numpy(), seg,
                                          [0.1, 0.45], O.sym_Pi(2, 1e-6), [0.0, 1.0])
        info["kernel"] = "binom2_fast_kernel" if hmm.last_kernels()["binom2_lean"] == 1 else "binom2_kernel"
    y = tiled_cols(out["y"], cols)
    g = tiled_cols(out["gamma"], cols)
    info.update({"chains_checked": int(len(cols) * (len(seg) - 1)), "cells_checked": int(len(cols)),
                 "state_mismatches": int((y != yo).sum()), "max_gamma_err": float(np.abs(g - go).max()),
                 "oracle_rc": int(rc), "against": "oracle/hb_oracle.c on the same resident inputs"})
    return info


def time_launches(fn, n, torch):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run_allele_extra(cells, rank, dev, hmm, world, dist, steps=5):
    """Device-resident throughput of the allele half of C4 on the same GPU(s): max over ranks."""
    import torch
    wl = WORKLOADS["allele"]
    positions = wl["positions"]
    d = device_inputs("allele", cells, positions, seed=300 + rank, dev=dev)
    Pi2 = hmm.sym_Pi(2, 1e-6)
    pooled = None
    if world == 1:
        try:
            pooled = run_pooled_round(d, cells, hmm)
        except Exception as exc:
            pooled = {"error": repr(exc)[:200]}
    xn = hmm.pack_counts(hmm.tile_rows(d["r"]), hmm.tile_rows(d["n"]))
    del d["r"], d["n"]
    torch.cuda.empty_cache()
    out = hmm.hmm_binom_tiled_dev(xn, None, cells, [0.1, 0.45], Pi2, [0.0, 1.0], seg_off=d["seg"])
    for _ in range(3):
        hmm.hmm_binom_tiled_dev(xn, None, cells, [0.1, 0.45], Pi2, [0.0, 1.0], seg_off=d["seg"], out=out)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        hmm.hmm_binom_tiled_dev(xn, None, cells, [0.1, 0.45], Pi2, [0.0, 1.0], seg_off=d["seg"], out=out)
    e1.record()
    torch.cuda.synchronize()
    hmm.status_dev()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    val = world * cells * positions / (ms * 1e-3)
    peak, _ = measured_peak()
    ach = cells * positions * wl["bytes_per_step"] / (ms * 1e-3) / 1e9
    rerun = hmm.binom2_general_rerun()
    bb = None
    if world == 1:
        try:
            bb = run_betabinom(xn, cells, positions, d["seg"], hmm, wl, peak, steps)
        except Exception as exc:
            bb = {"error": repr(exc)[:200]}
    return {"workload": wl["name"], "value": val, "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": ncu_traffic("allele"), "algorithmic_bytes_per_step": wl["bytes_per_step"],
                         "kernel": "binom2_fast_kernel<VIT,FB>"},
            "general_kernel_rerun": rerun, "pooled_round": pooled, "betabinom": bb}


def run_betabinom(xn, cells, positions, seg, hmm, wl, peak, steps, rho=0.02):
    """BASELINE config 3 with north_star's beta-binomial emissions (rho > 0; an extension — the reference is
    binomial): the same resident packed counts and kernel, emission table built for the over-dispersed model
    up to the workload's maximum count; sampled chains against the oracle at the same rho."""
    import torch
    from honeybadger_b200 import _lib
    from oracle import oracle as O
    lib = _lib.load()
    Pi2, prob, delta = hmm.sym_Pi(2, 1e-6), [0.1, 0.45], [0.0, 1.0]
    _lib.check(lib.hb_set_binom_lut_max(1245))
    try:
        out = hmm.hmm_binom_tiled_dev(xn, None, cells, prob, Pi2, delta, seg_off=seg, rho=rho)
        hmm.status_dev()
        fn = lambda: hmm.hmm_binom_tiled_dev(xn, None, cells, prob, Pi2, delta, seg_off=seg, rho=rho, out=out)  # noqa: E731
        ms = time_launches(fn, steps, torch)
        hmm.status_dev()
        cols = np.random.default_rng(13).choice(cells, size=32, replace=False)
        w = tiled_cols(xn, cols)
        r_np, n_np = (w & 0xFFFF).astype(np.int32), ((w >> 16) & 0xFFFF).astype(np.int32)
        yo, go, _, rc = O.fbv_binom_batch(r_np, n_np, seg, prob, O.sym_Pi(2, 1e-6), delta, rho=rho)
        y, g = tiled_cols(out["y"], cols), tiled_cols(out["gamma"], cols)
        par = {"chains_checked": int(len(cols) * (len(seg) - 1)), "state_mismatches": int((y != yo).sum()),
               "max_gamma_err": float(np.abs(g - go).max()), "oracle_rc": int(rc)}
    finally:
        _lib.check(lib.hb_set_binom_lut_max(511))
    ach = cells * positions * wl["bytes_per_step"] / (ms * 1e-3) / 1e9
    return {"rho": rho, "ms_per_step": ms, "value": cells * positions / (ms * 1e-3), "roofline_frac": ach / peak,
            "general_kernel_rerun": hmm.binom2_general_rerun(), "parity": par,
            "note": "beta-binomial emissions (mean p, intra-class correlation rho) — extension, reference is binomial"}


def run_pooled_round(d, cells, hmm):
    """The reference's own call shape on the same resident matrices: one `lapply(heights, ...)` round of
    calcAlleleCnvBoundaries (R/HoneyBADGER.R:1395-1417) — pooled counts of K = 6 cell sets, K binomial chains
    of all 200k SNPs — through hb_allele_pooled_viterbi_dev, host results included.  The chains run parallel
    along the gene axis (csrc/hb_longchain.cuh); `ms_thread_per_chain` is the same round with one thread per
    chain (hb_set_longchain_mode(1)), same states."""
    import time
    import torch
    from honeybadger_b200 import honeybadger as hb
    idx = np.arange(cells)
    member = np.stack([np.ones(cells, bool)] + [idx % 2 == k for k in range(2)] + [idx % 3 == k for k in range(3)])
    res = {}
    ys = {}
    for mode, key in ((0, "ms"), (1, "ms_thread_per_chain")):
        hmm.set_longchain_mode(mode)
        try:
            ts = []
            for _ in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                y, _, _ = hb.allele_round_dev(d["r"], d["n"], member, 0.1, 0.45, 1e-6)
                ts.append((time.perf_counter() - t0) * 1e3)
            res[key] = float(np.median(ts[1:]))
            ys[mode] = y
            if mode == 0:
                res["longchain"] = hmm.longchain_stats()
        finally:
            hmm.set_longchain_mode(0)
    res.update({"chains": int(member.shape[0]), "positions": int(d["r"].shape[0]), "cells": int(cells),
                "same_states": bool(np.array_equal(ys[0], ys[1])),
                "api": "hb_allele_pooled_viterbi_dev (pooled counts + emissions + K long chains + download)"})
    return res


def run_e2e(kind, wl, d, args, world, dev, hmm, dist):
    """Same workload through hb_hmm_*_host: host (pinned) column-major buffers in R's layout, every
    byte crossing PCIe inside the timed region.  Cells are streamed in column batches through a
    bounded pinned window so that host memory stays small at any N."""
    import torch
    cells, positions = wl["cells"], wl["positions"]
    batch = min(cells, 10000 if kind == "expr" else 1000)
    nb = (cells + batch - 1) // batch
    S = wl["S"]
    seg = d["seg"]
    Pi3, Pi2 = hmm.sym_Pi(3, 1e-6), hmm.sym_Pi(2, 1e-6)
    # R layout: column-major [positions x cells] == C-order [cells, positions]
    if kind == "expr":
        xin = torch.empty((cells, positions), dtype=torch.float64).pin_memory()
        xin.copy_(d["x"].t())
        sd_h = hmm.chain_sd_dev(d["x"]).cpu().numpy()
        h2d = positions * cells * 8
    else:
        rin = torch.empty((cells, positions), dtype=torch.int32).pin_memory()
        nin = torch.empty((cells, positions), dtype=torch.int32).pin_memory()
        rin.copy_(d["r"].t())
        nin.copy_(d["n"].t())
        h2d = positions * cells * 8
    yout = torch.empty((batch, positions), dtype=torch.int32).pin_memory()
    gout = torch.empty((S, batch, positions), dtype=torch.float64).pin_memory()
    # states cross PCIe as bytes (widened to R integers on the host); the neutral-state posterior plane is
    # rebuilt on the host from the S - 1 planes that do cross (HB_E2E_FULL_PLANES=1 downloads all S)
    full = os.environ.get("HB_E2E_FULL_PLANES", "0") not in ("", "0")
    d2h = positions * cells * (1 + 8 * (S if full else S - 1))
    from honeybadger_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    segp = np.ascontiguousarray(seg, dtype=np.int64)
    mean = np.array([-DEV, 0.0, DEV])
    delta3, delta2, prob = np.array([0.0, 1.0, 0.0]), np.array([0.0, 1.0]), np.array([0.1, 0.45])
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731

    def one_pass():
        for b in range(nb):
            b0, bn = b * batch, min(batch, cells - b * batch)
            # gamma for a short last batch: planes are [T x bn] each, contiguous at the front
            if kind == "expr":
                rc = lib.hb_hmm_norm_host(xin[b0:b0 + bn].data_ptr(), positions, bn, vp(segp), segp.size - 1,
                                          vp(mean), vp(np.ascontiguousarray(sd_h[b0:b0 + bn])), 0, vp(Pi3),
                                          vp(delta3), 3, yout.data_ptr(), gout.data_ptr(), None)
            else:
                rc = lib.hb_hmm_binom_host(rin[b0:b0 + bn].data_ptr(), nin[b0:b0 + bn].data_ptr(), positions,
                                           bn, vp(segp), segp.size - 1, vp(prob), 0.0, vp(Pi2), vp(delta2),
                                           3, yout.data_ptr(), gout.data_ptr(), None)
            _lib.check(rc)
            lib.hb_last_host_io(C.byref(io[0]), C.byref(io[1]), C.byref(io[2]))
            for q in range(3):
                tot[q] += io[q].value

    io = [C.c_int64(0) for _ in range(3)]
    tot = [0, 0, 0]
    one_pass()  # warm-up (workspace allocation, LUT)
    tot = [0, 0, 0]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        one_pass()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    val = world * cells * positions * args.e2e_steps / float(t.item())
    # bytes counted by the library itself (hb_last_host_io)
    h2d, d2h = tot[0] // args.e2e_steps, tot[1] // args.e2e_steps
    return {"value": val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "posterior_bytes_rebuilt_on_host_per_step": tot[2] // args.e2e_steps,
            "api": "hb_hmm_norm_host" if kind == "expr" else "hb_hmm_binom_host",
            "host_buffers": "pinned, R column-major layout; one call per window of %d cells, each call "
                            "pipelined internally (upload | layout+chains | download on three streams)" % batch,
            "steps": args.e2e_steps, "cpus_bound_to_gpu_numa_node": len(os.sched_getaffinity(0))}


if __name__ == "__main__":
    main()

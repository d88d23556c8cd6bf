#!/usr/bin/env python3
"""bench.py — BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout.

metric   : simulated packet-events/sec (packet-event = one packet sent by a Rat or one ack
           consumed), whole job over all ranks
workload : BASELINE configs[1] — the default Remy ConfigRange Evaluator pass (link 1..2 ppt
           step 0.25 x rtt 100..200 step 25 x nsrc 1..32 = 800 NetConfigs, on = off = 5000 ms,
           infinite buffer, no loss, fixed WhiskerTree), batched over R evaluator seeds per GPU
           so that one step fills the machine.  One step = one scoring pass over that batch.
value    : device time of the kernel launches (one simulation launch + one overflow scan per step) (CUDA events on the launching stream, inside
           the library), inputs already resident in HBM
e2e      : the same pass through the one-shot C-ABI call with HOST buffers
           (remy_b200_score_batch: H2D copies, scratch allocation, kernels, D2H results)
--impl reference : the reference's own CPU implementation (oracle/_ref, built from
           /root/reference's sources) on all host threads, on a bounded sample of the workload.

Launch: python bench.py --gpus N --steps K --warmup W         (N == 1)
        python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)
"""
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

from remy_b200 import abi, workloads  # noqa: E402

METRIC = "sim_packet_events_per_sec"
UNIT = "packet-events/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build_improve_step(args, dev):
    """BASELINE configs[2]: one ActionImprover::evaluate_replacements fan-out (src/breeder.cc:53-77) as ONE
    batch: score the tree once with use counts, take the most used whisker whose action the optimizer may
    vary (WhiskerTree::most_used, src/whiskertree.cc:84-109), generate its candidates with the C++ mirror of
    Whisker::next_generation, and score candidates x the default ConfigRange at carefulness 0.1."""
    from remy_b200 import hostlib
    import __graft_entry__ as g
    g.build_host()
    tree = workloads.load_golden_tree(args.tree)
    cfgs = workloads.default_config_range(ticks=args.ticks)
    seeds = workloads.seeds_for(2026, len(cfgs))
    _, _, counts = dev.score(tree, cfgs, seeds, want_senders=False, want_use_counts=True)
    host = hostlib.HostLib()
    ht = host.tree_from_flat(tree)
    order = np.argsort(-counts[0].astype(np.int64))
    for leaf in order:
        l = tree.leaves[leaf]
        if 0 <= l["window_increment"] <= 256 and 0 <= l["window_multiple"] <= 1 and 0.25 <= l["intersend"] <= 3:
            cands = ht.next_generation(int(leaf))
            return tree, cfgs, seeds, cands, int(leaf)
    raise SystemExit("no whisker with an optimizable action")


def build_workload(args, rank):
    tree = workloads.load_golden_tree(args.tree)
    base = workloads.default_config_range(ticks=args.ticks)
    cfgs = np.tile(base, args.replicas)
    # every (rank, replica) is its own Evaluator seed; seeds within a replica follow seeds_for()
    seeds = np.concatenate([workloads.seeds_for(1000 + rank * args.replicas + r, len(base)) for r in range(args.replicas)])
    return tree, cfgs, seeds


def algorithmic_bytes(cfgs, sims):
    """per event tick 44 n + 32 n_on bytes of sender state, per packet sent a
    16-byte queue record written and read once, per delivered packet 32 B in the delay line
    + 208 B of Memory/Utility/action."""
    n = cfgs["num_senders"].astype(np.float64)
    n_on = n * cfgs["mean_on_duration"] / (cfgs["mean_on_duration"] + cfgs["mean_off_duration"])
    n_cfg = len(cfgs)
    k = np.arange(len(sims)) % n_cfg
    return float((sims["event_ticks"] * (44.0 * n[k] + 32.0 * n_on[k]) + sims["packets_sent"] * 32.0
                  + sims["packets_received"] * 240.0).sum())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(gpu_index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned_like(a):
    """Copy of `a` in pinned host memory (numpy view of a pinned torch tensor)."""
    import torch
    t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
    out = t.numpy().view(a.dtype).reshape(a.shape)
    out[...] = a
    return out, t


def cpu_sample(args, tree, cfgs, seeds):
    """Bounded sample of the same workload for the CPU leg: every `stride`-th simulation of
    the first replica."""
    n_base = len(cfgs) // args.replicas
    idx = np.arange(0, n_base, args.cpu_stride)
    return cfgs[idx], seeds[idx], f"every {args.cpu_stride}th NetConfig of one 800-config replica ({len(idx)} simulations, {args.ticks:g} ticks each)"


def run_cpu(args, tree, cfgs, seeds):
    """Times the reference's own CPU implementation (oracle/_ref) or, if it was not built,
    the oracle port, on all host threads.  Returns (value, kind, cores, seconds, packet_events)."""
    threads = os.cpu_count() or 1
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libremy_ref.so")
    if os.path.exists(ref_path):
        lib, kind = abi.Reference(ref_path), "reference"
    else:
        import __graft_entry__ as g
        g.build_checkers()
        lib, kind = abi.Oracle(), "port"
    t0 = time.perf_counter()
    sims, _, _ = lib.score(tree, cfgs, seeds, want_senders=False, threads=threads)
    dt = time.perf_counter() - t0
    ev = int((sims["packets_sent"] + sims["packets_received"]).sum())
    return ev / dt, kind, threads, dt, ev, int(sims["event_ticks"].sum())


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else any library prints
    (NCCL's version banner, for instance) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ticks", type=float, default=1e5, help="simulation_ticks per NetConfig")
    ap.add_argument("--replicas", type=int, default=192, help="evaluator seeds per GPU (x 800 NetConfigs)")
    ap.add_argument("--tree", default="RemyCC-2014-100x")
    ap.add_argument("--cpu-stride", type=int, default=4)
    ap.add_argument("--workload", default="evaluator-pass", choices=["evaluator-pass", "improve-step"],
                    help="evaluator-pass: BASELINE configs[1] (the bench line of record); improve-step: BASELINE configs[2], "
                         "every Whisker::next_generation candidate of the most used whisker x the 800 NetConfigs in one batch")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}")
    n_gpus = max(world, 1)

    tree, cfgs, seeds = build_workload(args, rank)
    workload = (f"BASELINE configs[1]: default Remy ConfigRange (800 NetConfigs: link 1-2 ppt x rtt 100-200 ms x nsrc 1-32, "
                f"exp on/off 5 s, infinite buffer) x {args.replicas} evaluator seeds per GPU, fixed WhiskerTree {args.tree}, "
                f"{args.ticks:g} ticks")
    config = {"workload": workload, "sims_per_gpu": int(len(cfgs)), "ticks": args.ticks, "tree": args.tree,
              "l2": "per-slot simulation state (>200 MB) exceeds the 126 MB L2; a 256 MiB buffer is also written between steps"}

    # ------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        scfg, sseed, sample = cpu_sample(args, tree, cfgs, seeds)
        for _ in range(min(args.warmup, 1)):
            run_cpu(args, tree, scfg[:32], sseed[:32])
        tot_t, tot_ev, tot_ticks = 0.0, 0, 0
        for _ in range(args.steps):
            v, kind, cores, dt, ev, ticks = run_cpu(args, tree, scfg, sseed)
            tot_t += dt; tot_ev += ev; tot_ticks += ticks
        value = tot_ev / tot_t
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "event_ticks_per_sec": tot_ticks / tot_t,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ------------------------------------------------------------------ GPU arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = abi.Device(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush_buf.fill_(1)
        torch.cuda.synchronize()

    cands = None
    if args.workload == "improve-step":
        tree, cfgs, seeds, cands, leaf = build_improve_step(args, dev)
        config["workload"] = (f"BASELINE configs[2]: one RatBreeder improve fan-out — {len(cands)} Whisker::next_generation candidates of "
                              f"leaf {leaf} of {args.tree} x 800 NetConfigs (default ConfigRange), {args.ticks:g} ticks, one batch per GPU")
        config["sims_per_gpu"] = int(len(cands) * len(cfgs))
        args.replicas = 1

    # value: staged API, inputs resident in HBM
    batch = dev.create_batch(tree, cfgs, seeds, cands=cands, want_senders=False, want_use_counts=False)
    for _ in range(args.warmup):
        batch.run(timed=True)
        flush_l2()
    sampler = ClockSampler(local_rank)
    barrier()
    kernel_ms, launches = 0.0, 0
    for _ in range(args.steps):
        kernel_ms += batch.run(timed=True)  # CUDA events on the library's launching stream
        launches += batch.launches()
        flush_l2()
    barrier()
    clocks = sampler.stop()
    sims, _, _ = batch.fetch()
    batch.destroy()
    assert (sims["error"] == 0).all(), np.unique(sims["error"])
    pk = int((sims["packets_sent"] + sims["packets_received"]).sum())
    ticks = int(sims["event_ticks"].sum())
    alg_bytes = algorithmic_bytes(cfgs[:len(cfgs) // args.replicas], sims)

    # e2e: one-shot C-ABI call, host (pinned) buffers in and out
    p_cfgs, k1 = pinned_like(cfgs)
    p_seeds, k2 = pinned_like(seeds)
    for _ in range(1):
        dev.score(tree, p_cfgs, p_seeds, cands=cands, want_senders=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e_sims, _, _ = dev.score(tree, p_cfgs, p_seeds, cands=cands, want_senders=False)
        score = float(e_sims["utility"].sum())  # the step's result, read on the host
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert e_sims.tobytes() == sims.tobytes(), "e2e path and staged path disagree"
    h2d = int(p_cfgs.nbytes + p_seeds.nbytes + tree.nodes.nbytes * 96 // 112 + tree.leaves.nbytes + len(cfgs) * 4)
    d2h = int(e_sims.nbytes)

    # reduce over ranks: max time, summed work; the per-rank score crosses NVLink (tiny)
    t = torch.tensor([kernel_ms, e2e_s], dtype=torch.float64, device="cuda")
    w = torch.tensor([float(pk), float(ticks), float(len(sims)), alg_bytes, float(launches), score], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    kernel_ms_max, e2e_s_max = t.tolist()
    pk_all, ticks_all, sims_all, alg_all, launches_all, score_all = w.tolist()

    if rank == 0:
        secs = kernel_ms_max / 1e3
        value = pk_all * args.steps / secs
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg_all / n_gpus / (kernel_ms_max / args.steps / 1e3) / 1e9  # per GPU, per launch
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        scfg, sseed, sample = cpu_sample(args, tree, cfgs, seeds)
        cpu = None
        if n_gpus == 1:
            v, kind, cores, dt, ev, _ = run_cpu(args, tree, scfg, sseed)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "seconds": dt}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": kernel_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": {"value": pk_all * args.steps / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches_all),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic,
                             "dram_gbs": (traffic / (kernel_ms_max / args.steps / 1e3) / 1e9) if traffic else None,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                             "note": "achieved = SURVEY.md 8d algorithmic bytes (a design that keeps per-sender state in HBM: 44n+32n_on B "
                                     "per event tick + 32 B per packet sent + 240 B per ack) / step time. This kernel caches event minima and "
                                     "masks in registers/shared memory, so its real DRAM traffic (`traffic`, ncu dram bytes of the step's one "
                                     "simulation launch, profiles/traffic.json) is ~9x smaller and `frac` can exceed 1; `dram_gbs` is the HBM "
                                     "bandwidth actually used. The path is latency/issue bound (ncu --set full, profiles/r01_ncu_full_final.txt: "
                                     "30% issue slots, 14.4/32 lanes, long-scoreboard 6.0 cycles per issued instruction; captured with 4 resident "
                                     "blocks per SM, the library now builds with 3), not HBM bound."},
                "cpu_baseline": cpu,
                "event_ticks_per_sec": ticks_all * args.steps / secs, "sims_per_sec": sims_all * args.steps / secs,
                "evals_per_sec": sims_all / 800.0 * args.steps / secs, "score_sum": score_all}
        if cands is not None:
            line["candidates"] = int(len(cands))
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

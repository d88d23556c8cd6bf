#!/usr/bin/env python3
"""bench.py — BASELINE.json's metric on BASELINE.json's configs, one JSON line on stdout.

metric   : simulated packet-events/sec (packet-event = one packet sent by a Rat or one ack
           consumed), whole job over all ranks
workloads: evaluator-pass (default, the line of record) — BASELINE configs[1]: the default Remy ConfigRange
           Evaluator pass (link 1..2 ppt step 0.25 x rtt 100..200 step 25 x nsrc 1..32 = 800 NetConfigs,
           on = off = 5000 ms, infinite buffer, no loss, fixed WhiskerTree), batched over R evaluator seeds
           per GPU so that one step fills the machine ("scaling": "weak").
           improve-step — BASELINE configs[2]: every Whisker::next_generation candidate of the most used
           whisker x the 800 NetConfigs as ONE batch.  With --scaling strong that one batch is partitioned
           over the ranks by config (remy_b200/sharding.py): each rank scores every candidate on its configs,
           and only per-simulation utilities (8 bytes each) cross NVLink, summed per candidate in config
           order on every rank (src/evaluator.cc:96), so the scores are byte-identical for every N.
           policy-sweep — BASELINE configs[4]: 2^20 random NetConfigs x the candidates, partitioned the same way.
           One step = one scoring pass over the rank's batch.
value    : device time of the kernel launches of a step (one simulation launch + one overflow scan; CUDA events
           on the launching stream, inside the library), inputs already resident in HBM
e2e      : the same pass through the one-shot C-ABI call with HOST buffers
           (remy_b200_score_batch: H2D copies, scratch allocation, kernels, D2H results)
parity   : the CPU leg scores a sample of the very simulations the GPU just ran; their counters, PRNG end
           states, event times and utilities are compared in the run ("parity_checked", "parity_ok"), and a
           mismatch fails the bench.
--impl reference : the reference's own CPU implementation (oracle/_ref, built from /root/reference's sources)
           on all host threads, on a bounded sample of the workload (--cpu-shape pool: a fixed pool of
           hardware_concurrency() workers over the (candidate, config) items; async: the reference's own shape,
           one std::async thread per candidate scoring its configs serially, src/breeder.cc:57-69).

Launch: python bench.py --gpus N --steps K --warmup W         (N == 1)
        python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from remy_b200 import abi, sharding, workloads  # noqa: E402

METRIC = "sim_packet_events_per_sec"
UNIT = "packet-events/s"
INT_FIELDS = ("event_ticks", "packets_sent", "packets_received", "link_queued", "delay_inflight", "prng_state", "error")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build_hash():
    """Identifies the kernel build a profile capture belongs to: sha256 over the CUDA sources and the ABI header."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "remy_b200", "csrc")
    files = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".h")))
    for f in files + [os.path.join(ROOT, "include", "remy_b200.h")]:
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def improve_step_inputs(args, dev):
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
    order = np.argsort(-counts[0].astype(np.int64), kind="stable")
    for leaf in order:
        l = tree.leaves[leaf]
        if 0 <= l["window_increment"] <= 256 and 0 <= l["window_multiple"] <= 1 and 0.25 <= l["intersend"] <= 3:
            cands = ht.next_generation(int(leaf))
            return tree, cfgs, seeds, cands, int(leaf)
    raise SystemExit("no whisker with an optimizable action")


FISH_LAMBDAS, FISH_CUTS = [0.02, 0.01, 0.004], [2.0, 20.0]   # batches of five packets per ms, by rtt_diff (ms)


def evaluator_pass_inputs(args, rank, kind="rat"):
    tree = abi.fin_tree(FISH_LAMBDAS, FISH_CUTS) if kind == "fish" else workloads.load_golden_tree(args.tree)
    base = workloads.default_config_range(ticks=args.ticks)
    if kind == "flows":
        base["mean_on_duration"] = 80.0      # mean flow length in packets (the figure4 config files)
        base["mean_off_duration"] = 500.0
    cfgs = np.tile(base, args.replicas)
    # every (rank, replica) is its own Evaluator seed; seeds within a replica follow seeds_for()
    seeds = np.concatenate([workloads.seeds_for(1000 + rank * args.replicas + r, len(base)) for r in range(args.replicas)])
    return tree, cfgs, seeds


def per_packet_bytes(sims):
    """Compulsory queue and sender-state traffic of the packets themselves: a 16-byte link record
    written and read once per packet sent; per delivered packet 32 B in the delay line + 208 B of Memory,
    Utility and action state."""
    return float((sims["packets_sent"] * 32.0 + sims["packets_received"] * 240.0).sum())


def state_in_hbm_bytes(cfgs, cfg_index, sims):
    """SURVEY.md §8d's model of a design that keeps per-sender state in HBM: additionally 44 n + 32 n_on bytes per
    event tick.  Reported for reference only: this kernel keeps that state on chip."""
    n = cfgs["num_senders"].astype(np.float64)
    n_on = n * cfgs["mean_on_duration"] / (cfgs["mean_on_duration"] + cfgs["mean_off_duration"])
    return float((sims["event_ticks"] * (44.0 * n[cfg_index] + 32.0 * n_on[cfg_index])).sum()) + per_packet_bytes(sims)


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


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_cpu(tree, cfgs, seeds, cands=None, shape="pool", repeats=1, sender="rat"):
    """Times the reference's own CPU implementation (oracle/_ref) or, if it was not built, the oracle port, on all
    host threads; best of `repeats`.  Returns a dict (value = packet-events/s) and the per-simulation results."""
    from oracle import bindings as checkers  # the CPU baseline leg is the one place bench.py executes oracle/
    threads = os.cpu_count() or 1
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libremy_ref.so")
    if os.path.exists(ref_path):
        lib, kind = checkers.Reference(ref_path), "reference"
    else:
        import __graft_entry__ as g
        g.build_checkers()
        lib, kind, shape = checkers.Oracle(), "port", "pool"
    best, sims = None, None
    for _ in range(max(1, repeats)):
        t0 = time.perf_counter()
        if sender != "rat":
            f = lib.score_fish if sender == "fish" else lib.score_flows
            sims, _, _ = f(tree, cfgs, seeds, cands=cands, want_senders=False, threads=threads)
        elif shape == "async":
            sims = lib.score_async_per_candidate(tree, cfgs, seeds, cands=cands)
        else:
            sims, _, _ = lib.score(tree, cfgs, seeds, cands=cands, want_senders=False, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    ev = int((sims["packets_sent"] + sims["packets_received"]).sum())
    n_c = 1 if cands is None else len(cands)
    used = min(threads, n_c) if shape == "async" else threads
    return {"value": ev / best, "unit": UNIT, "cores": used, "host_threads": threads, "cpu_model": cpu_model(), "kind": kind,
            "shape": ("one std::async thread per candidate, configs serial (src/breeder.cc:57-69)" if shape == "async"
                      else "fixed pool of hardware_concurrency() workers over (candidate, config) items; drives the reference's "
                           "classes through oracle/ref_driver.cc's loop, pinned to the public score() by tests/test_oracle.py"),
            "seconds": best, "best_of": max(1, repeats), "event_ticks_per_sec": int(sims["event_ticks"].sum()) / best,
            "sims": int(len(sims))}, sims


def check_parity(gpu_sims, cpu_sims):
    """GPU results of the sampled simulations against the CPU leg's: integers, PRNG end state and event times
    bit-exact, utilities within 1e-9 relative (the tolerance north_star states)."""
    bad = []
    for f in INT_FIELDS:
        if not (gpu_sims[f] == cpu_sims[f]).all():
            bad.append(f)
    if not (gpu_sims["last_tick"] == cpu_sims["last_tick"]).all():
        bad.append("last_tick")
    rel = np.abs(gpu_sims["utility"] - cpu_sims["utility"]) <= 1e-9 * np.abs(cpu_sims["utility"])
    if not rel.all():
        bad.append("utility")
    return bad


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
    ap.add_argument("--cpu-stride", type=int, default=4, help="CPU leg: every n-th NetConfig of one replica")
    ap.add_argument("--cpu-shape", default="pool", choices=["pool", "async"])
    ap.add_argument("--cpu-repeats", type=int, default=1)
    ap.add_argument("--workload", default="evaluator-pass", choices=["evaluator-pass", "improve-step", "policy-sweep", "fish-pass", "flows-pass"],
                    help="fish-pass: the evaluator pass with Fish senders over a FinTree (Evaluator<FinTree>::score, src/evaluator.cc:105-132); "
                         "flows-pass: with Rat senders behind ByteSwitchedSender, flows of mean 80 packets (src/sendergang.cc:108-138)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="improve-step / policy-sweep: strong = ONE batch partitioned over the ranks (default for them); "
                         "weak = every rank its own batch (default for evaluator-pass)")
    ap.add_argument("--sweep-configs", type=int, default=1 << 20, help="policy-sweep: number of random NetConfigs")
    ap.add_argument("--sweep-candidates", type=int, default=1)
    ap.add_argument("--parity-sample", type=int, default=200, help="simulations the CPU leg re-scores and compares (0 = off)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline / parity leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer (one-shot C-ABI) leg: for the very large sweeps")
    ap.add_argument("--parity-at-scale", action="store_true", help="N > 1: rank 0 still re-scores a sample of ITS shard on the CPU")
    args = ap.parse_args()
    pass_like = args.workload in ("evaluator-pass", "fish-pass", "flows-pass")
    if args.scaling is None:
        args.scaling = "weak" if pass_like else "strong"
    kind = {"fish-pass": "fish", "flows-pass": "flows"}.get(args.workload, "rat")

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}")
    n_gpus = max(world, 1)
    l2_note = "per-slot simulation state (>200 MB) exceeds the 126 MB L2; a 256 MiB buffer is also written between steps"

    # ------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        if not pass_like:
            tree = workloads.load_golden_tree(args.tree)
            base = workloads.default_config_range(ticks=args.ticks)
            cands = workloads.candidate_grid(0, base=(2, 0.9, 1.0), n=(os.cpu_count() or 1))
            idx = np.arange(0, len(base), 8)
            scfg, sseed = base[idx], workloads.seeds_for(2026, len(base))[idx]
            sample = (f"{len(cands)} candidates (one per host thread) x every 8th NetConfig of the default ConfigRange "
                      f"({len(idx)} configs), {args.ticks:g} ticks")
            workload = f"BASELINE configs[2] shape: candidates x NetConfigs of one improve step, {args.ticks:g} ticks (CPU sample)"
        else:
            tree, cfgs, seeds = evaluator_pass_inputs(args, 0, kind)
            n_base = len(cfgs) // args.replicas
            idx = np.arange(0, n_base, args.cpu_stride)
            scfg, sseed, cands = cfgs[idx], seeds[idx], None
            sample = f"every {args.cpu_stride}th NetConfig of one 800-config replica ({len(idx)} simulations, {args.ticks:g} ticks each)"
            workload = (f"BASELINE configs[1]: default Remy ConfigRange (800 NetConfigs: link 1-2 ppt x rtt 100-200 ms x nsrc 1-32, "
                        f"exp on/off 5 s, infinite buffer) x {args.replicas} evaluator seeds per GPU, fixed WhiskerTree {args.tree}, "
                        f"{args.ticks:g} ticks")
        config = {"workload": workload, "ticks": args.ticks, "tree": args.tree, "l2": l2_note}
        for _ in range(min(args.warmup, 1)):
            run_cpu(tree, scfg[:32], sseed[:32], cands, args.cpu_shape, 1, kind)
        tot_t, tot_ev, tot_ticks, cpu = 0.0, 0, 0, None
        for _ in range(args.steps):
            cpu, sims = run_cpu(tree, scfg, sseed, cands, args.cpu_shape, 1, kind)
            tot_t += cpu["seconds"]; tot_ev += int((sims["packets_sent"] + sims["packets_received"]).sum())
            tot_ticks += int(sims["event_ticks"].sum())
        value = tot_ev / tot_t
        cpu.update({"value": value, "sample": sample})
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "event_ticks_per_sec": tot_ticks / tot_t, "cpu_baseline": cpu,
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

    # ---- the rank's batch: (tree, candidates, this rank's configs) ----
    cands, shard, layout, n_cfgs_all = None, None, None, None
    if pass_like:
        tree, cfgs, seeds = evaluator_pass_inputs(args, rank, kind)
        workload = (f"BASELINE configs[1]: default Remy ConfigRange (800 NetConfigs: link 1-2 ppt x rtt 100-200 ms x nsrc 1-32, "
                    f"exp on/off 5 s, infinite buffer) x {args.replicas} evaluator seeds per GPU, fixed WhiskerTree {args.tree}, "
                    f"{args.ticks:g} ticks")
        if kind == "fish":
            workload = (f"Evaluator<FinTree> pass: the default ConfigRange (800 NetConfigs) x {args.replicas} evaluator seeds per GPU, Fish senders, "
                        f"FinTree lambda {FISH_LAMBDAS} split at rtt_diff {FISH_CUTS}, {args.ticks:g} ticks")
        elif kind == "flows":
            workload = (f"default ConfigRange (800 NetConfigs) x {args.replicas} evaluator seeds per GPU, Rat senders behind ByteSwitchedSender "
                        f"(flows of mean 80 packets, idle 500 ms), WhiskerTree {args.tree}, {args.ticks:g} ticks")
        n_units = 800.0
    else:
        if args.workload == "improve-step":
            tree, cfgs, seeds, cands, leaf = improve_step_inputs(args, dev)
            workload = (f"BASELINE configs[2]: one RatBreeder improve fan-out - {len(cands)} Whisker::next_generation candidates of "
                        f"leaf {leaf} of {args.tree} x 800 NetConfigs (default ConfigRange), {args.ticks:g} ticks")
        else:
            tree = workloads.load_golden_tree(args.tree)
            cfgs = workloads.random_sweep(args.sweep_configs, ticks=args.ticks)
            seeds = workloads.seeds_for(2026, len(cfgs))
            cands = workloads.candidate_grid(0, base=(2, 0.9, 1.0), n=args.sweep_candidates) if args.sweep_candidates > 1 else None
            workload = (f"BASELINE configs[4]: {len(cfgs)} random NetConfigs (link U[0.5,10] ppt, delay U[25,200] ms, nsrc U{{1..32}}, buffer in "
                        f"{{10, 0.1 BDP, BDP, infinite}}, loss in {{0, .01, .05}}, on = off = 5 s) x {args.sweep_candidates} candidate(s), "
                        f"{args.ticks:g} ticks, {args.tree}")
        n_cfgs_all = len(cfgs)
        n_units = float(n_cfgs_all)
        if args.scaling == "strong" and world > 1:
            # ONE batch over all ranks: deal the cost-sorted configs round-robin (every rank keeps every candidate)
            layout = sharding.shard_layout(cfgs, world)
            shard = layout[rank]
            cfgs, seeds = cfgs[shard], seeds[shard]
            workload += f"; ONE batch partitioned over {world} GPUs by config"
        elif world > 1:
            workload += "; every GPU scores its own copy of the batch"
    n_cands = 1 if cands is None else len(cands)
    config = {"workload": workload, "sims_per_gpu": int(n_cands * len(cfgs)), "ticks": args.ticks, "tree": args.tree, "l2": l2_note}

    # ---- value: staged API, inputs resident in HBM ----
    batch = dev.create_batch(tree, cfgs, seeds, cands=cands, want_senders=False, want_use_counts=False, fish=kind == "fish", flows=kind == "flows")
    one_shot = {"rat": dev.score, "fish": dev.score_fish, "flows": dev.score_flows}[kind]
    for _ in range(args.warmup):
        batch.run(timed=True)
        flush_l2()
    sampler = ClockSampler(local_rank)
    barrier()
    kernel_ms, launches, rerun = 0.0, 0, 0
    for _ in range(args.steps):
        kernel_ms += batch.run(timed=True)  # CUDA events on the library's launching stream
        launches += batch.launches()
        rerun = batch.rerun()
        flush_l2()
    barrier()
    clocks = sampler.stop()
    sims, _, _ = batch.fetch()
    batch.destroy()
    assert (sims["error"] == 0).all(), np.unique(sims["error"])
    pk = int((sims["packets_sent"] + sims["packets_received"]).sum())
    ticks = int(sims["event_ticks"].sum())
    cfg_index = np.arange(len(sims)) % len(cfgs)
    pp_bytes = per_packet_bytes(sims)
    model_bytes = state_in_hbm_bytes(cfgs, cfg_index, sims)

    # ---- e2e: one-shot C-ABI call, host (pinned) buffers in and out ----
    p_cfgs, k1 = pinned_like(cfgs)
    p_seeds, k2 = pinned_like(seeds)
    h2d = int(p_cfgs.nbytes + p_seeds.nbytes + tree.nodes.nbytes * 96 // 112 + tree.leaves.nbytes + len(cfgs) * 4)
    d2h = int(sims.nbytes)
    if args.no_e2e:
        e2e_s, score = float("nan"), float(sims["utility"].sum())
    else:
        one_shot(tree, p_cfgs, p_seeds, cands=cands, want_senders=False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e_sims, _, _ = one_shot(tree, p_cfgs, p_seeds, cands=cands, want_senders=False)
            score = float(e_sims["utility"].sum())  # the step's result, read on the host
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        assert e_sims.tobytes() == sims.tobytes(), "e2e path and staged path disagree"

    # ---- per-candidate scores: only per-simulation utilities cross NVLink (strong scaling), summed in config order ----
    scores_sha, reduce_ms = None, None
    if cands is not None or args.workload == "policy-sweep":
        util = torch.from_numpy(np.ascontiguousarray(sims["utility"].reshape(n_cands, len(cfgs)))).cuda()
        if shard is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            full = sharding.gather_utilities(util, n_cfgs_all, rank, world, dist, layout)
            ev1.record(); torch.cuda.synchronize()
            reduce_ms = ev0.elapsed_time(ev1)
        else:
            full = util
        cand_scores = sharding.scores_in_config_order(full.cpu().numpy())
        scores_sha = hashlib.sha256(cand_scores.tobytes()).hexdigest()[:16]

    # ---- reduce over ranks: max time, summed work ----
    t = torch.tensor([kernel_ms, e2e_s], dtype=torch.float64, device="cuda")
    w = torch.tensor([float(pk), float(ticks), float(len(sims)), pp_bytes, float(launches), score, model_bytes, float(rerun)],
                     dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    kernel_ms_max, e2e_s_max = t.tolist()
    pk_all, ticks_all, sims_all, pp_all, launches_all, score_all, model_all, rerun_all = w.tolist()

    # ---- CPU leg on rank 0 (N = 1 only): baseline timing + in-run parity on the same simulations ----
    cpu, parity_n, parity_bad = None, 0, []
    if rank == 0 and (n_gpus == 1 or args.parity_at_scale) and not args.no_cpu:
        rng = np.random.default_rng(7)
        if pass_like:
            n_base = len(cfgs) // args.replicas
            kidx = np.arange(0, n_base, args.cpu_stride)
            sample = f"every {args.cpu_stride}th NetConfig of one 800-config replica ({len(kidx)} simulations, {args.ticks:g} ticks each)"
            scand = None
        else:
            kidx = np.sort(rng.choice(len(cfgs), size=min(len(cfgs), max(1, args.parity_sample)), replace=False))
            scand = cands
            sample = f"{len(kidx)} random NetConfigs of the batch x {n_cands} candidate(s) ({args.ticks:g} ticks each)"
            if cands is not None and len(cands) > 8:   # bound the CPU work: a few candidates on a few configs
                csel = np.sort(rng.choice(len(cands), size=8, replace=False))
                kidx = kidx[:max(1, args.parity_sample // 8)]
                scand = cands[csel]
                sample = f"8 random candidates x {len(kidx)} random NetConfigs of the batch ({args.ticks:g} ticks each)"
        cpu, c_sims = run_cpu(tree, cfgs[kidx], seeds[kidx], scand, args.cpu_shape, args.cpu_repeats, kind)
        cpu["sample"] = sample
        if args.parity_sample > 0:
            g = sims.reshape(n_cands, len(cfgs))
            if scand is not None and scand is not cands:
                g = g[csel]
            g = g[:, kidx].reshape(-1)
            parity_n = len(g)
            parity_bad = check_parity(g, c_sims)

    if rank == 0:
        secs = kernel_ms_max / 1e3
        step_s = secs / args.steps
        value = pk_all * args.steps / secs
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # measured DRAM traffic of this build's launch on this workload, if a capture of the same build exists
        bh = build_hash()
        wkey = f"{args.workload}/{args.ticks:g}/{args.replicas if pass_like else n_cands}/{args.tree}/{config['sims_per_gpu']}"
        log(f"traffic key: {wkey}   build hash: {bh}")
        traffic, counters, tnote = None, None, "no ncu capture of this build and workload under profiles/traffic.json"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("build_hash") == bh and tj.get("workload_key") == wkey:
                traffic = float(tj["dram_bytes_per_launch"])
                tnote = tj.get("source", "profiles/traffic.json")
            elif tj.get("build_hash") != bh:
                tnote = f"profiles/traffic.json was captured on build {tj.get('build_hash')}, this is {bh}: refused"
            else:
                tnote = f"profiles/traffic.json is for workload {tj.get('workload_key')}, this is {wkey}: refused"
            if tj.get("build_hash") == bh:
                counters = tj.get("counters")
        except (OSError, ValueError, KeyError):
            pass
        pp_gbs = pp_all / n_gpus / step_s / 1e9
        if traffic is not None:
            achieved, basis = traffic / step_s / 1e9, "measured DRAM bytes of the step's simulation launch (ncu, same build) / launch time"
        else:
            achieved, basis = pp_gbs, "per-packet algorithmic bytes (32 B per packet sent + 240 B per ack, SURVEY.md 8d) / launch time"
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "basis": basis, "traffic_note": tnote, "build_hash": bh,
                    "per_packet_algorithmic_gbs": pp_gbs, "per_packet_frac": pp_gbs / peak,
                    "state_in_hbm_model_gbs": model_all / n_gpus / step_s / 1e9,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                    "kernel_counters": counters,
                    "note": "The kernel is bound by memory latency and issue slots, not by HBM bandwidth (kernel_counters: ncu --set full of the "
                            "same build, profiles/). state_in_hbm_model_gbs is SURVEY.md 8d's figure for a design that keeps per-sender state in "
                            "HBM (44n+32n_on B per event tick on top of the per-packet bytes); this kernel keeps that state on chip, so the "
                            "figure can exceed the peak and is not a roofline fraction."}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": kernel_ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": ({"value": pk_all * args.steps / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
                        if not args.no_e2e else None),
                "gpu_launches": int(launches_all), "roofline": roofline, "cpu_baseline": cpu,
                "parity_checked": parity_n, "parity_ok": (not parity_bad) if parity_n else None,
                "event_ticks_per_sec": ticks_all * args.steps / secs, "sims_per_sec": sims_all * args.steps / secs,
                "evals_per_sec": sims_all / n_units * args.steps / secs, "sims_rerun_with_larger_rings": int(rerun_all),
                "score_sum": score_all}
        if cands is not None:
            line["candidates"] = int(len(cands))
        if scores_sha is not None:
            line["candidate_scores_sha256"] = scores_sha
            line["score_gather_ms"] = reduce_ms
        emit(line)
        if parity_bad:
            log(f"bench.py: PARITY FAILURE on the in-run sample: fields {parity_bad} differ between the GPU and the CPU reference")
    if world > 1:
        dist.destroy_process_group()
    return 1 if parity_bad else 0


if __name__ == "__main__":
    sys.exit(main())

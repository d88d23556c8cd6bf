"""ctypes bindings of the two CPU checkers: oracle/liboracle.so (the C restatement of the reference's
algorithm) and oracle/_ref/libremy_ref.so (the reference's own sources, compiled by oracle/build_ref.sh).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke(), tools/ and bench.py's CPU baseline leg import
this module; the product package (remy_b200/) does not, and remy_b200.abi.Device never loads either library.
"""
import ctypes as C
import os

import numpy as np

from remy_b200.abi import (NUM_AXES, REPO_ROOT, FlatTree, FlatTreeC, _PTR, _ScoreLib, _ptr, _score_with, leaf_action_dtype,
                           net_config_dtype, sim_result_dtype, tree_node_dtype)


class Oracle(_ScoreLib):
    """oracle/liboracle.so — CPU restatement (test infrastructure)."""

    def __init__(self, path=None):
        super().__init__(path or os.path.join(REPO_ROOT, "oracle", "liboracle.so"),
                         "remy_oracle_score_batch", (C.c_uint32,))

    def score(self, tree, cfgs, seeds, cands=None, want_senders=True, want_use_counts=False, threads=1):
        return super().score(tree, cfgs, seeds, cands, want_senders, want_use_counts, extra=(threads,))

    def score_fish(self, tree, cfgs, seeds, cands=None, want_senders=True, want_use_counts=False, threads=1):
        """Evaluator<FinTree>::score restated: Fish senders over a FinTree (leaf lambda in `intersend`)."""
        return _score_with(self.lib.remy_oracle_score_batch_kind, self.path, tree, cfgs, seeds, cands, want_senders, want_use_counts,
                           [(C.c_uint32, threads), (C.c_int, 1)])

    def score_flows(self, tree, cfgs, seeds, cands=None, want_senders=True, want_use_counts=False, threads=1):
        """Rat senders behind ByteSwitchedSender (src/sendergang.cc:108-138, 256-278) restated: flows of a random length in packets."""
        return _score_with(self.lib.remy_oracle_score_batch_kind, self.path, tree, cfgs, seeds, cands, want_senders, want_use_counts,
                           [(C.c_uint32, threads), (C.c_int, 2)])

    def trace_sim(self, tree, cfg, seed, cap=1 << 22, fish=False):
        """Every use_whisker / use_fin call of one simulation, in call order: (sim result, leaf[n], mem[n, 6])."""
        f = self.lib.remy_oracle_trace_sim_kind
        f.restype = C.c_uint64
        f.argtypes = [C.POINTER(FlatTreeC), _PTR, C.c_uint32, C.c_int, _PTR, _PTR, C.c_uint64]
        cfg = np.ascontiguousarray(cfg, dtype=net_config_dtype).reshape(1)
        sim = np.zeros(1, dtype=sim_result_dtype)
        buf = np.zeros(cap * 7, dtype=np.float64)
        n = int(f(tree.byref(), _ptr(cfg), int(seed), int(fish), _ptr(sim), _ptr(buf), cap))
        if n > cap:
            return self.trace_sim(tree, cfg, seed, cap=n, fish=fish)
        rec = buf[:n * 7].reshape(n, 7)
        return sim[0], rec[:, 0].astype(np.uint32), np.ascontiguousarray(rec[:, 1:])

    def trace_medians(self, tree, cfgs, seeds, fish=False):
        """score(tree, seeds[k], {cfgs[k]}, trace=true) for k = 0.. on one tree -> (sims, use counts,
        medians[n_leaves, 6]) of the P^2 estimators MemoryRange::track feeds."""
        L = self.lib
        L.remy_oracle_p2_sizeof.restype = C.c_size_t
        sz = L.remy_oracle_p2_sizeof()
        L.remy_oracle_p2_init.argtypes = [_PTR]
        L.remy_oracle_p2_median.argtypes = [_PTR]
        L.remy_oracle_p2_median.restype = C.c_double
        L.remy_oracle_trace_medians_kind.argtypes = [C.POINTER(FlatTreeC), _PTR, _PTR, C.c_uint32, C.c_int, _PTR, _PTR, _PTR]
        cfgs = np.ascontiguousarray(cfgs, dtype=net_config_dtype)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        nl = tree.n_leaves
        acc = np.zeros(nl * NUM_AXES * sz, dtype=np.uint8)
        for i in range(nl * NUM_AXES):
            L.remy_oracle_p2_init(acc.ctypes.data + i * sz)
        sims = np.zeros(len(cfgs), dtype=sim_result_dtype)
        counts = np.zeros(nl, dtype=np.uint32)
        L.remy_oracle_trace_medians_kind(tree.byref(), _ptr(cfgs), _ptr(seeds), len(cfgs), int(fish), _ptr(sims), _ptr(counts), _ptr(acc))
        med = np.array([L.remy_oracle_p2_median(acc.ctypes.data + i * sz) for i in range(nl * NUM_AXES)]).reshape(nl, NUM_AXES)
        return sims, counts, med


class Reference(_ScoreLib):
    """oracle/_ref/libremy_ref.so — the reference's own sources (test infrastructure)."""

    def __init__(self, path=None):
        super().__init__(path or os.path.join(REPO_ROOT, "oracle", "_ref", "libremy_ref.so"),
                         "remy_ref_score_batch", (C.c_uint32, C.c_int))

    def score(self, tree, cfgs, seeds, cands=None, want_senders=True, want_use_counts=False, threads=1, mode=0):
        return super().score(tree, cfgs, seeds, cands, want_senders, want_use_counts, extra=(threads, mode))

    def score_async_per_candidate(self, tree, cfgs, seeds, cands=None, mode=0):
        """The reference's own fan-out shape (src/breeder.cc:57-69): one std::async thread per candidate, each
        scoring all configs serially.  Returns the per-simulation results."""
        f = self.lib.remy_ref_score_batch_async
        f.restype = C.c_int
        f.argtypes = [C.POINTER(FlatTreeC), _PTR, C.c_uint32, _PTR, _PTR, C.c_uint32, _PTR, C.c_int]
        cfgs = np.ascontiguousarray(cfgs, dtype=net_config_dtype)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        n_cands = 1 if cands is None else len(cands)
        if cands is not None:
            from remy_b200.abi import leaf_override_dtype
            cands = np.ascontiguousarray(cands, dtype=leaf_override_dtype)
        sims = np.zeros(n_cands * len(cfgs), dtype=sim_result_dtype)
        rc = f(tree.byref(), _ptr(cands), n_cands, _ptr(cfgs), _ptr(seeds), len(cfgs), _ptr(sims), mode)
        if rc != 0:
            raise RuntimeError(f"remy_ref_score_batch_async returned {rc}")
        return sims

    def score_fish(self, tree, cfgs, seeds, cands=None, want_senders=True, want_use_counts=False, threads=1, mode=0):
        """The reference's own Fish / FinTree code (src/evaluator.cc:105-132)."""
        return _score_with(self.lib.remy_ref_score_batch_fish, self.path, tree, cfgs, seeds, cands, want_senders, want_use_counts,
                           [(C.c_uint32, threads), (C.c_int, mode)])

    def score_flows(self, tree, cfgs, seeds, cands=None, want_senders=True, want_use_counts=False, threads=1):
        """The reference's own ByteSwitchedSender template instantiated for Rat."""
        return _score_with(self.lib.remy_ref_score_batch_flows, self.path, tree, cfgs, seeds, cands, want_senders, want_use_counts,
                           [(C.c_uint32, threads)])

    def score_chained(self, tree, cfgs, prng_seed):
        """The reference's public Evaluator<WhiskerTree>::score on several configs in ONE call (one chained PRNG,
        src/evaluator.cc:84,92-94) -> (score, tput_delay[n_cfgs, stride, 2], use counts)."""
        f = self.lib.remy_ref_score_chained
        f.restype = C.c_int
        f.argtypes = [C.POINTER(FlatTreeC), _PTR, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), _PTR, _PTR]
        cfgs = np.ascontiguousarray(cfgs, dtype=net_config_dtype)
        stride = int(cfgs["num_senders"].max())
        td = np.zeros((len(cfgs), stride, 2))
        counts = np.zeros(tree.n_leaves, dtype=np.uint32)
        score = C.c_double()
        rc = f(tree.byref(), _ptr(cfgs), len(cfgs), int(prng_seed), stride, C.byref(score), _ptr(td), _ptr(counts))
        if rc != 0:
            raise RuntimeError(f"remy_ref_score_chained returned {rc}")
        return score.value, td, counts

    def trace_split(self, tree, cfgs, seeds, generation=0, split=True):
        """The reference's own trace == true scoring (per config, on one tree) and, if `split`, the
        body of Breeder::apply_best_split -> (score, use counts, medians[n_leaves, 6], resulting FlatTree)."""
        f = self.lib.remy_ref_trace_split
        f.argtypes = [C.POINTER(FlatTreeC), _PTR, _PTR, C.c_uint32, C.c_uint32, C.c_int, _PTR, _PTR, C.POINTER(C.c_double),
                      _PTR, C.c_uint32, _PTR, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        cfgs = np.ascontiguousarray(cfgs, dtype=net_config_dtype)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        med = np.zeros((tree.n_leaves, NUM_AXES))
        counts = np.zeros(tree.n_leaves, dtype=np.uint32)
        score = C.c_double()
        cap = len(tree.nodes) + 4096
        nodes = np.zeros(cap, dtype=tree_node_dtype)
        leaves = np.zeros(cap, dtype=leaf_action_dtype)
        nn, nl = C.c_uint32(), C.c_uint32()
        rc = f(tree.byref(), _ptr(cfgs), _ptr(seeds), len(cfgs), generation, int(split), _ptr(med), _ptr(counts), C.byref(score),
               _ptr(nodes), cap, _ptr(leaves), cap, C.byref(nn), C.byref(nl))
        if rc != 0:
            raise RuntimeError(f"remy_ref_trace_split returned {rc}")
        return score.value, counts, med, FlatTree(nodes[:nn.value].copy(), leaves[:nl.value].copy())

"""ctypes view of remy_b200/host/libremy_host.so (the C++ mirror of the reference's
Evaluator / WhiskerTree / ConfigRange API) for the tests."""
import ctypes as C
import os
import struct

import numpy as np

from . import abi

_V = C.c_void_p


class HostLib:
    def __init__(self, path=None):
        path = path or os.path.join(abi.REPO_ROOT, "remy_b200", "host", "libremy_host.so")
        self.lib = L = C.CDLL(path)
        L.remy_host_last_error.restype = C.c_char_p
        for name in ("remy_host_tree_default", "remy_host_tree_parse", "remy_host_tree_from_flat"):
            getattr(L, name).restype = _V
        L.remy_host_tree_parse.argtypes = [C.c_char_p, C.c_size_t]
        L.remy_host_tree_from_flat.argtypes = [C.POINTER(abi.FlatTreeC)]
        L.remy_host_tree_free.argtypes = [_V]
        for name in ("remy_host_tree_serialize", "remy_host_tree_str"):
            getattr(L, name).restype = C.c_long
            getattr(L, name).argtypes = [_V, C.c_char_p, C.c_size_t]
        L.remy_host_tree_flatten.argtypes = [_V, _V, C.c_uint32, _V, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.remy_host_next_generation.restype = C.c_long
        L.remy_host_next_generation.argtypes = [_V, C.c_uint32, C.c_int, C.c_int, C.c_int, _V, C.c_uint32]
        L.remy_host_config_range_expand.restype = C.c_long
        L.remy_host_config_range_expand.argtypes = [C.c_char_p, C.c_size_t, C.c_double, _V, C.c_uint32, C.c_char_p, C.c_size_t, C.POINTER(C.c_long)]
        L.remy_host_evaluator_score.argtypes = [_V, C.c_char_p, C.c_size_t, C.c_uint, C.c_double, C.POINTER(C.c_double), _V, C.c_uint32,
                                                _V, C.c_uint32, C.POINTER(C.c_uint32)]
        for name in ("remy_host_solve_problem", "remy_host_answer_roundtrip"):
            getattr(L, name).restype = C.c_long
            getattr(L, name).argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.remy_host_problem_dna.restype = C.c_long
        L.remy_host_problem_dna.argtypes = [_V, C.c_char_p, C.c_size_t, C.c_uint, C.c_char_p, C.c_size_t]
        L.remy_host_improve_whisker.restype = C.c_long
        L.remy_host_improve_whisker.argtypes = [_V, C.c_char_p, C.c_size_t, C.c_uint, C.c_uint32, _V]
        L.remy_host_ratbreeder_improve.argtypes = [_V, C.c_char_p, C.c_size_t, C.c_uint, C.POINTER(C.c_double), _V]
        for name in ("remy_host_fintree_default", "remy_host_fintree_parse", "remy_host_fintree_from_flat"):
            getattr(L, name).restype = _V
        L.remy_host_fintree_parse.argtypes = [C.c_char_p, C.c_size_t]
        L.remy_host_fintree_from_flat.argtypes = [C.POINTER(abi.FlatTreeC)]
        L.remy_host_fintree_free.argtypes = [_V]
        for name in ("remy_host_fintree_serialize", "remy_host_fintree_str"):
            getattr(L, name).restype = C.c_long
            getattr(L, name).argtypes = [_V, C.c_char_p, C.c_size_t]
        L.remy_host_fintree_flatten.argtypes = [_V, _V, C.c_uint32, _V, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.remy_host_fin_next_generation.restype = C.c_long
        L.remy_host_fin_next_generation.argtypes = [_V, C.c_uint32, _V, C.c_uint32]
        L.remy_host_fish_score.argtypes = [_V, C.c_char_p, C.c_size_t, C.c_uint, C.c_double, C.POINTER(C.c_double), _V, C.c_uint32]
        L.remy_host_improve_fin.restype = C.c_long
        L.remy_host_improve_fin.argtypes = [_V, C.c_char_p, C.c_size_t, C.c_uint, C.c_uint32, _V]
        L.remy_host_fish_score_trace.argtypes = [_V, _V, C.c_uint32, C.c_uint, C.POINTER(C.c_double)]
        L.remy_host_fintree_medians.argtypes = [_V, _V, C.c_uint32]
        L.remy_host_fish_apply_best_split.argtypes = [_V, C.c_char_p, C.c_size_t, C.c_uint, C.c_uint]
        L.remy_host_fishbreeder_improve.argtypes = [_V, C.c_char_p, C.c_size_t, C.c_uint, C.POINTER(C.c_double)]
        L.remy_host_p2_median.restype = C.c_double
        L.remy_host_p2_median.argtypes = [_V, C.c_size_t]
        L.remy_host_track.argtypes = [_V, C.c_uint32, _V, C.c_size_t]
        L.remy_host_medians.argtypes = [_V, _V, C.c_uint32]
        L.remy_host_feed_records.argtypes = [_V, _V, C.c_uint64, C.c_uint32]
        L.remy_host_set_counts.argtypes = [_V, _V, C.c_uint32]
        L.remy_host_split_most_used.argtypes = [_V, C.c_uint]
        L.remy_host_score_trace.argtypes = [_V, _V, C.c_uint32, C.c_uint, C.POINTER(C.c_double)]
        L.remy_host_apply_best_split.argtypes = [_V, C.c_char_p, C.c_size_t, C.c_uint, C.c_uint]
        L.remy_host_score_next_generation.restype = C.c_long
        L.remy_host_score_next_generation.argtypes = [_V, C.c_char_p, C.c_size_t, C.c_uint, C.c_double, C.c_uint32, _V, C.c_uint32]

    def err(self):
        return self.lib.remy_host_last_error().decode()

    def set_chained_prng(self, on):
        """the reference's literal PRNG chaining through the configs of a score() call (remy.hh::set_chained_prng)"""
        self.lib.remy_host_set_chained_prng(int(bool(on)))

    def seeds(self, prng_seed, n):
        """Evaluator<WhiskerTree>::seeds(prng_seed, n) of the C++ host mirror."""
        out = np.zeros(n, dtype=np.uint32)
        self.lib.remy_host_seeds.argtypes = [C.c_uint, C.c_void_p, C.c_uint32]
        self._check(self.lib.remy_host_seeds(prng_seed, out.ctypes.data, n) == 0)
        return out

    def p2_median(self, samples):
        a = np.ascontiguousarray(samples, dtype=np.float64)
        return self.lib.remy_host_p2_median(a.ctypes.data, len(a))

    def _check(self, ok):
        if not ok:
            raise RuntimeError(self.err())

    def tree_default(self):
        return HostTree(self, self.lib.remy_host_tree_default())

    def tree_parse(self, data):
        h = self.lib.remy_host_tree_parse(data, len(data))
        self._check(h)
        return HostTree(self, h)

    def tree_from_flat(self, flat):
        h = self.lib.remy_host_tree_from_flat(flat.byref())
        self._check(h)
        return HostTree(self, h)

    def _bytes_call(self, fn, *args):
        n = fn(*args, None, 0)
        self._check(n >= 0)
        buf = C.create_string_buffer(n + 1)
        self._check(fn(*args, buf, n) == n)
        return buf.raw[:n]

    def fintree_default(self):
        return HostFinTree(self, self.lib.remy_host_fintree_default())

    def fintree_parse(self, data):
        h = self.lib.remy_host_fintree_parse(data, len(data))
        self._check(h)
        return HostFinTree(self, h)

    def fintree_from_flat(self, flat):
        h = self.lib.remy_host_fintree_from_flat(flat.byref())
        self._check(h)
        return HostFinTree(self, h)

    def solve_problem(self, problem):
        """Evaluator::parse_problem_and_evaluate -> AnswerBuffers.Outcome bytes (needs a GPU)."""
        return self._bytes_call(self.lib.remy_host_solve_problem, problem, len(problem))

    def answer_roundtrip(self, answer):
        return self._bytes_call(self.lib.remy_host_answer_roundtrip, answer, len(answer))

    def expand(self, range_bytes, carefulness=1.0):
        out = np.zeros(1 << 16, dtype=abi.net_config_dtype)
        dna = C.create_string_buffer(4096)
        n_dna = C.c_long(0)
        n = self.lib.remy_host_config_range_expand(range_bytes, len(range_bytes), carefulness, out.ctypes.data, len(out), dna, 4096, C.byref(n_dna))
        self._check(n >= 0)
        return out[:n].copy(), dna.raw[:n_dna.value]


class HostFinTree:
    """FinTree (remy_b200/host/fish.hh) behind the C wrappers."""

    def __init__(self, host, handle):
        self.host, self.h = host, handle

    def __del__(self):
        if self.h:
            self.host.lib.remy_host_fintree_free(self.h)
            self.h = None

    def _string(self, fn):
        n = fn(self.h, None, 0)
        self.host._check(n >= 0)
        buf = C.create_string_buffer(n + 1)
        fn(self.h, buf, n)
        return buf.raw[:n]

    def serialize(self):
        return self._string(self.host.lib.remy_host_fintree_serialize)

    def str(self):
        return self._string(self.host.lib.remy_host_fintree_str).decode()

    def flatten(self):
        nn, nl = C.c_uint32(0), C.c_uint32(0)
        self.host._check(self.host.lib.remy_host_fintree_flatten(self.h, None, 0, None, 0, C.byref(nn), C.byref(nl)) == 0)
        nodes = np.zeros(nn.value, dtype=abi.tree_node_dtype)
        leaves = np.zeros(nl.value, dtype=abi.leaf_action_dtype)
        self.host._check(self.host.lib.remy_host_fintree_flatten(self.h, nodes.ctypes.data, len(nodes), leaves.ctypes.data, len(leaves),
                                                                 C.byref(nn), C.byref(nl)) == 0)
        return abi.FlatTree(nodes, leaves)

    def next_generation(self, leaf):
        out = np.zeros(64, dtype=np.float64)
        n = self.host.lib.remy_host_fin_next_generation(self.h, leaf, out.ctypes.data, len(out))
        self.host._check(n >= 0)
        return out[:n].copy()

    def score(self, range_bytes, prng_seed, carefulness=1.0):
        """Evaluator<FinTree>::score (needs a GPU) -> (score, use counts)."""
        flat = self.flatten()
        score = C.c_double(0)
        counts = np.zeros(flat.n_leaves, dtype=np.uint32)
        rc = self.host.lib.remy_host_fish_score(self.h, range_bytes, len(range_bytes), prng_seed, carefulness, C.byref(score),
                                                counts.ctypes.data, len(counts))
        self.host._check(rc == 0)
        return score.value, counts

    def score_trace(self, cfgs, prng_seed):
        """Evaluator<FinTree>::score(tree, seed, configs, trace=true, ticks) (needs a GPU)."""
        cfgs = np.ascontiguousarray(cfgs, dtype=abi.net_config_dtype)
        score = C.c_double(0)
        self.host._check(self.host.lib.remy_host_fish_score_trace(self.h, cfgs.ctypes.data, len(cfgs), prng_seed, C.byref(score)) == 0)
        return score.value

    def medians(self, n_leaves):
        out = np.zeros((n_leaves, abi.NUM_AXES))
        self.host._check(self.host.lib.remy_host_fintree_medians(self.h, out.ctypes.data, n_leaves) == 0)
        return out

    def apply_best_split(self, range_bytes, seed, generation=0):
        self.host._check(self.host.lib.remy_host_fish_apply_best_split(self.h, range_bytes, len(range_bytes), seed, generation) == 0)

    def fishbreeder_improve(self, range_bytes, seed):
        score = C.c_double(0)
        self.host._check(self.host.lib.remy_host_fishbreeder_improve(self.h, range_bytes, len(range_bytes), seed, C.byref(score)) == 0)
        return score.value

    def improve_fin(self, range_bytes, prng_seed, leaf):
        """FinImprover::improve until the score stops rising -> (calls, score, lambda, candidates, batches)."""
        out = np.zeros(4, dtype=np.float64)
        n = self.host.lib.remy_host_improve_fin(self.h, range_bytes, len(range_bytes), prng_seed, leaf, out.ctypes.data)
        self.host._check(n >= 0)
        return int(n), float(out[0]), float(out[1]), int(out[2]), int(out[3])


class HostTree:
    def __init__(self, host, handle):
        self.host, self.h = host, handle

    def __del__(self):
        if self.h:
            self.host.lib.remy_host_tree_free(self.h)
            self.h = None

    def _string(self, fn):
        n = fn(self.h, None, 0)
        self.host._check(n >= 0)
        buf = C.create_string_buffer(n + 1)
        fn(self.h, buf, n)
        return buf.raw[:n]

    def serialize(self):
        return self._string(self.host.lib.remy_host_tree_serialize)

    def str(self):
        return self._string(self.host.lib.remy_host_tree_str).decode()

    def flatten(self):
        nn, nl = C.c_uint32(0), C.c_uint32(0)
        self.host._check(self.host.lib.remy_host_tree_flatten(self.h, None, 0, None, 0, C.byref(nn), C.byref(nl)) == 0)
        nodes = np.zeros(nn.value, dtype=abi.tree_node_dtype)
        leaves = np.zeros(nl.value, dtype=abi.leaf_action_dtype)
        self.host._check(self.host.lib.remy_host_tree_flatten(self.h, nodes.ctypes.data, len(nodes), leaves.ctypes.data, len(leaves),
                                                              C.byref(nn), C.byref(nl)) == 0)
        return abi.FlatTree(nodes, leaves)

    def next_generation(self, leaf, oi=True, om=True, os_=True):
        out = np.zeros(512, dtype=abi.leaf_override_dtype)
        n = self.host.lib.remy_host_next_generation(self.h, leaf, int(oi), int(om), int(os_), out.ctypes.data, len(out))
        self.host._check(n >= 0)
        return out[:n].copy()

    def score(self, range_bytes, prng_seed, carefulness=1.0):
        flat = self.flatten()
        score = C.c_double(0)
        counts = np.zeros(flat.n_leaves, dtype=np.uint32)
        td = np.zeros(1 << 18, dtype=np.float64)
        n_td = C.c_uint32(0)
        rc = self.host.lib.remy_host_evaluator_score(self.h, range_bytes, len(range_bytes), prng_seed, carefulness, C.byref(score),
                                                      counts.ctypes.data, len(counts), td.ctypes.data, len(td), C.byref(n_td))
        self.host._check(rc == 0)
        return score.value, counts, td[:n_td.value].reshape(-1, 2).copy()

    def problem_dna(self, range_bytes, prng_seed):
        """Evaluator::DNA(tree) -> ProblemBuffers.Problem bytes."""
        return self.host._bytes_call(self.host.lib.remy_host_problem_dna, self.h, range_bytes, len(range_bytes), prng_seed)

    def improve_whisker(self, range_bytes, prng_seed, leaf):
        """WhiskerImprover::improve repeated until the score stops rising -> (calls, score, (incr, mult, intersend), candidates, batches)."""
        out = np.zeros(6, dtype=np.float64)
        n = self.host.lib.remy_host_improve_whisker(self.h, range_bytes, len(range_bytes), prng_seed, leaf, out.ctypes.data)
        self.host._check(n >= 0)
        return int(n), float(out[0]), (int(out[1]), float(out[2]), float(out[3])), int(out[4]), int(out[5])

    def track(self, leaf, mem):
        """MemoryRange::track of leaf `leaf` with mem[n, 6]."""
        a = np.ascontiguousarray(mem, dtype=np.float64)
        self.host._check(self.host.lib.remy_host_track(self.h, leaf, a.ctypes.data, len(a)) == 0)

    def feed_records(self, records):
        """records[n, words] uint64 in the layout of remy_b200_score_trace."""
        a = np.ascontiguousarray(records, dtype=np.uint64)
        self.host._check(self.host.lib.remy_host_feed_records(self.h, a.ctypes.data, a.shape[0], a.shape[1]) == 0)

    def medians(self, n_leaves):
        out = np.zeros((n_leaves, abi.NUM_AXES))
        self.host._check(self.host.lib.remy_host_medians(self.h, out.ctypes.data, n_leaves) == 0)
        return out

    def set_counts(self, counts):
        a = np.ascontiguousarray(counts, dtype=np.uint32)
        self.host._check(self.host.lib.remy_host_set_counts(self.h, a.ctypes.data, len(a)) == 0)

    def split_most_used(self, generation=0):
        self.host._check(self.host.lib.remy_host_split_most_used(self.h, generation) == 0)

    def score_trace(self, cfgs, prng_seed):
        """Evaluator::score(tree, seed, configs, trace=true, ticks) (needs a GPU)."""
        cfgs = np.ascontiguousarray(cfgs, dtype=abi.net_config_dtype)
        score = C.c_double(0)
        self.host._check(self.host.lib.remy_host_score_trace(self.h, cfgs.ctypes.data, len(cfgs), prng_seed, C.byref(score)) == 0)
        return score.value

    def apply_best_split(self, range_bytes, seed, generation=0):
        self.host._check(self.host.lib.remy_host_apply_best_split(self.h, range_bytes, len(range_bytes), seed, generation) == 0)

    def ratbreeder_improve(self, range_bytes, seed):
        score = C.c_double(0)
        stats = np.zeros(2, dtype=np.float64)
        rc = self.host.lib.remy_host_ratbreeder_improve(self.h, range_bytes, len(range_bytes), seed, C.byref(score), stats.ctypes.data)
        self.host._check(rc == 0)
        return score.value, int(stats[0]), int(stats[1])

    def score_next_generation(self, range_bytes, prng_seed, leaf, carefulness=1.0):
        out = np.zeros(512, dtype=np.float64)
        n = self.host.lib.remy_host_score_next_generation(self.h, range_bytes, len(range_bytes), prng_seed, carefulness, leaf, out.ctypes.data, len(out))
        self.host._check(n >= 0)
        return out[:n].copy()


def encode_range(low, high, incr):
    """RemyBuffers.Range bytes (dna.proto:89-93)."""
    return b"".join(bytes([0xE9 + 8 * i, 0x03]) + struct.pack("<d", v) for i, v in enumerate((low, high, incr)))


def _varint(v):
    out = b""
    while v >= 0x80:
        out += bytes([(v & 0x7F) | 0x80])
        v >>= 7
    return out + bytes([v])


def encode_config_range(link, rtt, nsrc, buf, off, on, ticks, loss=(0, 0, 0), unicorn=False):
    """RemyBuffers.ConfigRange (dna.proto:95-104) or, with unicorn=True, the ConfigRangeUnicorn
    form in which field 77 is a Range (dna.proto:106-119)."""
    def msg(num, body):
        return _varint((num << 3) | 2) + _varint(len(body)) + body
    out = b""
    for num, r in ((71, link), (72, rtt), (73, nsrc), (74, buf), (75, off), (76, on)):
        out += msg(num, encode_range(*r))
    out += msg(77, encode_range(ticks, ticks, 0)) if unicorn else _varint((77 << 3) | 0) + _varint(int(ticks))
    out += msg(78, encode_range(*loss))
    return out

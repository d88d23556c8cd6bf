"""ctypes mirror of include/remy_b200.h (layouts only — no compute here).

Used by tests/ and bench.py to call the C-ABI of the CUDA library
(remy_b200/csrc/libremy_b200.so) and, with the same structs, the checkers under
oracle/.  The product path is the C-ABI library; nothing in this module falls
back to a CPU implementation.
"""
import ctypes as C
import os

import numpy as np

NUM_AXES = 6
MAX_SENDERS = 32
NO_OVERRIDE = 0xFFFFFFFF
BUFFER_INFINITE = 0xFFFFFFFF

ERR_NO_WHISKER = 0x01
ERR_NO_CHILD = 0x02
ERR_LINK_OVERFLOW = 0x04
ERR_DELAY_OVERFLOW = 0x08
ERR_BAD_CONFIG = 0x10
ERR_ASSERT = 0x20

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# numpy dtypes with exactly the C layout (checked against ctypes sizes below)
net_config_dtype = np.dtype([
    ("mean_on_duration", "<f8"), ("mean_off_duration", "<f8"), ("link_ppt", "<f8"),
    ("delay", "<f8"), ("stochastic_loss_rate", "<f8"), ("simulation_ticks", "<f8"),
    ("num_senders", "<u4"), ("buffer_size", "<u4")], align=True)

tree_node_dtype = np.dtype([
    ("lower", "<f8", (NUM_AXES,)), ("upper", "<f8", (NUM_AXES,)),
    ("axis_mask", "<u4"), ("child_begin", "<u4"), ("child_count", "<u4"), ("leaf_index", "<u4")], align=True)

leaf_action_dtype = np.dtype([
    ("window_multiple", "<f8"), ("intersend", "<f8"), ("window_increment", "<i4"), ("generation", "<u4")], align=True)

leaf_override_dtype = np.dtype([
    ("window_multiple", "<f8"), ("intersend", "<f8"), ("window_increment", "<i4"), ("leaf_index", "<u4")], align=True)

sim_result_dtype = np.dtype([
    ("utility", "<f8"), ("last_tick", "<f8"), ("event_ticks", "<u8"), ("packets_sent", "<u8"),
    ("packets_received", "<u8"), ("link_queued", "<u4"), ("delay_inflight", "<u4"),
    ("prng_state", "<u4"), ("error", "<u4")], align=True)

sender_result_dtype = np.dtype([
    ("tick_share_sending", "<f8"), ("total_delay", "<f8"), ("packets_received", "<u4"), ("packets_sent", "<u4")],
    align=True)

assert net_config_dtype.itemsize == 56
assert tree_node_dtype.itemsize == 112
assert leaf_action_dtype.itemsize == 24
assert leaf_override_dtype.itemsize == 24
assert sim_result_dtype.itemsize == 56
assert sender_result_dtype.itemsize == 24


class FlatTreeC(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("leaves", C.c_void_p), ("n_nodes", C.c_uint32), ("n_leaves", C.c_uint32)]


class FlatTree:
    """Host-side flattened WhiskerTree: numpy arrays + the C view of them."""

    def __init__(self, nodes, leaves):
        self.nodes = np.ascontiguousarray(nodes, dtype=tree_node_dtype)
        self.leaves = np.ascontiguousarray(leaves, dtype=leaf_action_dtype)
        self.c = FlatTreeC(self.nodes.ctypes.data, self.leaves.ctypes.data, len(self.nodes), len(self.leaves))

    @property
    def n_leaves(self):
        return len(self.leaves)

    def byref(self):
        return C.byref(self.c)


def default_tree():
    """WhiskerTree() of the reference (src/whiskertree.cc:10-15): one leaf over
    [0, 163840)^axes with the default action (src/whisker.hh:59-66: increment 1,
    multiple 1, intersend 3) and the default 4 active axes (memoryrange.hh:30-31)."""
    nodes = np.zeros(1, dtype=tree_node_dtype)
    nodes[0]["lower"][:] = 0.0
    nodes[0]["upper"][:] = 163840.0
    nodes[0]["axis_mask"] = 0b1111
    leaves = np.zeros(1, dtype=leaf_action_dtype)
    leaves[0] = (1.0, 3.0, 1, 0)
    return FlatTree(nodes, leaves)


def make_configs(rows):
    """rows: iterable of dicts with NetConfig field names (missing -> NetConfig
    defaults, src/network.hh:27-36) plus simulation_ticks."""
    out = np.zeros(len(rows), dtype=net_config_dtype)
    for i, r in enumerate(rows):
        out[i] = (r.get("mean_on_duration", 5000.0), r.get("mean_off_duration", 5000.0),
                  r.get("link_ppt", 1.0), r.get("delay", 150.0), r.get("stochastic_loss_rate", 0.0),
                  r["simulation_ticks"], r.get("num_senders", 8), r.get("buffer_size", BUFFER_INFINITE))
    return out


_PTR = C.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_PTR)


def _score_with(fn, path, tree, cfgs, seeds, cands, want_senders, want_use_counts, extra):
    """Call a function with the shape of remy_b200_score_batch plus `extra` (ctype, value) arguments."""
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(FlatTreeC), _PTR, C.c_uint32, _PTR, _PTR, C.c_uint32, C.c_uint32, _PTR, _PTR, _PTR] + [t for t, _ in extra]
    cfgs = np.ascontiguousarray(cfgs, dtype=net_config_dtype)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    n_cands = 1 if cands is None else len(cands)
    if cands is not None:
        cands = np.ascontiguousarray(cands, dtype=leaf_override_dtype)
    n = n_cands * len(cfgs)
    stride = max(1, min(int(cfgs["num_senders"].max()) if len(cfgs) else 1, 4096))
    sims = np.zeros(n, dtype=sim_result_dtype)
    senders = np.zeros((n, stride), dtype=sender_result_dtype) if want_senders else None
    counts = np.zeros((n_cands, tree.n_leaves), dtype=np.uint32) if want_use_counts else None
    rc = fn(tree.byref(), _ptr(cands), n_cands, _ptr(cfgs), _ptr(seeds), len(cfgs), stride, _ptr(sims), _ptr(senders), _ptr(counts),
            *[v for _, v in extra])
    if rc != 0:
        raise RuntimeError(f"{path}: returned {rc}")
    return sims, senders, counts


def fin_tree(lambdas, cuts=()):
    """A FinTree (src/fintree.cc:7-12) over rtt_diff, the Fish's only signal: one fin per interval of
    [0, 163840) split at `cuts`; a leaf's lambda travels in the `intersend` field."""
    n = len(lambdas)
    assert n == len(cuts) + 1
    nodes = np.zeros(1 if n == 1 else n + 1, dtype=tree_node_dtype)
    leaves = np.zeros(n, dtype=leaf_action_dtype)
    for nd in nodes:
        nd["upper"][:] = 163840.0
        nd["axis_mask"] = 1 << 4
    edges = [0.0] + list(cuts) + [163840.0]
    for i, lam in enumerate(lambdas):
        leaves[i] = (0.0, lam, 0, 0)
        if n > 1:
            nodes[i + 1]["lower"][4], nodes[i + 1]["upper"][4], nodes[i + 1]["leaf_index"] = edges[i], edges[i + 1], i
    if n > 1:
        nodes[0]["child_begin"], nodes[0]["child_count"] = 1, n
    return FlatTree(nodes, leaves)


class _ScoreLib:
    """Shared calling convention of remy_b200_score_batch and the two checkers."""

    def __init__(self, path, symbol, extra_args=()):
        self.path = path
        self.lib = C.CDLL(path)
        self.fn = getattr(self.lib, symbol)
        self.fn.restype = C.c_int
        self.fn.argtypes = [C.POINTER(FlatTreeC), _PTR, C.c_uint32, _PTR, _PTR, C.c_uint32, C.c_uint32,
                            _PTR, _PTR, _PTR] + list(extra_args)

    def score(self, tree, cfgs, seeds, cands=None, want_senders=True, want_use_counts=False, extra=()):
        cfgs = np.ascontiguousarray(cfgs, dtype=net_config_dtype)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        assert len(seeds) == len(cfgs)
        n_cands = 1 if cands is None else len(cands)
        if cands is not None:
            cands = np.ascontiguousarray(cands, dtype=leaf_override_dtype)
        n = n_cands * len(cfgs)
        stride = int(cfgs["num_senders"].max()) if len(cfgs) else 1
        stride = max(1, min(stride, 4096))
        sims = np.zeros(n, dtype=sim_result_dtype)
        senders = np.zeros((n, stride), dtype=sender_result_dtype) if want_senders else None
        counts = np.zeros((n_cands, tree.n_leaves), dtype=np.uint32) if want_use_counts else None
        rc = self.fn(tree.byref(), _ptr(cands), n_cands, _ptr(cfgs), _ptr(seeds), len(cfgs), stride,
                     _ptr(sims), _ptr(senders), _ptr(counts), *extra)
        if rc != 0:
            raise RuntimeError(f"{self.path}: score_batch returned {rc}: {self.last_error()}")
        return sims, senders, counts

    def last_error(self):
        return ""


class Device(_ScoreLib):
    """remy_b200/csrc/libremy_b200.so — the product: CUDA kernels behind the C ABI.
    Raises if the library is missing or no CUDA device can be initialised; there
    is no CPU fallback."""

    def __init__(self, device=0, path=None):
        """device: a CUDA device index, or a list of them (None = all visible): every batch is then
        partitioned over those devices (remy_b200_init_devices)."""
        path = path or os.path.join(REPO_ROOT, "remy_b200", "csrc", "libremy_b200.so")
        if not os.path.exists(path):
            raise RuntimeError(f"CUDA extension not built: {path} (run python -c 'import __graft_entry__ as g; g.build()')")
        super().__init__(path, "remy_b200_score_batch")
        self.lib.remy_b200_last_error.restype = C.c_char_p
        self.lib.remy_b200_init.argtypes = [C.c_int]
        self.lib.remy_b200_abi_version.restype = C.c_uint32
        self.abi_version = self.lib.remy_b200_abi_version()
        if self.abi_version < 2:   # an older build of the library (A/B tooling): one device, no staged Fish batches
            if not isinstance(device, int):
                raise RuntimeError(f"{path}: ABI v{self.abi_version} binds one device only")
            rc = self.lib.remy_b200_init(device)
            if rc != 0:
                raise RuntimeError(f"remy_b200_init({device}) failed with {rc}: {self.last_error()}")
            self.n_devices = 1
            self._bind_batch_api()
            return
        self.lib.remy_b200_init_devices.argtypes = [C.POINTER(C.c_int), C.c_int]
        if isinstance(device, int):
            rc = self.lib.remy_b200_init(device)
        elif device is None:
            rc = self.lib.remy_b200_init_devices(None, 0)
        else:
            arr = (C.c_int * len(device))(*device)
            rc = self.lib.remy_b200_init_devices(arr, len(device))
        if rc != 0:
            raise RuntimeError(f"remy_b200_init({device}) failed with {rc}: {self.last_error()}")
        self.n_devices = self.lib.remy_b200_device_count()
        self._bind_batch_api()

    def _bind_batch_api(self):
        self.lib.remy_b200_batch_create.restype = C.c_int
        self.lib.remy_b200_batch_create.argtypes = [C.POINTER(FlatTreeC), _PTR, C.c_uint32, _PTR, _PTR, C.c_uint32,
                                                    C.c_uint32, C.c_int, C.c_int, C.POINTER(_PTR)]
        self.lib.remy_b200_batch_run.restype = C.c_int
        self.lib.remy_b200_batch_run.argtypes = [_PTR, C.POINTER(C.c_float)]
        self.lib.remy_b200_batch_fetch.restype = C.c_int
        self.lib.remy_b200_batch_fetch.argtypes = [_PTR, _PTR, _PTR, _PTR]
        self.lib.remy_b200_batch_destroy.argtypes = [_PTR]
        for name in ("remy_b200_batch_launches", "remy_b200_batch_rerun", "remy_b200_batch_devices")[:3 if self.abi_version >= 2 else 1]:
            getattr(self.lib, name).restype = C.c_uint32
            getattr(self.lib, name).argtypes = [_PTR]
        if self.abi_version >= 2:
            for name in ("remy_b200_batch_create_fish", "remy_b200_batch_create_flows"):
                getattr(self.lib, name).restype = C.c_int
                getattr(self.lib, name).argtypes = self.lib.remy_b200_batch_create.argtypes

    def last_error(self):
        return self.lib.remy_b200_last_error().decode()

    def score_fish(self, tree, cfgs, seeds, cands=None, want_senders=True, want_use_counts=False):
        """remy_b200_score_batch_fish: Fish senders over a FinTree on the GPU."""
        try:
            return _score_with(self.lib.remy_b200_score_batch_fish, self.path, tree, cfgs, seeds, cands, want_senders, want_use_counts, [])
        except RuntimeError as e:
            raise RuntimeError(f"{e}: {self.last_error()}")

    def score_flows(self, tree, cfgs, seeds, cands=None, want_senders=True, want_use_counts=False):
        """remy_b200_score_batch_flows: Rat senders behind ByteSwitchedSender (mean_on_duration = mean flow length in packets)."""
        try:
            return _score_with(self.lib.remy_b200_score_batch_flows, self.path, tree, cfgs, seeds, cands, want_senders, want_use_counts, [])
        except RuntimeError as e:
            raise RuntimeError(f"{e}: {self.last_error()}")

    def score_trace(self, tree, cfgs, seeds, fish=False):
        """remy_b200_score_trace[_fish] -> (sims, senders, use counts, [(leaf[n], values[n, words - 1]) per config])."""
        sink_t = C.CFUNCTYPE(C.c_int, _PTR, C.c_uint32, C.POINTER(C.c_uint64), C.c_uint64, C.c_uint32)
        f = self.lib.remy_b200_score_trace_fish if fish else self.lib.remy_b200_score_trace
        f.restype = C.c_int
        f.argtypes = [C.POINTER(FlatTreeC), _PTR, _PTR, C.c_uint32, C.c_uint32, _PTR, _PTR, _PTR, sink_t, _PTR]
        cfgs = np.ascontiguousarray(cfgs, dtype=net_config_dtype)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        stride = max(1, int(cfgs["num_senders"].max())) if len(cfgs) else 1
        sims = np.zeros(len(cfgs), dtype=sim_result_dtype)
        senders = np.zeros((len(cfgs), stride), dtype=sender_result_dtype)
        counts = np.zeros((1, tree.n_leaves), dtype=np.uint32)
        chunks = [[] for _ in range(len(cfgs))]
        order = []

        def sink(_user, k, rec, n, words):
            order.append(k)
            if n:
                chunks[k].append(np.ctypeslib.as_array(rec, shape=(n * words,)).reshape(n, words).copy())
            return 0

        cb = sink_t(sink)
        rc = f(tree.byref(), _ptr(cfgs), _ptr(seeds), len(cfgs), stride, _ptr(sims), _ptr(senders), _ptr(counts), cb, None)
        if rc != 0:
            raise RuntimeError(f"remy_b200_score_trace failed with {rc}: {self.last_error()}")
        assert order == sorted(order)
        out = []
        for c in chunks:
            a = np.concatenate(c) if c else np.zeros((0, 1), dtype=np.uint64)
            out.append((a[:, 0].astype(np.uint32), np.ascontiguousarray(a[:, 1:]).view(np.float64)))
        return sims, senders, counts, out

    def create_batch(self, tree, cfgs, seeds, cands=None, want_senders=False, want_use_counts=False, fish=False, flows=False):
        return DeviceBatch(self, tree, cfgs, seeds, cands, want_senders, want_use_counts, fish, flows)


class DeviceBatch:
    """Staged scoring: inputs stay resident in HBM between runs."""

    def __init__(self, dev, tree, cfgs, seeds, cands, want_senders, want_use_counts, fish=False, flows=False):
        self.dev = dev
        self.tree = tree
        self.cfgs = np.ascontiguousarray(cfgs, dtype=net_config_dtype)
        self.seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        self.cands = None if cands is None else np.ascontiguousarray(cands, dtype=leaf_override_dtype)
        self.n_cands = 1 if cands is None else len(cands)
        self.n = self.n_cands * len(self.cfgs)
        self.stride = max(1, int(self.cfgs["num_senders"].max()))
        self.want_senders = want_senders
        self.want_use_counts = want_use_counts
        h = _PTR()
        create = (dev.lib.remy_b200_batch_create_fish if fish else dev.lib.remy_b200_batch_create_flows if flows
                  else dev.lib.remy_b200_batch_create)
        rc = create(tree.byref(), _ptr(self.cands), self.n_cands, _ptr(self.cfgs),
                                            _ptr(self.seeds), len(self.cfgs), self.stride,
                                            int(want_senders), int(want_use_counts), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"remy_b200_batch_create failed with {rc}: {dev.last_error()}")
        self.h = h

    def run(self, timed=False):
        ms = C.c_float(0.0)
        rc = self.dev.lib.remy_b200_batch_run(self.h, C.byref(ms) if timed else None)
        if rc != 0:
            raise RuntimeError(f"remy_b200_batch_run failed with {rc}: {self.dev.last_error()}")
        return ms.value

    def launches(self):
        return self.dev.lib.remy_b200_batch_launches(self.h)

    def rerun(self):
        """simulations the last run repeated with larger link rings"""
        return self.dev.lib.remy_b200_batch_rerun(self.h)

    def devices(self):
        return self.dev.lib.remy_b200_batch_devices(self.h)

    def fetch(self):
        sims = np.zeros(self.n, dtype=sim_result_dtype)
        senders = np.zeros((self.n, self.stride), dtype=sender_result_dtype) if self.want_senders else None
        counts = np.zeros((self.n_cands, self.tree.n_leaves), dtype=np.uint32) if self.want_use_counts else None
        rc = self.dev.lib.remy_b200_batch_fetch(self.h, _ptr(sims), _ptr(senders), _ptr(counts))
        if rc != 0:
            raise RuntimeError(f"remy_b200_batch_fetch failed with {rc}: {self.dev.last_error()}")
        return sims, senders, counts

    def destroy(self):
        if self.h:
            self.dev.lib.remy_b200_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

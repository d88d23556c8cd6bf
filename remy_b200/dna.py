"""Minimal proto2 wire reader/writer for the reference's .dna / .cfg files (host logic).

Schemas: reference protobufs/dna.proto (field numbers quoted inline).  This is
host-side glue for tests, tools and bench.py — flattening a RemyBuffers.WhiskerTree
into the POD arrays of include/remy_b200.h and decoding ConfigRange[Unicorn]
files.  The C++ host library (remy_b200/host) has its own reader; tests cross-check
the two and python-protobuf.
"""
import struct

import numpy as np

from . import abi

WILDCARD_UPPER = 163840.0  # src/memory.cc:134-143: missing upper fields match anything


def _varint(buf, i):
    shift = 0
    val = 0
    while True:
        b = buf[i]
        i += 1
        val |= (b & 0x7F) << shift
        if not (b & 0x80):
            return val, i
        shift += 7


def parse_message(buf):
    """-> list of (field_number, wire_type, value); value is int, 8/4 raw bytes, or bytes."""
    out = []
    i = 0
    n = len(buf)
    while i < n:
        key, i = _varint(buf, i)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(buf, i)
        elif wt == 1:
            v = bytes(buf[i:i + 8]); i += 8
        elif wt == 2:
            ln, i = _varint(buf, i)
            v = bytes(buf[i:i + ln]); i += ln
        elif wt == 5:
            v = bytes(buf[i:i + 4]); i += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        out.append((fn, wt, v))
    if i != n:
        raise ValueError("truncated message")
    return out


def _f64(b):
    return struct.unpack("<d", b)[0]


def _zigzag(v):
    return (v >> 1) ^ -(v & 1)


def parse_memory(buf, is_lower):
    """RemyBuffers.Memory (dna.proto:47-54): fields 21..26 doubles; missing -> wildcard."""
    vals = [0.0 if is_lower else WILDCARD_UPPER] * abi.NUM_AXES
    for fn, wt, v in parse_message(buf):
        if 21 <= fn <= 26 and wt == 1:
            vals[fn - 21] = _f64(v)
    return vals


def parse_memory_range(buf):
    """RemyBuffers.MemoryRange (dna.proto:31-45): 11 lower, 12 upper, 13 active_axis."""
    lower = [0.0] * abi.NUM_AXES
    upper = [WILDCARD_UPPER] * abi.NUM_AXES
    axes = []
    for fn, wt, v in parse_message(buf):
        if fn == 11:
            lower = parse_memory(v, True)
        elif fn == 12:
            upper = parse_memory(v, False)
        elif fn == 13:
            if wt == 0:
                axes.append(v)
            else:  # packed
                i = 0
                while i < len(v):
                    a, i = _varint(v, i)
                    axes.append(a)
    if not axes:
        axes = [0, 1, 2, 3]  # src/memoryrange.cc:114-118
    mask = 0
    for a in axes:
        mask |= 1 << a
    return lower, upper, mask


class TreeNode:
    __slots__ = ("lower", "upper", "mask", "children", "leaf")

    def __init__(self):
        self.lower = [0.0] * abi.NUM_AXES
        self.upper = [WILDCARD_UPPER] * abi.NUM_AXES
        self.mask = 0b1111
        self.children = []
        self.leaf = None  # (window_increment, window_multiple, intersend)


def parse_whisker_tree(buf):
    """RemyBuffers.WhiskerTree (dna.proto:3-15) -> (TreeNode, extras dict of raw sub-messages 4,5,6)."""
    node = TreeNode()
    extras = {}
    for fn, wt, v in parse_message(buf):
        if fn == 1:
            node.lower, node.upper, node.mask = parse_memory_range(v)
        elif fn == 2:
            node.children.append(parse_whisker_tree(v)[0])
        elif fn == 3:
            incr, mult, inter = 0, 0.0, 0.0
            dom = None
            for f2, w2, v2 in parse_message(v):
                if f2 == 31:
                    incr = _zigzag(v2)
                elif f2 == 32:
                    mult = _f64(v2)
                elif f2 == 33:
                    inter = _f64(v2)
                elif f2 == 34:
                    dom = parse_memory_range(v2)
            node.leaf = (incr, mult, inter, dom)
        elif fn in (4, 5, 6):
            extras[fn] = v
    return node, extras


def flatten(root):
    """TreeNode -> abi.FlatTree.  Nodes breadth-first (children contiguous, stored order);
    leaves numbered depth-first like the reference's serialisation/str() order."""
    order = [root]
    child_begin = {}
    i = 0
    while i < len(order):
        nd = order[i]
        if nd.children:
            child_begin[id(nd)] = len(order)
            order.extend(nd.children)
        i += 1
    leaf_index = {}
    leaves = []

    def dfs(nd):
        if nd.leaf is not None and not nd.children:
            leaf_index[id(nd)] = len(leaves)
            leaves.append(nd.leaf)
        for c in nd.children:
            dfs(c)

    dfs(root)
    nodes = np.zeros(len(order), dtype=abi.tree_node_dtype)
    for k, nd in enumerate(order):
        nodes[k]["lower"][:] = nd.lower
        nodes[k]["upper"][:] = nd.upper
        nodes[k]["axis_mask"] = nd.mask
        if nd.children:
            nodes[k]["child_begin"] = child_begin[id(nd)]
            nodes[k]["child_count"] = len(nd.children)
        else:
            # a leaf node's lookup domain is the WhiskerTree node's _domain
            # (whiskertree.cc:64); the Whisker's own domain duplicates it.
            nodes[k]["leaf_index"] = leaf_index[id(nd)]
    lv = np.zeros(len(leaves), dtype=abi.leaf_action_dtype)
    for k, (incr, mult, inter, _dom) in enumerate(leaves):
        lv[k] = (mult, inter, incr, 0)
    return abi.FlatTree(nodes, lv)


def load_dna(path):
    with open(path, "rb") as f:
        root, _ = parse_whisker_tree(f.read())
    return flatten(root)


def parse_range(buf):
    """RemyBuffers.Range (dna.proto:89-93): 61 low, 62 high, 63 incr."""
    r = {"low": 0.0, "high": 0.0, "incr": 0.0}
    for fn, wt, v in parse_message(buf):
        if fn == 61:
            r["low"] = _f64(v)
        elif fn == 62:
            r["high"] = _f64(v)
        elif fn == 63:
            r["incr"] = _f64(v)
    return r


_CR_FIELDS = {71: "link_ppt", 72: "rtt", 73: "num_senders", 74: "buffer_size", 75: "mean_off_duration",
              76: "mean_on_duration", 78: "stochastic_loss_rate"}


def parse_config_range(buf):
    """RemyBuffers.ConfigRange (dna.proto:95-104) or ConfigRangeUnicorn (dna.proto:106-119):
    field 77 is a uint32 in the former and a Range in the latter; accept either."""
    out = {}
    for fn, wt, v in parse_message(buf):
        if fn in _CR_FIELDS:
            out[_CR_FIELDS[fn]] = parse_range(v)
        elif fn == 77:
            if wt == 0:
                out["simulation_ticks"] = {"low": float(v), "high": float(v), "incr": 0.0}
            else:
                out["simulation_ticks"] = parse_range(v)
        elif fn == 79:
            out["num_threads"] = v
        elif fn == 80:
            out["cooperative"] = bool(v)
        elif fn == 81:
            out["delay_delta"] = _f64(v)
        elif fn == 82:
            out["iterations"] = v
    return out


def cfg_to_netconfig(cr):
    """Single-point .cfg -> NetConfig dict for the Rat/TimeSwitched path:
    every *.low; mean_on <= 0 (a ByteSwitchedSender flow length, src/sendergang.cc:128-134)
    maps to sender-runner's default 5000 ms (src/sender-runner.cc:53)."""
    on = cr["mean_on_duration"]["low"]
    buf = cr["buffer_size"]["low"]
    return dict(link_ppt=cr["link_ppt"]["low"], delay=cr["rtt"]["low"],
                num_senders=int(cr["num_senders"]["low"]),
                buffer_size=int(min(buf, 4294967295.0)),
                mean_off_duration=cr["mean_off_duration"]["low"],
                mean_on_duration=on if on > 0 else 5000.0,
                stochastic_loss_rate=cr.get("stochastic_loss_rate", {"low": 0.0})["low"],
                simulation_ticks=cr["simulation_ticks"]["low"])

"""The C-ABI library loads and exports every symbol include/remy_b200.h declares; struct
layouts used by the Python mirror match the header (compiled with gcc); without a GPU the
entry points fail loudly instead of computing anything."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from helpers import ROOT
from remy_b200 import abi

HEADER = os.path.join(ROOT, "include", "remy_b200.h")
LIB = os.path.join(ROOT, "remy_b200", "csrc", "libremy_b200.so")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(remy_b200_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib(built):
    built.build_cuda()
    return C.CDLL(LIB)


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    lib.remy_b200_abi_version.restype = C.c_uint32
    assert lib.remy_b200_abi_version() == 2


def test_struct_layouts_match_header():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "remy_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(RemyNetConfig), sizeof(RemyTreeNode), sizeof(RemyLeafAction),
         sizeof(RemyLeafOverride), sizeof(RemySimResult), sizeof(RemySenderResult), sizeof(RemyFlatTree));
  printf("%zu %zu %zu %zu\n", offsetof(RemyNetConfig, num_senders), offsetof(RemyTreeNode, axis_mask),
         offsetof(RemySimResult, prng_state), offsetof(RemyLeafOverride, leaf_index));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")], text=True).split()
    sizes = [abi.net_config_dtype.itemsize, abi.tree_node_dtype.itemsize, abi.leaf_action_dtype.itemsize,
             abi.leaf_override_dtype.itemsize, abi.sim_result_dtype.itemsize, abi.sender_result_dtype.itemsize,
             C.sizeof(abi.FlatTreeC)]
    assert [int(x) for x in out[:7]] == sizes
    offs = [abi.net_config_dtype.fields["num_senders"][1], abi.tree_node_dtype.fields["axis_mask"][1],
            abi.sim_result_dtype.fields["prng_state"][1], abi.leaf_override_dtype.fields["leaf_index"][1]]
    assert [int(x) for x in out[7:]] == offs


def test_no_gpu_means_loud_failure(lib):
    """There is no CPU fallback in the product: without a device init() returns REMY_ENODEV
    and create() refuses to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib.remy_b200_init.argtypes = [C.c_int]
    lib.remy_b200_last_error.restype = C.c_char_p
    assert lib.remy_b200_init(0) == -2
    assert b"no CUDA device" in lib.remy_b200_last_error()
    assert lib.remy_b200_init_devices(None, 0) == -2   # "all devices" of a machine without any
    assert lib.remy_b200_device_count() == 0
    with pytest.raises(RuntimeError):
        abi.Device(0)
    tree = abi.default_tree()
    h = C.c_void_p()
    rc = lib.remy_b200_batch_create(tree.byref(), None, 1, None, None, 0, 1, 0, 0, C.byref(h))
    assert rc != 0 and not h.value


def test_product_does_not_link_the_oracle(lib):
    """The shipped library must not depend on anything under oracle/ or tests/emu."""
    out = subprocess.check_output(["ldd", LIB], text=True)
    assert "oracle" not in out and "emu" not in out
    syms = subprocess.check_output(["nm", "-D", "--defined-only", LIB], text=True)
    assert "remy_oracle" not in syms and "remy_emu" not in syms and "remy_ref" not in syms

"""The thread-per-sub-tree BLAS builder (voidin_b200/csrc/t4_seq.cuh, what k_t4 runs per thread) compiled for the
host and checked bit-for-bit against the CPU oracle on small meshes: nodes, numbering and primitive order.
CPU only; the same code path is covered on the GPU by tests/test_gpu_parity.py through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from voidin_b200 import scenes as S

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "t4_host.cpp")
HDR = os.path.join(HERE, "..", "voidin_b200", "csrc", "t4_seq.cuh")
LIB = os.path.join(HERE, "native", "libt4_host.so")


@pytest.fixture(scope="module")
def t4():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared",
                               "-o", LIB, SRC])
    return C.CDLL(LIB)


def run(t4, v, idx, stride=1, column=0, cap=32):
    v = np.ascontiguousarray(v, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1)
    n = idx.size // 3
    nodes = np.zeros(2 * n, dtype=O.BVH_NODE)
    order = np.zeros(n, dtype=np.uint32)
    m = C.c_uint32(0)
    rc = t4.t4_host_build_cap(v.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_uint32(stride),
                              C.c_uint32(column), nodes.ctypes.data_as(C.c_void_p), C.byref(m), order.ctypes.data_as(C.c_void_p),
                              C.c_uint32(cap))
    return rc, nodes[: m.value], order


def check(t4, v, idx, **kw):
    rc, nodes, order = run(t4, v, idx, **kw)
    orc, onodes, _, oorder, _ = O.blas_build(v, idx)
    if orc == O.EDEGENERATE:
        assert rc == -2
        return
    assert orc == 0 and rc == 0
    assert len(nodes) == len(onodes)
    assert nodes.tobytes() == onodes.tobytes()
    assert (order == oorder).all()


def test_soups_all_sizes(t4):
    for n in range(1, 33):
        for seed in range(12):
            v, i = S.soup(n, 1000 * n + seed, 0.2)
            check(t4, v, i)


def test_shipped_capacity_eight(t4):
    """k_t4 is instantiated with a capacity of 8 primitives in libbvh_cuda.so (768 threads per block): the same
    instantiation, with the block's [word][thread] layout, on every size it takes."""
    rng = np.random.default_rng(8)
    for n in range(1, 9):
        for seed in range(40):
            v, i = S.soup(n, 50_000 + 100 * n + seed, float(rng.choice([0.5, 0.2, 0.02])))
            check(t4, v, i, cap=8)
            check(t4, v, i, cap=8, stride=768, column=int(rng.integers(0, 768)))


def test_layout_stride_and_column(t4):
    v, i = S.soup(32, 77, 0.3)
    check(t4, v, i, stride=5, column=3)
    check(t4, v, i, stride=160, column=159)


def test_shared_vertices_and_flat_axes(t4):
    # grids: shared vertices, one zero-extent axis (every candidate on it is NaN), many equal centroids per plane
    for nx, ny in ((4, 4), (2, 8), (1, 16), (16, 1), (3, 5)):
        v, i = S.grid_mesh(nx, ny)
        if i.size // 3 <= 32:
            check(t4, v, i)


def test_random_quantised_and_signed_zeros(t4):
    rng = np.random.default_rng(2024)
    for trial in range(400):
        n = int(rng.integers(1, 33))
        # coordinates on a coarse lattice: ties between centroids and planes, duplicate points, +-0.0 faces
        v = rng.integers(-2, 3, size=(3 * n, 3)).astype(np.float32) * np.float32(0.5)
        z = rng.random(v.shape) < 0.15
        v[z & (v == 0)] = np.float32(-0.0)
        i = np.arange(3 * n, dtype=np.uint32)
        check(t4, v, i)


def test_huge_and_tiny_scales(t4):
    rng = np.random.default_rng(5)
    for scale in (1e-30, 1e-12, 1e12, 3e18):
        for n in (4, 9, 17, 32):
            v = (rng.random((3 * n, 3)) - 0.5).astype(np.float32) * np.float32(scale)
            check(t4, v, np.arange(3 * n, dtype=np.uint32))

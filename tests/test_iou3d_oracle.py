"""CPU tests: pin the IoU3D oracle (oracle/iou3d_oracle.c) against closed forms, the independent
fp64 half-space oracle, the reference-generated wrapper fixture, and check the product's per-lane
geometry (box3d_geom.cuh compiled for the host) is bit-identical to the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import boxgen
from oracle import iou3d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests/golden/iou3d_wrapper_golden.npz"))


def test_known_answers():
    b = GOLD["ka_boxes"]
    vol, iou, nf, _ = iou3d.iou_box3d(b[:1], b, with_counts=True)
    np.testing.assert_allclose(iou[0], GOLD["ka_expect"], atol=2e-6)
    assert nf[0, 5] == 0 and nf[0, 6] == 0 or iou[0, 6] == 0.0   # disjoint: zero faces
    assert nf[0, 0] == 12


@pytest.mark.parametrize("L,seed", [(1.0, 0), (3.0, 1)])
def test_oracle_vs_fp64_halfspace(L, seed):
    a = boxgen.random_boxes(12, L, seed)
    b = boxgen.random_boxes(12, L, seed + 9)
    vol, iou = iou3d.iou_box3d(a, b)
    for i in range(12):
        for j in range(12):
            v64, u64 = iou3d.iou_halfspace_fp64(a[i], b[j])
            assert abs(u64 - iou[i, j]) < 2e-5, (i, j, u64, iou[i, j])
            assert abs(v64 - vol[i, j]) < 2e-5 * max(1.0, v64)


def test_rigid_invariance_and_range():
    rng = np.random.default_rng(5)
    a = boxgen.random_boxes(40, 1.5, 2)
    b = boxgen.random_boxes(40, 1.5, 3)
    _, iou = iou3d.iou_box3d(a, b)
    R = boxgen.random_rotations(1, rng)[0]
    t = rng.uniform(-3, 3, 3)
    a2 = (a.astype(np.float64) @ R.T + t).astype(np.float32)
    b2 = (b.astype(np.float64) @ R.T + t).astype(np.float32)
    _, iou2 = iou3d.iou_box3d(a2, b2)
    assert np.abs(iou - iou2).max() < 5e-5
    assert iou.min() >= 0 and iou.max() <= 1 + 1e-6
    _, self_iou = iou3d.iou_box3d_paired(a, a)[:2]
    np.testing.assert_allclose(self_iou, 1.0, atol=1e-5)


@pytest.mark.parametrize("name", ["dense", "sparse", "mid"])
def test_wrapper_matches_reference_fixture(name):
    dt, gt = GOLD[f"{name}_dt"], GOLD[f"{name}_gt"]
    cop, nz = iou3d.check_boxes(dt)
    assert np.array_equal(cop, GOLD[f"{name}_coplanar_ok"])
    assert np.array_equal(nz, GOLD[f"{name}_nonzero_ok"])
    iou, nbad = iou3d.box3d_overlap(dt, gt)
    assert np.array_equal(iou, GOLD[f"{name}_iou"])
    assert nbad[0] == (~cop).sum() and nbad[1] == (~nz).sum()


def test_threaded_equals_serial():
    a = boxgen.random_boxes(60, 1.0, 0)
    b = boxgen.random_boxes(50, 1.0, 1)
    r1 = iou3d.iou_box3d(a, b, threads=1, with_counts=True)
    r4 = iou3d.iou_box3d(a, b, threads=4, with_counts=True)
    for x, y in zip(r1, r4):
        assert np.array_equal(x, y)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("h") / "libharness.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", out,
                           os.path.join(ROOT, "tests/csrc/geom_host_harness.cpp")])
    return ctypes.CDLL(out)


@pytest.mark.parametrize("L,seed", [(1.0, 0), (3.0, 4), (10.0, 8)])
def test_product_geometry_on_host_is_bit_exact(harness, L, seed):
    a = boxgen.random_boxes(80, L, seed)
    b = boxgen.random_boxes(70, L, seed + 31)
    N, M = len(a), len(b)
    v = np.zeros((N, M), np.float32); u = np.zeros((N, M), np.float32); nf = np.zeros((N, M), np.int32)
    P = lambda x, t=ctypes.c_float: x.ctypes.data_as(ctypes.POINTER(t))
    harness.harness_iou(P(a), N, P(b), M, P(v), P(u), P(nf, ctypes.c_int32))
    vo, uo, nfo, _ = iou3d.iou_box3d(a, b, threads=4, with_counts=True)
    assert np.array_equal(nf, nfo)
    assert np.array_equal(v, vo) and np.array_equal(u, uo)


def test_product_row_checks_on_host(harness):
    for name in ["dense", "sparse", "mid"]:
        dt = np.ascontiguousarray(GOLD[f"{name}_dt"])
        fl = np.zeros(len(dt), np.int32)
        harness.harness_check(dt.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), len(dt),
                              ctypes.c_float(1e-4), ctypes.c_float(1e-8),
                              fl.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        assert np.array_equal((fl & 1).astype(bool), GOLD[f"{name}_coplanar_ok"])
        assert np.array_equal((fl & 2).astype(bool), GOLD[f"{name}_nonzero_ok"])

"""Host-side cluster builder of the shadow sweep (csrc/ore_clusters.h), compiled with g++ and checked on the CPU.

The beam kernel culls whole 32-sphere clusters by their bounding sphere, so the one property everything rests on is
containment: every member ball (centre, R') lies inside its cluster's ball - or the cluster is 'always open' (+inf).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ray-tracer-engine_b200", "csrc")


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("clu") / "libcluster_probe.so")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC,
                           os.path.join(ROOT, "tests", "cluster_probe.cpp"), "-o", out], env=env)
    lib = C.CDLL(out)
    fp = C.POINTER(C.c_float)
    lib.ore_probe_build_clusters.argtypes = [fp, fp, C.c_int, fp, fp, fp, C.c_int, C.c_int]
    lib.ore_probe_build_clusters.restype = C.c_int

    def build(ex, sh):
        n = len(ex)
        n_clu = (n + 31) // 32
        ss = np.zeros((max(1, n_clu) * 32, 4), dtype=np.float32)
        xs = np.zeros_like(ss)
        cl = np.zeros(((max(1, n_clu) + 3) // 4 * 4, 4), dtype=np.float32)
        exc = np.ascontiguousarray(ex, dtype=np.float32).reshape(-1, 4) if n else np.zeros((1, 4), dtype=np.float32)
        shc = np.ascontiguousarray(sh, dtype=np.float32).reshape(-1, 4) if n else np.zeros((1, 4), dtype=np.float32)
        got = lib.ore_probe_build_clusters(exc.ctypes.data_as(fp), shc.ctypes.data_as(fp), n, ss.ctypes.data_as(fp),
                                           xs.ctypes.data_as(fp), cl.ctypes.data_as(fp), len(ss), len(cl))
        assert got == n_clu
        return ss, xs, cl, n_clu

    return build


def make(rng, n, extent=40.0):
    pos = rng.uniform(0, extent, size=(n, 3)).astype(np.float32)
    member = (rng.uniform(0, 1.3, size=n) ** 2).astype(np.float32)
    ex = np.concatenate([pos, member[:, None]], axis=1).astype(np.float32)
    sh = ex.copy()
    sh[:, 3] = np.nextafter(member * np.float32(1.000004), np.float32(np.inf))   # R' >= effective radius
    return ex, sh


def check(ex, sh, ss, xs, cl, n_clu):
    n = len(ex)
    key = lambda a: sorted(map(tuple, np.nan_to_num(a[:n].astype(np.float64), nan=-7.25e-5).tolist()))
    assert key(ss) == key(sh), "sorted shadow records must be a permutation of the input"
    # the exact records follow the same permutation: same centre bits in both sorted arrays
    assert np.array_equal(ss[:n, :3].view(np.uint32), xs[:n, :3].view(np.uint32))
    for j in range(n_clu):
        mem = ss[j * 32:min(j * 32 + 32, n)].astype(np.float64)
        tame = np.all(np.isfinite(mem) & (np.abs(mem) < 1e15))
        c, r = cl[j, :3].astype(np.float64), float(cl[j, 3])
        if not tame:
            assert r == np.inf, (j, mem, r)
            continue
        assert np.isfinite(r)
        d = np.sqrt(((mem[:, :3] - c) ** 2).sum(axis=1)) + mem[:, 3]
        assert np.all(d <= r), (j, float(d.max()), r)


@pytest.mark.parametrize("n", [0, 1, 5, 31, 32, 33, 64, 150, 1024, 5000])
def test_every_member_ball_lies_inside_its_cluster_ball(probe, n):
    rng = np.random.default_rng(n)
    ex, sh = make(rng, n)
    ss, xs, cl, n_clu = probe(ex, sh)
    if n:
        check(ex, sh, ss, xs, cl, n_clu)
    assert n_clu == (n + 31) // 32
    assert np.all(cl[n_clu:] == 0)


def test_clusters_are_spatially_compact(probe):
    """Morton order: the mean cluster radius of 4096 uniformly placed spheres must be far below the scene extent"""
    rng = np.random.default_rng(9)
    ex, sh = make(rng, 4096, extent=100.0)
    ss, xs, cl, n_clu = probe(ex, sh)
    assert np.mean(cl[:n_clu, 3]) < 30.0


def test_untame_members_open_their_cluster(probe):
    rng = np.random.default_rng(3)
    ex, sh = make(rng, 200)
    weird = [(1e17, 0, 0, 1.0), (0, -3e19, 0, 1e10), (np.inf, 0, 0, 1.0), (-np.inf, np.inf, 0, 2.0), (np.nan, 1, 1, 0.5),
             (1, 1, 1, np.nan), (5e15, 5e15, 5e15, 1e3), (3, 3, 3, 2e15)]
    for k, w in enumerate(weird):
        ex[k * 11 + 2] = np.array(w, dtype=np.float32)
        sh[k * 11 + 2] = np.array(w, dtype=np.float32)
    ss, xs, cl, n_clu = probe(ex, sh)
    check(ex, sh, ss, xs, cl, n_clu)
    assert np.count_nonzero(np.isinf(cl[:n_clu, 3])) >= 1
    assert np.count_nonzero(np.isfinite(cl[:n_clu, 3])) >= 1, "tame clusters must stay bounded"


def test_degenerate_layouts(probe):
    # all centres identical, all on a line, zero radii
    for ex in (np.tile(np.array([[1, 2, 3, 0.25]], dtype=np.float32), (70, 1)),
               np.stack([np.linspace(0, 9, 70), np.zeros(70), np.zeros(70), np.zeros(70)], axis=1).astype(np.float32)):
        ss, xs, cl, n_clu = probe(ex, ex.copy())
        check(ex, ex.copy(), ss, xs, cl, n_clu)


# ---- round 2: leaves of 8 spheres and super-clusters of 32 leaves over the same Morton order -----------------------

@pytest.fixture(scope="module")
def hierarchy(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("clu2") / "libcluster_probe2.so")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC,
                           os.path.join(ROOT, "tests", "cluster_probe.cpp"), "-o", out], env=env)
    lib = C.CDLL(out)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    lib.ore_probe_build_hierarchy.argtypes = [fp, fp, C.c_int, fp, ip, fp, fp, C.c_int, C.c_int, C.c_int]
    lib.ore_probe_build_hierarchy.restype = C.c_int

    def build(ex, sh):
        n = len(ex)
        n_sort = max(1, (n + 31) // 32) * 32
        n_leaf = (n + 7) // 8
        ss = np.zeros((n_sort, 4), dtype=np.float32)
        order = np.zeros(n_sort, dtype=np.int32)
        lv = np.zeros((max(32, (n_leaf + 31) // 32 * 32), 4), dtype=np.float32)
        sp = np.zeros((max(4, ((n_leaf + 31) // 32 + 3) // 4 * 4), 4), dtype=np.float32)
        exc = np.ascontiguousarray(ex, dtype=np.float32).reshape(-1, 4) if n else np.zeros((1, 4), dtype=np.float32)
        shc = np.ascontiguousarray(sh, dtype=np.float32).reshape(-1, 4) if n else np.zeros((1, 4), dtype=np.float32)
        got = lib.ore_probe_build_hierarchy(exc.ctypes.data_as(fp), shc.ctypes.data_as(fp), n, ss.ctypes.data_as(fp),
                                            order.ctypes.data_as(ip), lv.ctypes.data_as(fp), sp.ctypes.data_as(fp),
                                            len(ss), len(lv), len(sp))
        assert got == n_leaf
        return ss, order, lv, sp, n_leaf

    return build


def contained(members, ball):
    mem = members.astype(np.float64)
    tame = np.all(np.isfinite(mem) & (np.abs(mem) < 1e15))
    c, r = ball[:3].astype(np.float64), float(ball[3])
    if not tame:
        return r == np.inf
    d = np.sqrt(((mem[:, :3] - c) ** 2).sum(axis=1)) + np.abs(mem[:, 3])
    return bool(np.isfinite(r) and np.all(d <= r))


@pytest.mark.parametrize("n", [1, 7, 8, 9, 255, 256, 257, 1024, 5000, 16384])
def test_leaves_and_super_clusters_contain_their_members(hierarchy, n):
    rng = np.random.default_rng(n + 1)
    ex, sh = make(rng, n)
    if n >= 255:   # a few untame members
        for k in (3, n // 2, n - 2):
            sh[k, 0] = ex[k, 0] = np.float32(1e17 if k % 2 else np.inf)
    ss, order, lv, sp, n_leaf = hierarchy(ex, sh)
    # the permutation: sorted position p holds original sphere order[p], every sphere exactly once
    assert sorted(order[:n].tolist()) == list(range(n))
    assert np.array_equal(ss[:n].view(np.uint32), sh[order[:n]].view(np.uint32))
    for j in range(n_leaf):
        assert contained(ss[8 * j:min(8 * j + 8, n)], lv[j]), ("leaf", j)
    n_sup = (n_leaf + 31) // 32
    for k in range(n_sup):
        assert contained(ss[256 * k:min(256 * k + 256, n)], sp[k]), ("super", k)
    assert np.all(lv[n_leaf:, 3] == -1) and np.all(sp[n_sup:, 3] == -1)   # padding: never touched


def test_leaves_are_much_tighter_than_the_round1_clusters(hierarchy, probe):
    rng = np.random.default_rng(5)
    ex, sh = make(rng, 1024, extent=25.0)
    ss, order, lv, sp, n_leaf = hierarchy(ex, sh)
    _, _, cl, n_clu = probe(ex, sh)
    assert np.mean(lv[:n_leaf, 3]) < 0.72 * np.mean(cl[:n_clu, 3])

"""The C++20 façade (include/axiom/collision/collision_world.hpp) compiles against the C ABI and
behaves like the documented Broadphase/Narrowphase call shape (reference: CLAUDE.md:162-178)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "axiom-physics-engine_b200")
EXE = os.path.join(PKG, "facade_smoke")


def _build():
    subprocess.check_call(["make", "-C", PKG, "facade_smoke", "bench_step", "slab_host"], stdout=subprocess.DEVNULL)


def test_facade_compiles_against_the_engine_headers():
    """Inside the Axiom tree the façade must use the engine's own Result / Transform / AABB
    (AXIOM_COLLISION_HAS_ENGINE_TYPES), not its stand-ins.  Needs the reference checkout; skipped on the GPU box."""
    ref = "/root/reference/include"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present")
    src = os.path.join(ROOT, "tests", "cpp", "facade_smoke.cpp")
    r = subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-DAXIOM_EXPECT_ENGINE_TYPES", "-I" + ref,
                        "-I" + os.path.join(ROOT, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def test_facade_links_and_refuses_without_device():
    import torch
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        # no device: create() must fail with the GPU-init code, never fall back to a CPU path
        assert r.returncode == 77, r.stdout + r.stderr
        assert "500" in r.stdout


@pytest.mark.gpu
def test_facade_c0_counts():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout.split()
    assert out[:2] == ["2875", "1381"]   # C0 fixture: candidate pairs, contacts
    assert int(out[2]) >= 1381           # contact points over all manifolds


@pytest.mark.gpu
def test_cpp_host_at_c1_size():
    """The C++20 host (tests/cpp/bench_step.cpp) through the façade at config C1: 100 k bodies, the same
    counts as the ctypes path."""
    _build()
    exe = os.path.join(os.path.dirname(EXE), "bench_step")
    r = subprocess.run([exe, "100000", "46.4", "2", "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    n, pairs, contacts, dev_ms, wall_ms = r.stdout.split()
    import axcd
    w = axcd.CollisionWorld.for_scene(axcd.config_scene("C1"))
    st = w.step()
    assert (int(n), int(pairs), int(contacts)) == (100000, st.numPairs, st.numContacts)
    assert 0 < float(dev_ms) < float(wall_ms)
    w.close()


@pytest.mark.gpu
def test_cpp_slab_host_two_ranks_over_nccl():
    """tests/cpp/slab_host.cpp: a C++20 host, two ranks as host threads on two GPUs, the ghost exchange over NCCL
    inside libaxcd.so.  The program itself checks that the union of the ranks' pair and contact sets equals the
    single-GPU run byte for byte (exit 20 / 21 otherwise)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _build()
    exe = os.path.join(os.path.dirname(EXE), "slab_host")
    r = subprocess.run([exe, "200000", "58.5", "7", "2", "5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    ranks, n, pairs, contacts, ghosts, wall_ms, graph = r.stdout.strip().splitlines()[-1].split()   # NCCL may print its banner first
    assert (int(ranks), int(n)) == (2, 200000) and int(pairs) > 500000 and int(ghosts) > 0 and int(graph) == 1

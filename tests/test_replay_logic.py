"""CPU-only: the native replay loop (tpp_mlir_b200/csrc/harness/replay.cpp - the stand-in for tpp-run's JIT-compiled
hot loop) against a recording stub of the C-ABI (tests/stubs/xsmm_abi_stub.cpp). Checks WHAT it asks the runtime to do:
the invoke stream of any run equals the plain loop's, whatever mix of rotation / partial / single-step graphs replays
it; the grouped end-to-end pipeline uploads before it launches, downloads after, and waits before it reuses buffers."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
i64, p = ctypes.c_int64, ctypes.c_void_p


class Set(ctypes.Structure):
    _fields_ = [("acts", p * 9), ("weights", p * 8), ("biases", p * 8)]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("replay_stub") / "libreplay_stub.so"
    # -Bsymbolic: replay.cpp's calls bind to the stub inside this library even when the real ABI library has been
    # loaded into the process (RTLD_GLOBAL) by another test
    cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tpp_mlir_b200", "csrc", "harness", "replay.cpp"),
           os.path.join(ROOT, "tests", "stubs", "xsmm_abi_stub.cpp"), "-o", str(out)]
    subprocess.run(cmd, check=True)
    lib = ctypes.CDLL(str(out))
    lib.tpp_replay_mlp.argtypes = [i64, i64, p, p, i64, i64, i64, i64, p, i64, i64, i64, i64]
    lib.tpp_replay_mlp.restype = None
    lib.tpp_replay_mlp_graph.argtypes = [i64, i64, p, p, i64, i64, i64, i64, p, i64, p, i64, i64, i64, i64]
    lib.tpp_replay_mlp_graph.restype = i64
    lib.tpp_replay_mlp_e2e_pipelined.argtypes = [i64, i64, p, p, i64, i64, i64, i64, p, i64, i64, i64, i64, i64, p, p]
    lib.tpp_replay_mlp_e2e_pipelined.restype = i64
    lib.stub_log_size.restype = i64
    lib.stub_log_get.argtypes = [i64, p]
    lib.stub_num_graphs.restype = i64
    lib.stub_set_wait_ring.argtypes = [i64]
    return lib


def make_sets(n, layers=3, contiguous_groups=0):
    """n operand sets with fake, distinct addresses; contiguous_groups = g: the input / output buffers of every g
    consecutive sets are consecutive 512 KiB slices (one host block per group)."""
    sets = (Set * n)()
    nbytes = 256 * 1024 * 2
    for s in range(n):
        base = 0x10000000 * (s + 1)
        for l in range(layers + 1):
            sets[s].acts[l] = base + 0x1000000 * l
        if contiguous_groups:
            g, j = divmod(s, contiguous_groups)
            sets[s].acts[0] = 0x7000000000 + g * 0x100000000 + j * nbytes
            sets[s].acts[layers] = 0x7800000000 + g * 0x100000000 + j * nbytes
        for l in range(layers):
            sets[s].weights[l] = base + 0x8000000 + 0x200000 * l
            sets[s].biases[l] = base + 0xC000000 + 0x1000 * l
    return sets


def log(lib):
    out, buf = [], (i64 * 6)()
    for i in range(lib.stub_log_size()):
        lib.stub_log_get(i, buf)
        out.append(tuple(buf))
    return out


HANDLES = (i64 * 3)(11, 22, 33)
SIZES = (i64 * 4)(1024, 1024, 1024, 1024)


@pytest.mark.parametrize("num_sets,first,steps", [(17, 0, 17), (17, 0, 40), (17, 5, 40), (148, 0, 2000), (148, 0, 20),
                                                   (148, 140, 170), (4, 3, 1), (4, 0, 3), (1, 0, 7)])
def test_graph_replay_issues_the_same_invoke_stream_as_the_plain_loop(lib, num_sets, first, steps):
    sets = make_sets(num_sets)
    lib.stub_reset()
    lib.tpp_replay_mlp(2, 3, HANDLES, SIZES, 256, 256, 1024, 1024, sets, num_sets, first, steps, 1)
    want = [e for e in log(lib) if e[0] == 1]
    assert len(want) == 3 * steps
    lib.stub_reset()
    graphs = (i64 * (num_sets + 2))()
    assert lib.tpp_replay_mlp_graph(2, 3, HANDLES, SIZES, 256, 256, 1024, 1024, sets, num_sets, graphs, first, steps, 1, 1) == 0
    events = log(lib)
    assert [e for e in events if e[0] == 1] == want, "graph replay changed the invoke stream"
    launches = [e for e in events if e[0] == 8]
    # a run is at most: one partial head, full rotations, one partial tail (single steps only when one step is left)
    head = min(steps, (num_sets - first % num_sets) % num_sets)
    full, tail = divmod(steps - head, num_sets)
    assert len(launches) == (1 if head else 0) + full + (1 if tail else 0)
    # a second identical run captures nothing new
    n_graphs = lib.stub_num_graphs()
    assert lib.tpp_replay_mlp_graph(2, 3, HANDLES, SIZES, 256, 256, 1024, 1024, sets, num_sets, graphs, first, steps, 1, 1) == 0
    assert lib.stub_num_graphs() == n_graphs


def test_partial_graphs_belong_to_one_loop_object(lib):
    """Two loops over different buffers that happen to use the same arrays one after the other (recycled addresses):
    the second one must capture its own partial graphs, not replay the first one's."""
    num_sets, steps = 8, 5
    sets = make_sets(num_sets)
    graphs = (i64 * (num_sets + 2))()
    lib.stub_reset()
    assert lib.tpp_replay_mlp_graph(2, 3, HANDLES, SIZES, 256, 256, 1024, 1024, sets, num_sets, graphs, 0, steps, 1, 1) == 0
    first_run = [e for e in log(lib) if e[0] == 1]
    # "new loop object": same arrays, different buffers, fresh graph slots
    for s in range(num_sets):
        for l in range(4):
            sets[s].acts[l] += 0x40
    for i in range(num_sets + 2):
        graphs[i] = 0
    n_before = lib.stub_log_size()
    assert lib.tpp_replay_mlp_graph(2, 3, HANDLES, SIZES, 256, 256, 1024, 1024, sets, num_sets, graphs, 0, steps, 1, 1) == 0
    second_run = [e for e in log(lib)[n_before:] if e[0] == 1]
    assert len(second_run) == len(first_run) == 3 * steps
    assert all(b[2] == a[2] + 0x40 for a, b in zip(first_run, second_run)), "stale graph replayed"


def test_grouped_e2e_pipeline_orders_copies_launches_and_waits(lib):
    gsz, ngrp, layers = 74, 3, 3
    depth = gsz * ngrp
    sets = make_sets(depth, contiguous_groups=gsz)
    graphs, streams = (i64 * (depth + 1))(), (p * depth)()
    lib.stub_reset()
    steps = 8 * depth
    ran = lib.tpp_replay_mlp_e2e_pipelined(2, layers, HANDLES, SIZES, 256, 256, 1024, 1024, sets, depth, steps, 1, 2,
                                           2 + gsz, graphs, streams)
    assert ran == steps
    events = log(lib)
    nbytes = 256 * 1024 * 2
    per_group = {}
    pos = 0
    for it in range(steps // gsz):
        grp = it % ngrp
        g_in, g_out = sets[grp * gsz].acts[0], sets[grp * gsz].acts[layers]
        if it >= ngrp:   # the group's previous output must have reached the host before its buffers are reused
            assert events[pos] == (4, g_out, 0, 0, 0, 0)
            pos += 1
        assert events[pos] == (2, g_in, gsz * nbytes, 0, 0, 0), "one upload for the whole group"
        assert events[pos + 1][0] == 8
        body = events[pos + 2: pos + 2 + 3 * gsz]
        assert all(e[0] == 1 for e in body)
        # the captured body is the group's gsz forward passes, in order, each on its own buffers
        for j in range(gsz):
            a = sets[grp * gsz + j].acts
            assert [e[2] for e in body[3 * j: 3 * j + 3]] == [a[0], a[1], a[2]]
            assert [e[4] for e in body[3 * j: 3 * j + 3]] == [a[1], a[2], a[3]]
        per_group.setdefault(grp, events[pos + 1][1])
        assert events[pos + 1][1] == per_group[grp], "a group always replays its own graph"
        pos += 2 + 3 * gsz
        assert events[pos] == (3, g_out, gsz * nbytes, 0, 0, 0), "one download for the whole group"
        pos += 1
    assert events[pos][0] == 5 and pos + 1 == len(events), "final drain"
    assert lib.stub_num_graphs() == ngrp


def test_grouped_e2e_drains_when_a_download_is_no_longer_on_record(lib):
    """xsmm_cuda_wait_host only remembers the most recent downloads; when it answers -1 the loop must drain the stream
    (xsmm_cuda_stream_sync) before it reuses the slot - never reuse the buffer blindly (replay.cpp:168)."""
    gsz, ngrp = 12, 2
    sets = make_sets(gsz * ngrp)   # scattered buffers: one download per step
    graphs, streams = (i64 * (gsz * ngrp + 1))(), (p * (gsz * ngrp))()
    lib.stub_reset()
    lib.stub_set_wait_ring(16)     # fewer than the 24 downloads in flight
    assert lib.tpp_replay_mlp_e2e_pipelined(2, 3, HANDLES, SIZES, 256, 256, 1024, 1024, sets, gsz * ngrp, 4 * gsz, 1, 2,
                                            2 + gsz, graphs, streams) == 4 * gsz
    ev = log(lib)
    waits = [i for i, e in enumerate(ev) if e[0] == 4]
    assert waits, "slots are reused, so there must be waits"
    forgotten = 0
    for i in waits:
        host = ev[i][1]
        earlier = [e[1] for e in ev[:i] if e[0] == 3]
        if host not in earlier[-16:]:
            forgotten += 1
            assert ev[i + 1][0] == 5, "a forgotten download must be followed by a stream sync"
    assert forgotten > 0

"""The C++ host mirror of the reference's API (self-play-ai_b200/host/selfplay_b200.hpp): compiles against the C ABI
everywhere; on a B200 it reproduces the survey's greedy DetEval game (main.rs:106-114 move rule)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT = json.load(open(os.path.join(ROOT, "tests", "golden", "survey_kats.json")))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host") / "host_mirror_test")
    lib = os.path.join(ROOT, "self-play-ai_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-o", out, os.path.join(ROOT, "tests", "host_mirror_test.cc"),
                           "-L" + lib, "-lselfplay_b200", "-Wl,-rpath," + lib])
    return out


def test_host_mirror_compiles_and_fails_loudly_without_gpu(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe, "10"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_host_mirror_plays_the_survey_greedy_game(exe):
    r = subprocess.run([exe, "800"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert [int(x) for x in lines[0].split()[1:]] == KAT["c4_det_greedy_800"]["actions"]
    assert lines[1] == "status 2" and lines[2] == "error -1"
    arenas = [int(l.split()[-1]) for l in r.stderr.strip().splitlines()]
    assert arenas == KAT["c4_det_greedy_800"]["arena_sizes"]


@pytest.mark.gpu
def test_host_mirror_chess_perft_and_greedy_plies(exe):
    """ChessMcts of the C++ mirror: device perft(3) of the start position, then four greedy DetEval plies with use_subtree
    against the oracle's search."""
    from oracle import pychess as P
    r = subprocess.run([exe, "100", "chess"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "perft3 8902 moves 20"
    got = [int(x) for x in lines[1].split()[1:]]
    f = P.Forest(1)
    f.reset(0, P.Game())
    want, arenas = [], []
    for _ in range(4):
        f.search(100, 1)
        mvs, cnt, ids = f.root_children(0)
        best = max(range(len(cnt)), key=lambda i: (cnt[i], i))
        want.append(mvs[best])
        arenas.append(f.arena_len(0))
        f.use_subtree(0, ids[best])
    assert got == want and lines[2] == "error -1"
    assert [int(l.split()[-1]) for l in r.stderr.strip().splitlines()] == arenas

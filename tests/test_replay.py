"""the two callers of the harmonic path -- the nav_core plugin's makePlan and
the anytime node's tick + services -- replayed without ROS by tests/native/replay_callers.cpp, a C++
program written against the library's public headers and linked with the drop-in libepic.so.

tests/golden/replay_callers.json is the output of the SAME source compiled against the reference's own
headers and linked with the untouched reference CPU sources (tools/make_replay_golden.py); every line a
client would observe (iteration counts, delta bits, hashes of u / locked / raw paths / pose lists) must be
identical for the library's CPU exports and, on a B200, for its *_gpu entry points."""
import json
import os

import pytest

import common
import replay


@pytest.fixture(scope="module")
def gold():
    with open(replay.GOLDEN) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def ours(libepic_built, tmp_path_factory):
    return replay.build(str(tmp_path_factory.mktemp("replay") / "replay_ours"))


def check(out, gold, extra=()):
    for k, v in gold.items():
        assert out.get(k) == v, "%s: got %s, the reference gives %s" % (k, out.get(k), v)
    for k, v in extra:
        assert out.get(k) == v, "%s: got %s, want %s" % (k, out.get(k), v)


@pytest.mark.parametrize("scenario", ["plan", "node"])
def test_callers_on_the_cpu_exports_match_the_reference(ours, gold, tmp_path, scenario):
    case = replay.plan_case(str(tmp_path)) if scenario == "plan" else replay.node_case(str(tmp_path))
    check(replay.run(ours, case, "cpu"), gold[scenario])


@pytest.mark.skipif(not os.path.isdir(replay.REF_INC), reason="the reference tree is only present in the build container")
def test_golden_is_what_the_reference_sources_produce(gold, tmp_path):
    """Pins the committed golden file: same harness, reference headers, reference CPU sources."""
    exe = replay.build(str(tmp_path / "replay_ref"), include=replay.REF_INC, libdir=replay.REF_LIBDIR,
                       lib="epic_ref_cpu", cpu_only=True)
    check(replay.run(exe, replay.plan_case(str(tmp_path)), "cpu"), gold["plan"])
    check(replay.run(exe, replay.node_case(str(tmp_path)), "cpu"), gold["node"])


@pytest.mark.skipif(not os.path.isdir(replay.REF_INC), reason="the reference tree is only present in the build container")
def test_harness_compiles_against_the_reference_headers_and_links_the_drop_in(libepic_built, tmp_path):
    """Source compatibility: the reference's headers + this repository's libepic.so (all 30 symbols resolve)."""
    replay.build(str(tmp_path / "replay_mixed"), include=replay.REF_INC)


@pytest.mark.gpu
@pytest.mark.parametrize("skip_static", ["1", "all", "0"])
@pytest.mark.parametrize("scenario", ["plan", "node"])
def test_callers_on_the_gpu_match_the_reference(ours, gold, tmp_path, scenario, skip_static, monkeypatch):
    # static-tile skipping inside solves (default), also for the node's update ticks ("all"), and off
    monkeypatch.setenv("EPIC_SKIP_STATIC", skip_static)
    if scenario == "plan":
        out = replay.run(ours, replay.plan_case(str(tmp_path)), "gpu")
        check(out, gold["plan"], extra=[("complete_gpu_result", "0")])
    else:
        out = replay.run(ours, replay.node_case(str(tmp_path)), "gpu")
        check(out, gold["node"], extra=[("gpu_initialised", "1"), ("gpu_uninitialised", "1")])

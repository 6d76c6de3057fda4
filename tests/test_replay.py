"""the two callers of the harmonic path -- the nav_core plugin's makePlan and
the anytime node's tick + services -- replayed without ROS by tests/native/replay_callers.cpp, a C++
program written against the library's public headers and linked with the drop-in libepic.so.

tests/golden/replay_callers.json is the output of the SAME source compiled against the reference's own
headers and linked with the untouched reference CPU sources (tests/golden/make_replay_golden.py); every line a
client would observe (iteration counts, delta bits, hashes of u / locked / raw paths / pose lists) must be
identical for the library's CPU exports and, on a B200, for its *_gpu entry points."""
import json
import os

import numpy as np
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


@pytest.fixture(scope="module")
def ours_dense(libepic_built, tmp_path_factory):
    """The same harness with the node's scatter lists replaced by the dense map-ingest extensions."""
    return replay.build(str(tmp_path_factory.mktemp("replay") / "replay_dense"), dense_ingest=True)


def check(out, gold, extra=()):
    for k, v in gold.items():
        assert out.get(k) == v, "%s: got %s, the reference gives %s" % (k, out.get(k), v)
    for k, v in extra:
        assert out.get(k) == v, "%s: got %s, want %s" % (k, out.get(k), v)


@pytest.mark.parametrize("scenario", ["plan", "node"])
def test_callers_on_the_cpu_exports_match_the_reference(ours, gold, tmp_path, scenario):
    case = replay.plan_case(str(tmp_path)) if scenario == "plan" else replay.node_case(str(tmp_path))
    check(replay.run(ours, case, "cpu"), gold[scenario])


def test_node_with_dense_map_ingest_is_indistinguishable_on_the_cpu_exports(ours_dense, gold, tmp_path):
    check(replay.run(ours_dense, replay.node_case(str(tmp_path)), "cpu"), gold["node"])


@pytest.mark.gpu
def test_node_with_dense_map_ingest_is_indistinguishable_on_the_gpu(ours_dense, gold, tmp_path):
    out = replay.run(ours_dense, replay.node_case(str(tmp_path)), "gpu")
    check(out, gold["node"], extra=[("gpu_initialised", "1"), ("gpu_uninitialised", "1")])


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


@pytest.mark.gpu
def test_plugin_replay_on_umass_matches_the_reference(ours, gold, tmp_path):
    """BASELINE.json config 2(ii): makePlan on maps/umass.png with the plugin's own costmap -> grid conversion;
    the golden lines come from the same harness linked with the untouched reference CPU sources."""
    out = replay.run(ours, replay.plan_umass_case(str(tmp_path)), "gpu")
    check(out, gold["plan_umass"], extra=[("complete_gpu_result", "0")])


@pytest.mark.skipif(not os.environ.get("EPIC_SLOW_TESTS"), reason="two minutes of CPU relaxation: set EPIC_SLOW_TESTS=1")
def test_plugin_replay_on_umass_on_the_cpu_exports(ours, gold, tmp_path):
    check(replay.run(ours, replay.plan_umass_case(str(tmp_path)), "cpu", timeout=1200), gold["plan_umass"])


def _fnv(data, h=1469598103934665603):
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _plan_state(tmp_path):
    """The Harmonic object of the plan scenario after initialize() + setGoal(), built as the harness builds it."""
    from epic_b200.harmonic import Harmonic
    case = replay.plan_case(str(tmp_path))
    raw = np.fromfile(case[1], dtype=np.uint8)
    h, w = (int(v) for v in raw[:8].view(np.uint32))
    cost = raw[8:].reshape(h, w)
    locked = (cost >= 250).astype(np.uint32)
    locked[0, :] = locked[-1, :] = locked[:, 0] = locked[:, -1] = 1
    u = np.full((h, w), -1e6, np.float32)
    gx, gy = int(case[3]), int(case[4])
    u[gy, gx], locked[gy, gx] = 0.0, 1
    x = np.float32((np.float32(float(case[5])) - np.float32(replay.OX)) / np.float32(replay.RES))
    y = np.float32((np.float32(float(case[6])) - np.float32(replay.OY)) / np.float32(replay.RES))
    return Harmonic(u, locked, 1e-3, 100), float(x), float(y), int(np.float32(h * w) / np.float32(0.05))


def _check_poses(hm, x, y, max_length, gold, process):
    r, poses = hm.compute_path_poses(x, y, 0.05, 0.5, max_length, replay.OX, replay.OY, replay.RES, process)
    assert r == 0 and len(poses) == int(gold["plan0_path_points"])
    assert "%016x" % _fnv(poses[1:].tobytes()) == gold["plan0_path_poses"], "pose list differs from the caller's own loop"


def test_pose_lists_equal_what_the_reference_plugin_publishes(libepic_built, gold, tmp_path):
    """harmonic_compute_path_poses_2d_cpu against the pose hash of the replayed plugin (golden from the
    reference sources): world coordinates and yaw in the callers' float arithmetic."""
    hm, x, y, max_length = _plan_state(tmp_path)
    hm.solve(process="cpu")
    assert hm.currentIteration == int(gold["plan"]["plan0_iterations"])
    _check_poses(hm, x, y, max_length, gold["plan"], "cpu")


@pytest.mark.gpu
def test_pose_lists_from_the_device_resident_field(libepic_built, gold, tmp_path):
    hm, x, y, max_length = _plan_state(tmp_path)
    hm.solve(process="gpu")
    hm.initialize_gpu()
    _check_poses(hm, x, y, max_length, gold["plan"], "gpu")
    hm.uninitialize_gpu()

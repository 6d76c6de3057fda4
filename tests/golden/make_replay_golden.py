#!/usr/bin/env python
"""tests/golden/replay_callers.json: the output of tests/native/replay_callers.cpp compiled against the
REFERENCE's own headers and linked with the untouched reference CPU sources (oracle/_ref, -DREPLAY_CPU_ONLY).
Run in the build container (needs /root/reference); the committed JSON travels to the GPU box."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import replay  # noqa: E402

with tempfile.TemporaryDirectory() as tmp:
    exe = replay.build(os.path.join(tmp, "replay_ref"), include=replay.REF_INC, libdir=replay.REF_LIBDIR,
                       lib="epic_ref_cpu", cpu_only=True)
    gold = {"plan": replay.run(exe, replay.plan_case(tmp), "cpu"), "node": replay.run(exe, replay.node_case(tmp), "cpu"),
            "plan_umass": replay.run(exe, replay.plan_umass_case(tmp), "cpu", timeout=1200)}
with open(replay.GOLDEN, "w") as f:
    json.dump(gold, f, indent=1, sort_keys=True)
print(json.dumps(gold, indent=1, sort_keys=True))

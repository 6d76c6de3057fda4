import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import common
from epic_b200.field import Field
for case in ("random256", "proc_maze", "basic", "umass", "maze", "random48x3"):
    u, locked, eps, stagger = common.case_input(case)
    res = {}
    for math in ("strict", "fast"):
        f = Field(u.shape, math=math); f.upload(u, locked)
        it, d = f.solve(eps, stagger)
        res[math] = (it, d, f.download_u()); f.close()
    free = locked == 0
    a, b = res["strict"][2][free], res["fast"][2][free]
    err = np.abs(a - b); tol = 1e-5 * np.abs(a) + 1e-5
    rel = err / np.maximum(np.abs(a), 1.0)
    print("%-10s strict it %d delta %.6g | fast it %d delta %.6g | max abs err %.3g at u=%.5g, max rel err %.3g, mean abs %.3g, violations of 1e-5|u|+1e-5: %d of %d" % (
        case, res["strict"][0], res["strict"][1], res["fast"][0], res["fast"][1], err.max(), a[np.argmax(err)], rel.max(), err.mean(), int((err > tol).sum()), err.size))

# streamlines on the fast field vs on the strict (= reference) field
import json
from epic_b200 import grids
gold = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "golden.json")))
for case in ("umass", "maze", "basic"):
    u, locked, eps, stagger = common.case_input(case)
    starts = [p["start"] for p in gold[case]["paths"] if p["step"] == 0.05] + [list(map(float, s)) for s in grids.free_cells(locked, 100, seed=11)]
    paths = {}
    for math in ("strict", "fast"):
        f = Field(u.shape, math=math); f.upload(u, locked); f.solve(eps, stagger)
        paths[math] = f.paths(starts, 0.05, 0.5, int(u.size / 0.05)); f.close()
    same_cells = 0; maxdev = 0.0; dk = 0; bad = 0
    for (ra, pa), (rb, pb) in zip(paths["strict"], paths["fast"]):
        if ra != rb:
            bad += 1; continue
        if ra != 0:
            continue
        n = min(len(pa), len(pb))
        maxdev = max(maxdev, float(np.abs(pa[:n] - pb[:n]).max()))
        dk = max(dk, abs(len(pa) - len(pb)))
        ca = np.floor(pa + np.float32(0.5)).astype(int); cb = np.floor(pb + np.float32(0.5)).astype(int)
        ka = np.ones(len(ca), bool); ka[1:] = (ca[1:] != ca[:-1]).any(1); kb = np.ones(len(cb), bool); kb[1:] = (cb[1:] != cb[:-1]).any(1)
        same_cells += int(ca[ka].shape == cb[kb].shape and np.array_equal(ca[ka], cb[kb]))
    print("%-6s %d paths: return codes differ %d, max point deviation %.4f cells, max |dk| %d, identical cell sequences %d" % (
        case, len(starts), bad, maxdev, dk, same_cells))

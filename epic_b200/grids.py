"""Occupancy-grid inputs for the harmonic solver: the PNG map format of the reference's Python
wrapper and the seeded synthetic grids named in BASELINE.json's configs.

Encoding everywhere (reference libepic/include/epic/constants.h:37-43): a grid is a pair
(`u` float32, `locked` uint32) of identical shape, last axis fastest; goal cells are
(u=0.0, locked=1), obstacle cells (u=-1e6, locked=1), free cells (u=-1e6, locked=0); every
border cell is locked (precondition stated at libepic/include/epic/harmonic/harmonic.h:35-37).

All generators are deterministic in (shape, parameters, seed) and can produce any slab of rows
[row0, row0+rows) of the full grid without building the rest, so that each rank of a sharded run
builds only what it owns.
"""
import numpy as np

LOG_GOAL = np.float32(0.0)
LOG_OBSTACLE = np.float32(-1e6)
LOG_FREE = np.float32(-1e6)
CELL_GOAL, CELL_OBSTACLE, CELL_FREE = 0, 1, 2


def grid_from_image(image):
    """(u, locked) from a 2-D uint8 grayscale image: 255 = goal, 0 = obstacle, anything else free.
    Same rule as the reference loader, libepic/python/epic/harmonic_map.py:62-100."""
    image = np.asarray(image)
    assert image.ndim == 2
    goal = image == 255
    u = np.where(goal, LOG_GOAL, LOG_FREE).astype(np.float32)
    locked = (goal | (image == 0)).astype(np.uint32)
    return u, locked


def load_png(filename):
    """Read a grayscale PNG map (cv2 as in the reference; PIL if cv2 is unavailable)."""
    try:
        import cv2
        image = cv2.imread(filename, cv2.IMREAD_GRAYSCALE)
    except ImportError:  # pragma: no cover
        from PIL import Image
        image = np.array(Image.open(filename).convert("L"))
    if image is None:
        raise IOError("Failed to load image file '%s'." % filename)
    return grid_from_image(image)


def _goal_cells(shape, goals, seed):
    """`goals` seeded interior cell coordinates, shape (goals, ndim)."""
    rng = np.random.RandomState(seed ^ 0x5EED)
    return np.stack([rng.randint(1, s - 1, size=goals) for s in shape], axis=1)


def _finish(obst, shape, row0, goal_cells):
    """Lock the border, place the goals, build (u, locked) for rows [row0, row0+len(obst))."""
    rows = obst.shape[0]
    glob = np.arange(row0, row0 + rows)
    obst[(glob == 0) | (glob == shape[0] - 1)] = True
    for ax in range(1, len(shape)):
        sl = [slice(None)] * len(shape)
        sl[ax] = 0
        obst[tuple(sl)] = True
        sl[ax] = shape[ax] - 1
        obst[tuple(sl)] = True
    u = np.full(obst.shape, LOG_FREE, dtype=np.float32)
    locked = obst.astype(np.uint32)
    for g in goal_cells:
        if row0 <= g[0] < row0 + rows:
            idx = (int(g[0]) - row0,) + tuple(int(v) for v in g[1:])
            u[idx] = LOG_GOAL
            locked[idx] = 1
    return u, locked


def random_obstacles(shape, p=0.2, goals=64, seed=1234, row0=0, rows=None):
    """BASELINE.json configs 3 and 5: every cell is an obstacle with probability p (one
    RandomState per x0 index, so slabs are independent), border locked, `goals` seeded goal
    cells.  Works for 2-D (rows) and 3-D (planes) shapes."""
    shape = tuple(int(s) for s in shape)
    rows = shape[0] - row0 if rows is None else rows
    obst = np.empty((rows,) + shape[1:], dtype=bool)
    for r in range(rows):
        rng = np.random.RandomState([seed & 0x7FFFFFFF, row0 + r])
        obst[r] = rng.random_sample(shape[1:]) < p
    return _finish(obst, shape, row0, _goal_cells(shape, goals, seed))


def procedural_maze(shape, corridor=8, wall=2, goals=4, seed=1234, row0=0, rows=None):
    """BASELINE.json config 4: a perfect maze (binary-tree construction: every coarse cell opens
    a passage either towards -row or towards +column, chosen by a per-coarse-row seeded coin) with
    corridors `corridor` cells wide and walls `wall` cells thick.  The construction is local, so
    any slab of rows can be generated on its own (one numpy row pattern per coarse row: a 65536-wide
    slab of 8192 rows takes seconds)."""
    shape = tuple(int(s) for s in shape)
    assert len(shape) == 2
    rows = shape[0] - row0 if rows is None else rows
    pitch = corridor + wall
    ncy, ncx = (shape[0] - wall) // pitch, (shape[1] - wall) // pitch
    obst = np.ones((rows, shape[1]), dtype=bool)
    cy_lo = max(0, (row0 - wall) // pitch - 1)
    cy_hi = min(ncy, (row0 + rows) // pitch + 2)
    col = np.arange(shape[1])
    cell_of = np.clip((col - wall) // pitch, 0, max(ncx - 1, 0))
    in_cell = (col >= wall) & (col < wall + ncx * pitch)
    off = col - (wall + cell_of * pitch)              # offset inside the coarse cell: [0, pitch)
    for cy in range(cy_lo, cy_hi):
        rng = np.random.RandomState([seed & 0x7FFFFFFF, cy])
        coin = rng.random_sample(ncx) < 0.5          # True: open towards +column
        coin[ncx - 1] = False                         # last column must open towards -row
        if cy == 0:
            coin[:] = True                            # first row must open towards +column
            coin[ncx - 1] = False
        y0 = wall + cy * pitch                        # first corridor row of this coarse row
        # corridor rows: the corridor itself, plus the wall on the +column side where the coin says so
        open_corr = in_cell & ((off < corridor) | coin[cell_of])
        ya, yb = max(y0, row0), min(y0 + corridor, row0 + rows)
        if ya < yb:
            obst[ya - row0:yb - row0, open_corr] = False
        if cy > 0:
            # the wall rows above: carved through (corridor columns only) where the cell opens towards -row
            open_wall = in_cell & (off < corridor) & ~coin[cell_of]
            ya, yb = max(y0 - wall, row0), min(y0, row0 + rows)
            if ya < yb:
                obst[ya - row0:yb - row0, open_wall] = False
    rng = np.random.RandomState(seed ^ 0x5EED)
    gc = np.stack([wall + rng.randint(0, ncy, size=goals) * pitch + corridor // 2,
                   wall + rng.randint(0, ncx, size=goals) * pitch + corridor // 2], axis=1)
    return _finish(obst, shape, row0, gc)


def free_cells(locked, count, seed=7):
    """`count` seeded (x, y) = (column, row) positions of unlocked cells of a 2-D grid."""
    ys, xs = np.nonzero(np.asarray(locked) == 0)
    rng = np.random.RandomState(seed)
    pick = rng.choice(len(ys), size=min(count, len(ys)), replace=False)
    return [(int(xs[i]), int(ys[i])) for i in pick]

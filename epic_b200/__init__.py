"""epic_b200 -- B200-native libepic (log-space harmonic-function relaxation and streamlines).

The product is the shared library epic_b200/lib/libepic.so (CUDA, sm_100a), a drop-in for the
reference's libepic.so.  This package holds its build recipe (csrc/), the Python mirror of the
reference's ctypes wrapper (libepic.py, harmonic.py), the slab wrapper and the multi-GPU driver
(field.py, sharded.py) and the input generators (grids.py).  There is no CPU fallback in here: a
GPU call on a machine without the library or without a B200 raises / returns libepic's error code.
"""

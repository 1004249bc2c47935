"""Builds tests/golden/urea_atomic_grids.npz: the atomic radial density grids of C, O, N, H exactly as critic2 builds them
(grid1%read_critic, grid1mod@proc.f90:206-318, restated in tests/test_oracle_promolecular_reference.py) from the reference's
data files dat/wfc/{c_,o_,n_,h_}_pbe.wfc, the cutoffs min(cutrad(z), rmax) (global.f90:53-56, crystalmod@env.f90:671-684)
and the library urea structure (dat/lib/crystal.dat).  With it the promolecular density of urea can be evaluated where
/root/reference does not exist (the GPU box) and compared with the 10x10x10 grid that the reference's nodata test
005_plot/016_cube_grid writes (tests/golden/cube_golden.json, "shift"/"plain_text").  Run in the build container."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_oracle_promolecular_reference as T  # noqa: E402

x2c, atoms, ispc, g = T.urea_system()
np.savez_compressed(os.path.join(HERE, "urea_atomic_grids.npz"), x2c=x2c, atoms=atoms, ispc=ispc, ngrid=g.ngrid, off=g.off, a=g.a, b=g.b,
                    rmax=g.rmax, rcut=g.rcut, rtab=g.rtab, ftab=g.ftab, z=np.array([6, 8, 7, 1]))
print("wrote urea_atomic_grids.npz", g.ngrid, os.path.getsize(os.path.join(HERE, "urea_atomic_grids.npz")), "bytes")

#!/usr/bin/env python
"""Write tests/golden/file_boundary_source.sww with the UNMODIFIED Python reference (build container only):
the time-space field that the File_boundary / Field_boundary golden case reads.  A 16 x 16 m basin with a
sloshing mound, stored every 0.5 s for 5 s (the reference's own SWW writer, smoothed vertex storage)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

anuga = pyref.import_anuga()
d = anuga.rectangular_cross_domain(16, 16, len1=16.0, len2=16.0)
d.set_flow_algorithm("DE1")
d.set_name("file_boundary_source")
d.set_datadir(HERE)
d.set_store(True)
d.set_multiprocessor_mode(2)
d.set_quantity("elevation", lambda x, y: -1.0 + 0.02 * x)
d.set_quantity("stage", lambda x, y: 0.3 * np.exp(-((x - 5.0) ** 2 + (y - 9.0) ** 2) / 6.0))
d.set_quantity("friction", 0.01)
Br = anuga.Reflective_boundary(d)
d.set_boundary({t: Br for t in d.get_boundary_tags()})
for t in d.evolve(yieldstep=0.5, finaltime=5.0):
    pass
print("written", os.path.join(HERE, "file_boundary_source.sww"), os.path.getsize(os.path.join(HERE, "file_boundary_source.sww")))

"""Stage times of the hierarchy setup (CSB_SETUP_TRACE=1) on the bench cube: first setup, then
free + setup twice more (host pool warm)."""
import os
import sys
import time

os.environ["CSB_SETUP_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from code_saturne_b200 import api  # noqa: E402
from code_saturne_b200.mesh_cartesian import cube_poisson  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
api.init(0)
s = cube_poisson(N)
m = api.Matrix.from_system(s, share_host=True)
c = api.SlesIt(api.SlesItType.FCG, 200, poly_degree=-1)
mg = api.Multigrid(pc=True)
mg.set_aggregation_blocks(int(os.environ.get("OMP_NUM_THREADS", "0")) or os.cpu_count())
c.transfer_pc(mg)
for k in range(3):
    t0 = time.perf_counter()
    if k:
        c.free()
    t1 = time.perf_counter()
    c.setup(m)
    api.synchronize()
    t2 = time.perf_counter()
    print(f"== setup {k}: free {1e3 * (t1 - t0):.1f} ms, setup {1e3 * (t2 - t1):.1f} ms, "
          f"stages {c.pc().setup_times()}", file=sys.stderr)

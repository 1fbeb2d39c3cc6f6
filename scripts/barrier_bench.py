import ctypes as C, sys
sys.path.insert(0, '/root/repo')
from code_saturne_b200 import api
api.init(0)
L = api.load_library()
L.csb_debug_barrier_bench.restype = C.c_double
for what, name in ((0, "grid_barrier"), (1, "all_sum<3>")):
    L.csb_debug_barrier_bench(100, what)
    print(name, "%.2f us each" % L.csb_debug_barrier_bench(2000, what))

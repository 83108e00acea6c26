import sys, ctypes as C
sys.path.insert(0, "/root/repo")
import sbp_env_b200 as sg
from sbp_env_b200 import _lib
from sbp_env_b200.runtime import get_context
lib = _lib.load(); ctx = get_context(0)
for which, name in ((3, "FFMA2 packed"), (4, "FFMA scalar"), (5, "FMNMX3"), (6, "FMNMX"), (7, "mix: 8 FFMA2 + 8 FMNMX per 16 units")):
    ms, units = C.c_double(), C.c_double()
    _lib.check(lib.sbp_microbench(ctx.handle, which, 0, 4096, C.byref(ms), C.byref(units)))
    rate = units.value / (ms.value / 1e3)
    print(name, "lane-ops/s", rate, "per SM per clk @1965MHz", rate / 148 / 1.965e9)

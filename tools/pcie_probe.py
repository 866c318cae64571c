"""Host <-> device copy bandwidth on this box for the kinds of host memory the end-to-end path can be given."""
import sys, json, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "shakti-fenics_b200"))
import numpy as np, torch
from shakti_b200 import capi
torch.cuda.init(); torch.zeros(1, device="cuda")
lib = capi.load()
n = 16_000_000
def bw(ptr, label):
    h, d = C.c_double(0), C.c_double(0); e = (C.c_double * 2)()
    rc = lib.shakti_debug_copy_bw(C.c_void_p(ptr), C.c_int64(8 * n), C.c_int(4), C.byref(h), C.byref(d), e)
    print(json.dumps(dict(memory=label, h2d_GBps=round(h.value, 1), d2h_GBps=round(d.value, 1), enqueue_ms=[round(e[0], 3), round(e[1], 3)], rc=rc)), flush=True)
pa = capi.PinnedArray(n); pa.array[:] = 1.0
bw(pa.array.ctypes.data, "shakti_alloc_pinned (cudaMallocHost)")
tp = torch.empty(n, dtype=torch.float64).pin_memory(); tp.fill_(1.0)
bw(tp.data_ptr(), "torch pin_memory")
pg = np.ones(n)
bw(pg.ctypes.data, "pageable numpy")
# torch's own measurement
d = torch.empty(n, dtype=torch.float64, device="cuda")
for name, fn in (("torch h2d", lambda: d.copy_(tp, non_blocking=True)), ("torch d2h", lambda: tp.copy_(d, non_blocking=True))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize(); e0.record()
    for _ in range(4): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, round(4 * 8 * n / 1e9 / (e0.elapsed_time(e1) / 1e3), 1), "GB/s")

"""Host <-> device copy bandwidth on this box for the kinds of host memory the end-to-end path can be given;
under torchrun every rank measures its own GPU AT THE SAME TIME (aggregate host-link bandwidth of the box)."""
import os, sys, json, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "shakti-fenics_b200"))
import numpy as np, torch
from shakti_b200 import capi
local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local); torch.zeros(1, device="cuda")
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = capi.load()
n = 16_000_000
def bw(ptr, label, reps=4):
    if dist is not None:
        dist.barrier(); torch.cuda.synchronize()
    h, d = C.c_double(0), C.c_double(0); e = (C.c_double * 2)()
    rc = lib.shakti_debug_copy_bw(C.c_void_p(ptr), C.c_int64(8 * n), C.c_int(reps), C.byref(h), C.byref(d), e)
    print(json.dumps(dict(rank=local, ranks=world, memory=label, h2d_GBps=round(h.value, 1), d2h_GBps=round(d.value, 1),
                          enqueue_ms=[round(e[0], 3), round(e[1], 3)], rc=rc)), flush=True)
pa = capi.PinnedArray(n); pa.array[:] = 1.0
bw(pa.array.ctypes.data, "shakti_alloc_pinned (cudaMallocHost)", reps=24 if world > 1 else 4)
if world == 1:
    tp = torch.empty(n, dtype=torch.float64).pin_memory(); tp.fill_(1.0)
    bw(tp.data_ptr(), "torch pin_memory")
    pg = np.ones(n)
    bw(pg.ctypes.data, "pageable numpy")
if dist is not None:
    dist.destroy_process_group()

import os, sys
from pathlib import Path
import numpy as np
ROOT = Path("/root/repo")
for p in (str(ROOT), str(ROOT / "shakti-fenics_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import torch, torch.distributed as dist
from common import make_case, make_model, make_oracle, relinf
from shakti_b200 import capi
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0: uid = torch.tensor(list(capi.comm_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(uid, 0); capi.comm_init(bytes(uid.cpu().tolist()), rank, world, local)
c = make_case(nx=48, ny=32, seed=9)
for kw in (dict(amg_smoother=0, amg_presmooth=1, amg_postsmooth=1), dict(amg_smoother=1), dict(amg_strength_theta=0.0), dict(amg_prolong_omega=0.0),
           dict(amg_max_levels=1), dict(amg_max_levels=2)):
    m = make_model(*c, device=local, precond="amg", linear_max_it=300, **kw)
    F, _ = m.assemble(3600.0, want_J=False)
    t = torch.from_numpy(F).cuda(); dist.all_reduce(t); F = t.cpu().numpy()
    try:
        dx, it, rr = m.linear_solve(F)
        msg = f"its {it} relres {rr:.2e}"
    except capi.ShaktiError as e:
        msg = "FAIL " + str(e)[:100]
    st = m.stats()
    print(f"rank {rank} {kw} -> {msg} levels {st['amg_levels']} opc {st['amg_operator_complexity']:.2f} n_owned {st['n_owned']}", flush=True)
    m.close()
capi.comm_finalize(); dist.destroy_process_group()

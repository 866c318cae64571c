"""Real multi-GPU parity check (one process per GPU, NCCL): run under torchrun, e.g.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
Every rank holds the global mesh; the library keeps its partition.  Rank 0 compares the summed
owned parts of the fields with the CPU oracle after a few transient steps."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "shakti-fenics_b200"), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def solve_md_leg(rank, world):
    """The reference-shaped entry point solvers.solve(md) on several GPUs: every rank saves its OWNED slice
    through the asynchronous double-buffered path, rank 0 assembles the (nti, nd) arrays and writes the .npy
    files, a checkpoint is written by rank 0 from all ranks' entries; the files must equal a plain stepping of
    the oracle.  Resume on the same number of GPUs continues to the same final fields."""
    import shutil
    import tempfile
    import torch.distributed as dist
    for p in (ROOT / "shakti-fenics_b200" / "source", ROOT / "shakti-fenics_b200" / "setups"):
        if str(p) not in sys.path:
            sys.path.insert(0, str(p))
    import solvers
    from _synthetic import md_from_case
    from common import make_oracle, relinf
    from shakti_b200 import configs, fem
    box = [tempfile.mkdtemp(prefix="shakti_mg_") if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    out = Path(box[0]) / "run"
    nt, nt_save = 8, 2            # nt divisible by nt_save, as every reference setup has it (solvers.py:110)

    def make(nsteps):
        case = configs.rect_steady(nx=60, ny=40, nsteps=nt)
        md = md_from_case(fem.comm_world(), case, str(ROOT / "shakti-fenics_b200" / "setups" / "setup_rect250k.py"),
                          nt_save=nt_save, nt_check=2 * nt_save, results_name=str(out))
        md.timesteps = case.timesteps[:nsteps]
        md.resume = True
        return md, case

    md, case = make(6)           # saves after steps 0, 2, 4; the last checkpoint is the one after step 4
    solvers.solve(md)
    md, case = make(nt)          # resumes from the checkpoint and finishes
    solvers.solve(md)
    good = True
    if rank == 0:
        N, b = np.load(out / "N.npy"), np.load(out / "b.npy")
        bc = solvers.get_bcs(md)[0].dofs
        o = make_oracle(case.xy, case.cells, case.fields, bc, case.N_bdry)
        errs, j = [], 0
        for i, dt in enumerate(o.dt_schedule(case.timesteps[:nt])):
            o.step(dt)
            if i % nt_save == 0 and j < N.shape[0]:
                errs.append(max(relinf(N[j], o.N), relinf(b[j], o.b)))
                j += 1
        good = N.shape == (nt // nt_save, case.n_vert) and j == nt // nt_save and max(errs) < 1e-8
        print(f"[{world} GPUs, solve(md) with owned-slice async saves + checkpoint/resume] rows {N.shape[0]} max field err "
              f"{max(errs):.1e}" + ("  OK" if good else "  MISMATCH"), flush=True)
        shutil.rmtree(box[0], ignore_errors=True)
    return good


def main():
    import torch
    import torch.distributed as dist
    from common import make_case, make_model, make_oracle, relinf
    from shakti_b200 import capi
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(capi.comm_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    capi.comm_init(bytes(uid.cpu().tolist()), rank, world, local)

    def gsum(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        dist.all_reduce(t)
        return t.cpu().numpy()

    ok = True
    # (label, mesh, model options): the AMG runs cover the three shapes of a distributed hierarchy --
    # everything replicated (tiny problem), distributed fine level + replicated rest, and several
    # distributed levels with the replicated part starting where coarsening ends
    runs = [("jacobi", (48, 32), dict(precond="jacobi")),
            ("amg/replicated", (48, 32), dict(precond="amg")),
            ("amg/fine-distributed", (48, 32), dict(precond="amg", amg_replicate_below=400)),
            ("amg/3-level-distributed", (160, 120), dict(precond="amg", amg_replicate_below=0, amg_coarse_size=64)),
            ("amg/no-graph", (96, 64), dict(precond="amg", amg_replicate_below=300, amg_cuda_graph=0)),
            ("amg/bicgstab", (96, 64), dict(precond="amg", amg_replicate_below=300, linear_solver="bicgstab"))]
    only = os.environ.get("SHAKTI_MG_ONLY")
    for label, (nx, ny), opt in runs:
        if only and only not in label:
            continue
        c = make_case(nx=nx, ny=ny, seed=9)
        m = make_model(*c, device=local, linear_max_it=5000, **opt)
        st = m.stats()
        assert st["n_owned"] < st["n_vert"] and st["n_local"] > st["n_owned"], st
        F, J = m.assemble(3600.0)
        F, J = gsum(F), gsum(J)
        dts = [360.0, 3600.0, 3600.0, 3600.0]
        its = list(m.run(dts))
        fields = {k: gsum(m.get_field(k)) for k in ("N", "b", "melt_n", "N_n")}
        q = gsum(m.get_flux())
        st = m.stats()
        if rank == 0:
            o = make_oracle(*c)
            Fo, Jo = o.assemble(3600.0)
            its_o = [o.step(dt)[0] for dt in dts]
            errs = dict(F=relinf(F, Fo), J=relinf(J, Jo), N=relinf(fields["N"], o.N), b=relinf(fields["b"], o.b),
                        melt=relinf(fields["melt_n"], o.melt_n), q=relinf(q, o.q), N_n=relinf(fields["N_n"], o.N_n))
            good = errs["F"] < 1e-12 and errs["J"] < 1e-12 and all(errs[k] < 1e-8 for k in ("N", "b", "melt", "q", "N_n")) \
                and its == its_o
            ok &= good
            print(f"[{world} GPUs, {label}, {nx}x{ny}] newton {its} (oracle {its_o}) krylov {st['linear_its']} amg levels "
                  f"{st['amg_levels']} errs " + " ".join(f"{k}={v:.1e}" for k, v in errs.items())
                  + ("  OK" if good else "  MISMATCH"), flush=True)
        m.close()
    ok &= solve_md_leg(rank, world)
    capi.comm_finalize()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK " + ("PASSED" if ok else "FAILED"), flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

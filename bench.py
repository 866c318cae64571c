#!/usr/bin/env python
"""Headline benchmark: SHAKTI transient time-steps/s at 16M dofs (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # the B200 path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on host cores

A "step" is one pass of reference source/solvers.py:179-229 (Newton solve for N with F+J
assembly and the Krylov/AMG linear solves, then the q / melt_n / b nodal updates) on the
synthetic config C4 of SURVEY.md §8d (4000 x 4000 vertices, jittered margin mesh, turbulent
K(b,Re)); N > 1 partitions that same mesh over N GPUs (strong scaling).  One JSON line is
printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for _p in (str(ROOT), str(ROOT / "shakti-fenics_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "time_steps_per_sec_at_16M_dofs"
UNIT = "steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nside", type=int, default=4000, help="vertices per side (4000 -> 16M dofs, the headline config)")
    ap.add_argument("--precond", default="amg")
    ap.add_argument("--linear-rtol", type=float, default=1e-12)
    ap.add_argument("--amg-refresh-every", type=int, default=None, help="override shakti_options.amg_refresh_every")
    ap.add_argument("--lagged-smoother-halo", action="store_true", help="multi-GPU: amg_smoother_halo = 0")
    ap.add_argument("--linear-forcing", type=float, default=None, help="override shakti_options.linear_forcing (0 = fixed tolerance)")
    ap.add_argument("--no-graph", action="store_true", help="do not replay the AMG V-cycle as a CUDA graph")
    ap.add_argument("--cpu-sample-nside", type=int, default=None,
                    help="mesh side of the bounded CPU sample (default: sized so that the CPU run takes ~2 minutes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample_nside(requested, steps):
    """Bounded CPU sample: 70-200 us of oracle time per dof and step (measured on the host cores of
    the GPU box and of the build container), sized for one to three minutes in total."""
    if requested:
        return requested
    dofs = 120.0 / (max(steps, 1) * 1.2e-4)
    return int(min(500, max(100, dofs ** 0.5)))


def cpu_oracle_rate(nside, steps, warmup, target_dofs):
    """Time the CPU oracle (numpy assembly + SuperLU, the stand-in for FEniCSx/PETSc LU) on a
    bounded sample: the same C4 fields on an nside x nside sub-size mesh.  Returns steps/s
    extrapolated LINEARLY in dofs to the target size (optimistic for the CPU: sparse LU is
    super-linear), the raw steps/s and a description."""
    from oracle.shakti_oracle import ShaktiOracle
    from shakti_b200 import configs
    case = configs.dofs16m(nside=nside, nsteps=max(steps + warmup + 1, 4))
    o = ShaktiOracle(case.xy, case.cells)
    for k in ("z_b", "z_s", "G", "inputs", "storage", "b", "N_n", "melt_n"):
        getattr(o, k)[:] = case.fields[k]
    o.q[:] = case.fields["q"]
    o.set_dirichlet(case.bc_dofs, case.N_bdry)
    o.start()
    dts = case.dts(steps + warmup)
    for dt in dts[:warmup]:
        o.step(dt)
    t0 = time.perf_counter()
    its = [o.step(dt)[0] for dt in dts[warmup:]]
    el = time.perf_counter() - t0
    raw = steps / el
    scaled = raw * (case.n_vert / float(target_dofs))
    sample = (f"{steps} oracle steps (numpy assembly + scipy SuperLU, 1 thread) on a {nside}x{nside}-vertex C4 mesh "
              f"({case.n_vert} dofs, {np.mean(its):.1f} Newton its/step): {raw:.4g} steps/s measured, scaled by "
              f"dofs ratio {case.n_vert}/{target_dofs} (linear; optimistic for LU)")
    return scaled, raw, sample, el


def run_reference(args):
    """--impl reference: the reference's algorithm on the host CPU (oracle port: FEniCSx/PETSc
    are not installable here).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    target = args.nside * args.nside
    nside_cpu = cpu_sample_nside(args.cpu_sample_nside, args.steps + 1)
    scaled, raw, sample, el = cpu_oracle_rate(nside_cpu, args.steps, min(args.warmup, 1), target)
    line = {
        "impl": "reference", "metric": METRIC, "value": scaled, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 / scaled, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C4 synthetic ice-sheet margin mesh {args.nside}x{args.nside} vertices ({target} dofs), "
                               "turbulent K(b,Re), dt=3600 s", "sample_nside": nside_cpu},
        "cpu_baseline": {"value": scaled, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "raw_steps_per_sec_on_sample": raw},
        "e2e": {"value": scaled, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    from shakti_b200 import capi, configs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(capi.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        capi.comm_init(bytes(uid.cpu().tolist()), rank, world, local_rank)
    assert args.gpus == world, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    t_setup = time.perf_counter()
    nsteps_total = args.warmup + 2 * args.steps + 4
    case = configs.dofs16m(nside=args.nside, nsteps=nsteps_total)
    nv = case.n_vert
    extra = {} if args.amg_refresh_every is None else {"amg_refresh_every": args.amg_refresh_every}
    if args.linear_forcing is not None:
        extra["linear_forcing"] = args.linear_forcing
    if args.no_graph:
        extra["amg_cuda_graph"] = 0
    if args.lagged_smoother_halo:
        extra["amg_smoother_halo"] = 0
    m = capi.Model(case.xy, case.cells, device=local_rank, precond=args.precond, linear_rtol=args.linear_rtol, **extra)
    configs.apply_case(m, case)
    dts = case.dts()
    setup_s = time.perf_counter() - t_setup

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (untimed)
    its_w = m.run(dts[: args.warmup])
    st0 = m.stats()
    # ---- timed region: exactly K steps, device time on the library stream, max over ranks
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    its, ms = m.run_timed(dts[args.warmup: args.warmup + args.steps])
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    st1 = m.stats()
    value = args.steps / (ms / 1e3)
    launches = st1["kernel_launches"] - st0["kernel_launches"]

    # ---- end-to-end: host buffers in the call, H2D of the forcing + D2H of b, N, qx, qy per step
    e2e = None
    if not args.no_e2e:
        pin = lambda: torch.empty(nv, dtype=torch.float64, pin_memory=True)
        h_in = pin()
        h_in.numpy()[:] = case.fields["inputs"]
        outs = [pin() for _ in range(4)]
        k0 = args.warmup + args.steps
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            m.step_host(dts[k0 + i], h_in.data_ptr(), *[o.data_ptr() for o in outs])
        barrier()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        e2e = {"value": args.steps / float(el.item()), "unit": UNIT, "h2d_bytes_per_step": 8 * nv,
               "d2h_bytes_per_step": 4 * 8 * nv,
               "how": "shakti_step_host: pinned host buffers, forcing H2D + b,N,qx,qy D2H every step (nt_save=1)"}

    # ---- roofline of the dominant kernel (SELL SpMV of the fine Jacobian), live CUDA events
    peak, peak_src = peaks()
    kern = {}
    for k in ("spmv", "assemble", "kbar", "nodal", "dot", "axpy"):
        t_ms = m.time_kernel(k, reps=20, dt=3600.0)
        by = m.kernel_bytes(k)
        kern[k] = {"ms": t_ms, "algorithmic_GB": by / 1e9, "GBps": by / 1e9 / (t_ms / 1e3), "frac": by / 1e9 / (t_ms / 1e3) / peak}
    traffic = None
    tf = ROOT / "profiles" / "r1_traffic.json"
    if tf.exists() and world == 1:
        t_ = json.loads(tf.read_text())["spmv_fine_fp64"]
        if t_["nside"] == args.nside:
            traffic = t_["traffic_bytes"]        # dram read+write per launch, ncu --set full (profiles/)
    roofline = {"bound": "hbm", "kernel": "spmv_sell_kernel<0,double> (fine Jacobian, fp64 SELL-32)", "achieved": kern["spmv"]["GBps"],
                "peak": peak, "unit": "GB/s", "frac": kern["spmv"]["frac"], "traffic": traffic,
                "algorithmic_bytes": m.kernel_bytes("spmv"), "peak_source": peak_src,
                "how": "12*nnz+20*Nv algorithmic bytes / mean of 20 launches, CUDA events on the library stream, matrix >> L2"}

    if rank != 0:
        return
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        scaled, raw, sample, _ = cpu_oracle_rate(cpu_sample_nside(args.cpu_sample_nside, 5), 2, 1, nv)
        cpu_baseline = {"value": scaled, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                        "raw_steps_per_sec_on_sample": raw}
    n_newton = st1["newton_its"] - st0["newton_its"]
    n_krylov = st1["linear_its"] - st0["linear_its"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C4 synthetic ice-sheet margin mesh {args.nside}x{args.nside} vertices ({nv} dofs, "
                               f"{case.cells.shape[0]} P1 triangles), turbulent K(b,Re), dt=3600 s",
                   "parallelism": f"mesh partition x{world}" if world > 1 else "single GPU",
                   "l2": "inputs larger than L2 (matrix + vectors >> 126 MB)",
                   "newton_its_per_step": n_newton / args.steps, "krylov_its_per_solve": n_krylov / max(n_newton, 1),
                   "linear_solver": "gmres", "precond": args.precond, "linear_rtol": args.linear_rtol,
                   "amg_levels": st1["amg_levels"], "amg_refreshes_in_timed_region": st1["amg_refreshes"] - st0["amg_refreshes"], "amg_operator_complexity": st1["amg_operator_complexity"],
                   "setup_seconds": setup_s, "warmup_newton_its": [int(v) for v in its_w]},
        "roofline": roofline, "kernels": kern, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    m.close()
    if dist is not None:
        capi.comm_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the per-timestep hot path: full-timestep Mcell-updates/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one pass of the example loop body: ``dt = compute_stable_timestep()``
then ``time_step(dt)`` (reference ``examples/3d_examples/*``), on synthetic fields.
Default workload at N=1: BASELINE.json configs[1], the 256^3 float32 vortex ring
(flow_type "navier_stokes": rotational-form update + diffusion + penalise + unbounded
FFT Poisson + curl).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

# stdout carries ONE JSON line: while the benchmark runs, file descriptor 1 points at stderr so that
# anything the libraries print (NCCL's version banner, ...) stays out of it; emit() restores it
_STDOUT_FD = None


def capture_stdout():
    """Called by main() only (importing this module must not touch the descriptors)."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)
    if _STDOUT_FD is not None:
        os.dup2(2, 1)


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "full_timestep_Mcell_updates_per_s"
UNIT = "Mcell-updates/s"
FALLBACK_HBM_GBS = 6650.0
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the fused z kernel, from the
# `ncu --set full` captures kept under profiles/ (r01_ncu_step256_summary.txt, r01_ncu_z512_summary.txt)
NCU_Z_PASS_TRAFFIC_BYTES = {"vortex_ring_256_f32": 1.076057e9 + 0.761642e9,
                            "vortex_ring_512_f32": 8.638908e9 + 6.418693e9}

WORKLOADS = {
    # name: (grid (z,y,x), flow_type, with immersed body)
    "vortex_ring_256_f32": ((256, 256, 256), "navier_stokes", False),
    "vortex_ring_128_f32": ((128, 128, 128), "navier_stokes", False),
    "vortex_ring_512_f32": ((512, 512, 512), "navier_stokes", False),
    "sphere_vbf_512x256x256_f32": ((256, 256, 512), "navier_stokes_with_forcing", True),
    # BASELINE configs[3]: rod in cross-flow; x_range 1.8, Laplacian filter order 1 (multiplicative),
    # synthetic rod surface grid (160 rings x 64 points + caps ~ 10 k points), 3 velocity
    # interpolations (rod sub-steps) + 1 full interaction per flow step
    "rod_fsi_512x256x256_f32": ((256, 256, 512), "navier_stokes_with_forcing", "rod"),
    "fsi_512_f32": ((512, 512, 512), "navier_stokes_with_forcing", True),
}


def rod_surface_points(x_range, y_range, z_range, n_elem=160, n_ring=64):
    """Straight rod of length 1 along x, diameter y_range / 5, centred in the box."""
    length, radius = 1.0, y_range / 10.0
    x0 = 0.5 * (x_range - length)
    xs = x0 + (np.arange(n_elem) + 0.5) * length / n_elem
    th = 2 * np.pi * np.arange(n_ring) / n_ring
    ring = np.stack([np.zeros(n_ring), radius * np.cos(th), radius * np.sin(th)])
    pts = [np.stack([np.full(n_ring, x), 0.5 * y_range + ring[1], 0.5 * z_range + ring[2]]) for x in xs]
    for xc in (x0, x0 + length):  # caps: concentric rings
        for frac in (0.25, 0.5, 0.75):
            m = max(int(n_ring * frac), 4)
            t2 = 2 * np.pi * np.arange(m) / m
            pts.append(np.stack([np.full(m, xc), 0.5 * y_range + frac * radius * np.cos(t2),
                                 0.5 * z_range + frac * radius * np.sin(t2)]))
    return np.concatenate(pts, axis=1)


def vortex_ring(x, y, z, real_t):
    """Gaussian-core vortex ring: R=0.125, a=0.2R, centre
    (0.5,0.5,0.35), axis +z, Gamma=1.  x,y,z: broadcastable coordinate arrays."""
    R, gamma = 0.125, 1.0
    a = 0.2 * R
    rho = np.sqrt((x - 0.5) ** 2 + (y - 0.5) ** 2)
    w_theta = gamma / (np.pi * a * a) * np.exp(-((rho - R) ** 2 + (z - 0.35) ** 2) / (a * a))
    rho = np.maximum(rho, 1e-12)
    wx = -w_theta * (y - 0.5) / rho
    wy = w_theta * (x - 0.5) / rho
    return np.stack([wx + 0 * z, wy + 0 * z, 0 * wx + 0 * z]).astype(real_t)


def sphere_points(centre, diameter, spacing):
    pts = []
    r = diameter / 2
    n_lat = max(int(np.pi * r / spacing), 2)
    for i in range(n_lat + 1):
        th = np.pi * i / n_lat
        n_lon = max(int(2 * np.pi * r * np.sin(th) / spacing), 1)
        for j in range(n_lon):
            ph = 2 * np.pi * j / n_lon
            pts.append([centre[0] + r * np.sin(th) * np.cos(ph), centre[1] + r * np.sin(th) * np.sin(ph),
                        centre[2] + r * np.cos(th)])
    return np.array(pts).T.copy()


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    def __init__(self, device_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device_index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        self.first = 0

    def mark(self):
        """Samples taken from now on belong to the timed region."""
        try:
            self.first = sum(1 for _ in open(self.path))
        except OSError:
            self.first = 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = open(self.path).read().splitlines()
        timed = lines[self.first:]
        # a 40 ms timed region may see no 20 ms sample of its own: fall back to the last samples of
        # the (>= 1 s, same workload) warm-up that precedes it
        for line in (timed if len(timed) >= 2 else lines[-10:]):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------ reference arm
def run_cpu_oracle(grid, flow_type, steps, warmup, real_t=np.float32):
    """The reference's CPU path restated (oracle/, numpy + scipy.fft on all host cores)."""
    from oracle.simulator import FlowSimulatorOracle3D

    cores = os.cpu_count() or 1
    sim = FlowSimulatorOracle3D(grid, 1.0, 1.0 / 1000.0, flow_type=flow_type, real_t=real_t,
                                fft_workers=cores)
    x, y, z = sim.local_x[None, None, :], sim.local_y[None, :, None], sim.local_z[:, None, None]
    sim.vorticity_field[...] = vortex_ring(x, y, z, real_t)
    sim.compute_flow_velocity((0.0, 0.0, 0.0))
    for _ in range(warmup):
        sim.time_step(sim.compute_stable_timestep(), (0.0, 0.0, 0.0))
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.time_step(sim.compute_stable_timestep(), (0.0, 0.0, 0.0))
    dt = (time.perf_counter() - t0) / max(steps, 1)
    cells = float(np.prod(grid))
    return cells / dt / 1e6, dt * 1e3, cores


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    grid = (128, 128, 128)
    name = args.workload or "vortex_ring_256_f32"
    # the CPU arm times the Eulerian step (rotational-form NS + Poisson) of the workload
    value, ms, cores = run_cpu_oracle(grid, "navier_stokes", args.steps, args.warmup)
    sample = (f"oracle/ (numpy + scipy.fft restatement of the reference CPU path; the reference itself needs "
              f"mpi4py/sopht/pystencils which are absent) on a {grid[0]}^3 float32 sub-sample of the "
              f"workload, scipy.fft workers={cores}, numpy stencils single-threaded")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "sample_grid": list(grid)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------- own arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", type=str, default=None)
    ap.add_argument("--impl", type=str, default="b200")
    ap.add_argument("--scaling", type=str, default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fused", action="store_true")
    ap.add_argument("--poisson-backend", type=str, default="auto")
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from sopht_mpi_b200 import _lib
    from sopht_mpi_b200.simulator import (PrescribedForcingGrid, RigidBodyFlowInteractionMPI,
                                          UnboundedFlowSimulator3D)
    from sopht_mpi_b200.utils.comm import init_process_group_if_needed

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    _lib.load()
    init_process_group_if_needed()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)

    name = args.workload or "vortex_ring_256_f32"
    grid, flow_type, with_body = WORKLOADS[name]
    grid = list(grid)
    if args.scaling == "weak" and world > 1:
        # fixed cells per GPU: the box doubles along z, then y, then x (256^3 -> 512x256x256 -> 512x512x256
        # -> 512^3 on 1/2/4/8 GPUs, BASELINE configs[4]); the decomposition stays z-slabs
        f, axis = world, 0
        while f > 1:
            grid[axis % 3] *= 2
            f //= 2
            axis += 1
    real_t = np.float32
    nu = 1.0 / 1000.0  # Re_Gamma = 1000
    is_rod = with_body == "rod"
    with_body = bool(with_body)
    extra = dict(filter_vorticity=True, filter_setting_dict={"order": 1, "type": "multiplicative"}) if is_rod else {}
    sim = UnboundedFlowSimulator3D(
        grid_size=tuple(grid), x_range=1.8 if is_rod else 1.0, kinematic_viscosity=nu, flow_type=flow_type,
        real_t=real_t, with_free_stream_flow=with_body, use_fused_kernels=not args.no_fused,
        poisson_backend=args.poisson_backend, rank_distribution=(0, 1, 1), **extra)
    gs = sim.ghost_size
    x = sim.local_x[None, None, :].astype(np.float64)
    y = sim.local_y[None, :, None].astype(np.float64)
    # keep the ring inside the (possibly non-cubic) box: scale every axis to the unit cube
    x = x / sim.x_range
    y = y / sim.y_range
    z = sim.local_z[:, None, None].astype(np.float64) / sim.z_range
    w0 = vortex_ring(x, y, z, real_t)
    host_w = torch.from_numpy(w0).pin_memory()
    sim.vorticity_field[...] = host_w.to(device)
    u_inf = [1.0, 0.0, 0.0] if with_body else [0.0, 0.0, 0.0]
    sim.compute_flow_velocity(free_stream_velocity=u_inf)

    interactor = None
    n_lag = 0
    if with_body:
        dx = float(sim.dx)
        diameter = 0.4 * min(grid[0], grid[1]) / grid[2]
        if is_rod:
            pts = rod_surface_points(sim.x_range, sim.y_range, sim.z_range)
            lag_vel = 0.01 * np.random.default_rng(1234).standard_normal(pts.shape)
        else:
            pts = sphere_points((0.25, 0.5 * sim.y_range, 0.5 * sim.z_range), diameter, dx)
            lag_vel = None
        n_lag = pts.shape[1]

        class _Body:
            pass

        interactor = RigidBodyFlowInteractionMPI(
            mpi_construct=sim.mpi_construct,
            mpi_ghost_exchange_communicator=sim.mpi_ghost_exchange_communicator,
            rigid_body=_Body(), eul_grid_forcing_field=sim.eul_grid_forcing_field,
            eul_grid_velocity_field=sim.velocity_field, virtual_boundary_stiffness_coeff=-6e5 / 4,
            virtual_boundary_damping_coeff=-3.5e2 / 4, dx=sim.dx, grid_dim=3,
            forcing_grid_cls=lambda grid_dim, rigid_body: PrescribedForcingGrid(
                grid_dim, pts, max_lag_grid_dx=dx))

    cells = float(np.prod(grid))

    def one_step():
        dt = sim.compute_stable_timestep(dt_prefac=0.5)
        if interactor is not None:
            if is_rod:  # rod sub-steps: flow forces on the body only (velocity interpolation + force)
                for _ in range(3):
                    interactor.compute_flow_forces_and_torques()
            interactor()
            interactor.time_step(dt)
        sim.time_step(dt=dt, free_stream_velocity=u_inf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # W warm-up steps, extended (never shortened) until the GPU has been busy for about a second: a
    # 256^3 step is 2 ms, and three of them do not bring the SM clock up from idle.  The number of
    # extra steps is decided on rank 0 and broadcast (the steps contain collectives).
    n_warm = max(args.warmup, 3)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_warm = time.perf_counter()
    for _ in range(n_warm):
        one_step()
    barrier()
    per_step = max((time.perf_counter() - t_warm) / n_warm, 1e-5)
    extra = torch.tensor([max(0, min(400, int(1.0 / per_step)) - n_warm)], dtype=torch.int64, device=device)
    if world > 1:
        dist.broadcast(extra, src=0)
    for _ in range(int(extra.item())):
        one_step()
    n_warm += int(extra.item())
    barrier()
    if sampler:
        sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        one_step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    clocks = sampler.stop() if sampler else {}

    # ---- per-stage device times (untimed extra pass) -> roofline of the dominant stage
    stage_ms = {}

    def timed(label, fn, reps=5):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        stage_ms[label] = a.elapsed_time(b) / reps

    dt_fix = sim.compute_stable_timestep(dt_prefac=0.5)
    solver = sim.unbounded_poisson_solver
    timed("poisson_vector_solve", lambda: solver.vector_field_solve(
        solution_vector_field=sim.stream_func_field, rhs_vector_field=sim.vorticity_field))
    timed("full_step", lambda: sim.time_step(dt=dt_fix, free_stream_velocity=u_inf))
    w_bytes = 4
    local_cells = cells / world
    peak, peak_src = hbm_peak()
    # ---- roofline of the dominant kernel: the fused z pass of the Poisson solve (forward FFT x Green's
    # spectrum x inverse FFT, in place).  Its launches are timed live with CUDA events recorded by the
    # library on the solve's stream (sb200_poisson_set_profiling), averaged over `steps` solves.
    # ALGORITHMIC bytes per launch: read 4 W + write 4 W per cell and component.
    roofline = None
    if solver.backend == "fft":
        solver.set_profiling(True)
        acc = {}
        for _ in range(max(args.steps, 5)):
            solver.vector_field_solve(solution_vector_field=sim.stream_func_field,
                                      rhs_vector_field=sim.vorticity_field)
            for k, v in solver.last_stage_ms().items():
                acc[k] = acc.get(k, 0.0) + v / max(args.steps, 5)
        solver.set_profiling(False)
        stage_ms.update({"poisson_" + k: v for k, v in acc.items()})
        zk = "z_fused_forward_green_inverse"
        if zk not in acc:  # z-slabs: the y and z passes are timed together (20 W per cell and component)
            zk = "y_forward_z_fused_y_inverse"
        z_bytes = (8 if zk.startswith("z_") else 20) * w_bytes * 3 * local_cells
        achieved = z_bytes / (acc[zk] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "sb_fft_strided32_kernel<float, MODE 1> (fused z pass of the Poisson "
                    "vector solve: forward FFT x Green x inverse FFT, in place)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": NCU_Z_PASS_TRAFFIC_BYTES.get(name) if world == 1 else None,
                    "peak_source": peak_src,
                    "launch_ms": acc[zk], "algorithmic_bytes_per_launch": z_bytes,
                    "algorithmic_bytes_per_cell": z_bytes / local_cells,
                    "share_of_step": acc[zk] / stage_ms["full_step"],
                    "per_gpu": world > 1,
                    "note": "the kernel is co-limited by FP32 issue (radix-16 butterflies, ~33 lane-ops per "
                            "complex point and transform) and HBM; see profiles/r01_fft_tuning.md"}
    nvlink = None
    if world > 1 and "poisson_all_to_all_z_to_kx" in stage_ms:
        # each all-to-all moves the x-pass output (nx/2 + 1 complex bins per row = 2 W per cell) of this
        # GPU's planes, minus the block that stays local
        kxl = ((grid[2] + 1 + world - 1) // world + 3) // 4 * 4
        a2a_bytes = 2 * w_bytes * 3 * (grid[0] / world) * grid[1] * kxl * (world - 1)
        nvlink = {"all_to_all_bytes_out_per_gpu": a2a_bytes,
                  "GBps_out_per_gpu": [a2a_bytes / (stage_ms[k] * 1e-3) / 1e9
                                       for k in ("poisson_all_to_all_z_to_kx", "poisson_all_to_all_kx_to_z")],
                  "peak_GBps_per_direction": 900.0,
                  "exchange": getattr(solver, "exchange_mode", None)}
    poisson_bytes = 86 * w_bytes * local_cells
    poisson_gbs = poisson_bytes / (stage_ms["poisson_vector_solve"] * 1e-3) / 1e9
    step_bytes = (107 if flow_type == "navier_stokes_with_forcing" else 101) * w_bytes * local_cells
    step_gbs = step_bytes / (ms_step * 1e-3) / 1e9

    # ---- end to end through the operator API with HOST buffers: every step uploads the
    # vorticity field from pinned host memory and reads vorticity + velocity back
    h2d = host_w.numel() * host_w.element_size()
    host_out_w = torch.empty_like(host_w).pin_memory()
    host_out_u = torch.empty_like(host_w).pin_memory()
    d2h = 2 * h2d
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        sim.vorticity_field.tensor.copy_(host_w, non_blocking=True)
        one_step()
        host_out_w.copy_(sim.vorticity_field.tensor, non_blocking=True)
        host_out_u.copy_(sim.velocity_field.tensor, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = cells / e2e_s.item() / 1e6

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms, cores = run_cpu_oracle((128, 128, 128), "navier_stokes", 3, 1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": ms,
                        "sample": "oracle/ restatement of the reference CPU path (numpy + scipy.fft, "
                                  f"workers={cores}) on a 128^3 float32 sub-sample, 3 steps"}

    # every rank runs the counted step (it contains collectives)
    launches_per_step = count_launches(sim, interactor, u_inf)
    if rank == 0:
        line = {
            "metric": METRIC, "value": cells / (ms_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": name, "grid_zyx": grid, "flow_type": flow_type, "lagrangian_points": n_lag,
                       "ghost_size": gs, "parallelism": f"z-slabs x{world}",
                       "poisson_backend": sim.unbounded_poisson_solver.backend,
                       "l2": "working set (>= 1.2 GB of fields per step) is far larger than the 126 MB L2",
                       "step_algorithmic_GBps": step_gbs, "step_hbm_frac_of_measured": step_gbs / peak,
                       "step_hbm_frac_of_nominal_8TBps": step_gbs / 8000.0,
                       "poisson_algorithmic_GBps": poisson_gbs, "poisson_hbm_frac_of_measured": poisson_gbs / peak,
                       "stage_ms": stage_ms, "nvlink": nvlink},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "operator API with host buffers: vorticity H2D from pinned memory, step, "
                            "vorticity+velocity D2H, every step"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        emit(line)
    if world > 1:
        # release the CUDA-IPC mappings of the other ranks' exchange buffers before any rank exits
        import gc

        del solver, sim, interactor
        gc.collect()
        torch.cuda.ipc_collect()
        dist.barrier()
        dist.destroy_process_group()


def count_launches(sim, interactor, u_inf):
    """kernels of libsophtb200 launched by one step (counted by wrapping the ctypes calls)."""
    from sopht_mpi_b200 import _lib

    lib = _lib.load()
    counts = {"n": 0}
    per_call = {"sb200_diffusion_timestep": 6, "sb200_laplacian_filter": 12, "sb200_penalise_field_boundary": 2,
                "sb200_poisson_solve": 5, "sb200_poisson_slab_spectral": 3, "sb200_velocity_from_stream_function": 2, "sb200_max_abs_sum": 2}
    originals = {}
    for fname in _lib.PROTOTYPES:
        fn = getattr(lib, fname)
        originals[fname] = fn

        def wrap(*a, _fn=fn, _name=fname):
            counts["n"] += per_call.get(_name, 1)
            return _fn(*a)

        setattr(lib, fname, wrap)
    try:
        dt = sim.compute_stable_timestep(dt_prefac=0.5)
        if interactor is not None:
            interactor()
        sim.time_step(dt=dt, free_stream_velocity=u_inf)
    finally:
        for fname, fn in originals.items():
            setattr(lib, fname, fn)
    return counts["n"]


if __name__ == "__main__":
    main()

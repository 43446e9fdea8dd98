#!/usr/bin/env python
"""bench.py -- particle-steps/s of the gevolution per-step particle-mesh hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (ours; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU reference arm)

A "step" is one cycle of the reference's main loop (main.cpp:372-879 without outputs):
deposit T00/Tij -> metric solve (12 FFTs) -> kick -> drift -> re-bin, GR, parabolic B.
Workload: the configuration BASELINE.json quotes the metric on -- 512^3 grid / 512^3
particles (23.7 GB of fields + 7.5 GB of particles; fits one 180 GB B200), strong scaling
over N z-slabs.  Synthetic data: quasi-uniform particles (one per cell, sigma = 0.05 cell,
q ~ N(0,(1e-3 a)^2)) and smooth analytic metric fields of cosmological amplitude.

`value`  : device-resident throughput (inputs in HBM when the timed region starts).
`e2e`    : same metric through the C ABI with HOST buffers -- every step uploads the particle
           and metric state from pinned host memory and downloads the updated state.
`roofline`, `kernels`: per-entry-point device time over the timed region (CUDA events on the
           library's stream) against algorithmic bytes.
`cpu_baseline`: the reference's own gevolution.hpp (oracle/_ref, single-rank LATfield2 shim)
           timed on the host cores on a bounded sample (smaller lattice, same path).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "gevolution-1.2_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "particle-steps/s (GR, 512^3)"
UNIT = "particle-steps/s"

# BASELINE.json configs (1-based as in SURVEY 8d; config 1 is the shipped settings.ini run, covered by tests/ and
# scripts/run_settings.py).  The bench line is quoted on config 3; the others are run on request (--config) and their
# lines are kept under profiles/.
CONFIGS = {
    2: dict(ngrid=256, species=("cdm",), vector_flag=0, hij=False, what="256^3 grid / 256^3 CDM particles, GR (phi, chi, B_i, Tij sources)"),
    3: dict(ngrid=512, species=("cdm",), vector_flag=0, hij=False, what="512^3 grid / 512^3 particles, GR, parabolic B"),
    4: dict(ngrid=512, species=("cdm", "b", "ncdm0", "ncdm1"), vector_flag=0, hij=False,
            what="512^3 grid, cdm + baryon + 2 massive-neutrino (0.1, 0.2 eV) species of 512^3 particles each, GR; synthetic ICs: the ncdm species are "
                 "a jittered lattice with Fermi-Dirac thermal momenta (the shipped class_tk.dat has no ncdm columns and CLASS is unavailable, SURVEY 8d)"),
    5: dict(ngrid=1024, species=("cdm",), vector_flag=1, hij=True, what="1024^3 grid / 1024^3 particles, GR with elliptic vector method (T0i deposit + projectFTvector) and an hij spectrum call"),
}


def env_int(name, default):
    return int(os.environ.get(name, default))


# ----------------------------------------------------------------------------- synthetic data
def local_particles(N, z0, nzl, a, seed):
    """one particle per cell of the local slab, centre + N(0, 0.05 cell) (clipped inside the cell)"""
    rng = np.random.default_rng(seed + 1000 * z0)
    g = (np.arange(N) + 0.5) / N
    gz = (np.arange(z0, z0 + nzl) + 0.5) / N
    n = nzl * N * N
    pos = np.empty((n, 3))
    pos[:, 0] = np.tile(g, nzl * N)
    pos[:, 1] = np.tile(np.repeat(g, N), nzl)
    pos[:, 2] = np.repeat(gz, N * N)
    disp = rng.standard_normal((n, 3))
    disp *= 0.05 / N
    np.clip(disp, -0.45 / N, 0.45 / N, out=disp)
    pos += disp
    vel = rng.standard_normal((n, 3))
    vel *= 1e-3 * a
    ids = np.arange(z0 * N * N, z0 * N * N + n, dtype=np.int64)
    return ids, pos, vel


def fermi_dirac_momenta(n, T_over_m, seed):
    """q/m of n particles drawn from the relativistic Fermi-Dirac distribution q^2 / (e^q + 1) (isotropic), the distribution
    applyMomentumDistribution samples (ic_basic.hpp:1460-1585): Gamma(3) proposals accepted with probability 1 / (1 + e^-q)"""
    rng = np.random.default_rng(seed)
    q = np.empty(n)
    todo = np.arange(n)
    while len(todo):
        prop = rng.gamma(3.0, 1.0, len(todo))
        ok = rng.random(len(todo)) < 1.0 / (1.0 + np.exp(-prop))
        q[todo[ok]] = prop[ok]
        todo = todo[~ok]
    mu = 2.0 * rng.random(n) - 1.0
    ph = 2.0 * np.pi * rng.random(n)
    st = np.sqrt(1.0 - mu * mu)
    q *= T_over_m
    return np.stack([q * st * np.cos(ph), q * st * np.sin(ph), q * mu], axis=1)


def analytic_field(N, z0, nzl, rms, seed, ncomp=1):
    """smooth periodic field from a few plane waves (cheap at 512^3), local slab [ncomp][nzl][N][N]"""
    rng = np.random.default_rng(seed)
    x = np.arange(N) / N
    z = np.arange(z0, z0 + nzl) / N
    out = np.zeros((ncomp, nzl, N, N))
    for c in range(ncomp):
        for _ in range(6):
            k = rng.integers(1, 5, 3)
            ph = rng.uniform(0, 2 * np.pi, 3)
            amp = rng.standard_normal() / np.sqrt(float(k @ k))
            out[c] += amp * (np.sin(2 * np.pi * k[2] * z + ph[2])[:, None, None] * np.sin(2 * np.pi * k[1] * x + ph[1])[None, :, None]
                             * np.sin(2 * np.pi * k[0] * x + ph[0])[None, None, :])
        out[c] *= rms / max(out[c].std(), 1e-300)
    return out


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines, self.t0 = device, None, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        """the timed region starts now: only samples taken from here on count (the sampler itself is started during the warm-up, so
        that a short timed region still sees one)"""
        self.t0 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for t, ln in self.lines if self.t0 is None or t >= self.t0]
        where = "timed region"
        if not inside and self.lines:                               # region shorter than the sampling period: the last sample before it
            inside, where = [self.lines[-1][1]], "last sample of the warm-up (timed region shorter than the 200 ms sampling period)"
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "sampled": where}


# ----------------------------------------------------------------------------- roofline bookkeeping
def algorithmic_bytes(N, nzl, np_local, P):
    """SURVEY.md section 8(d) per-unit figures x local units, per launch of each entry point"""
    V = nzl * N * N
    Vk = (N // 2 + 1) * N * (N // P)
    return {
        "projection_T00_Tij_project": 48 * np_local + 64 * V,      # particles 48 r; source 8 w, Sij 48 w, phi 8 r
        "projection_T00_project": 48 * np_local + 16 * V,
        "projection_Tij_project": 48 * np_local + 56 * V,
        "projection_T0i_project": 48 * np_local + 32 * V,
        "kick_drift": 96 * np_local + 40 * V,                       # 48 r + 48 w per particle; phi, chi, B x3 read once
        "updateVel": 72 * np_local + 40 * V,
        "moveParticles": 72 * np_local + 40 * V,
        "rebin_sort": 112 * np_local,                               # 56 r + 56 w per particle (out-of-place re-bin)
        "prepareFTsource_scalar": 32 * V,
        "prepareFTsource_tensor": 104 * V,
        "solveModifiedPoissonFT": 32 * Vk,
        "projectFTscalar": 112 * Vk,
        "evolveFTvector": 192 * Vk,
        "projectFTscalar_evolveFTvector": 208 * Vk,                  # 6 S read once, chi written, B_i read + written
        "projectFTvector": 96 * Vk,
        "projectFTtensor": 192 * Vk,
        "fft_forward": 16 * V, "fft_backward": 16 * V,              # ideal: 8 r + 8 w per real site and component
        "projection_init": 8 * V, "field_sum": 8 * V,
    }


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """per-launch DRAM traffic of the dominant kernel from the committed ncu capture, if any"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference_run(ngrid, steps, warmup):
    """the reference's own gevolution.hpp main-loop cycle on the host cores (oracle/_ref)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import common
    import oracle
    ref = oracle.load_ref()
    kind = "reference"
    if ref is None:
        raise RuntimeError("oracle/_ref/libgevref.so missing (it is built in the build container and travels with the snapshot)")
    cosmo, ds = common.shipped_cosmology(), common.shipped_settings()
    a0 = 1.0 / (1.0 + ds[3])
    ids, pos, vel = local_particles(ngrid, 0, ngrid, a0, 42)
    sim = ref.sim(ngrid, 1, 0, ds, cosmo)
    sim.set_particles(0, ids, pos, vel, (cosmo[0] + cosmo[1]) / len(ids))
    sim.set_field("phi", analytic_field(ngrid, 0, ngrid, 1e-5, 1))
    sim.set_field("chi", analytic_field(ngrid, 0, ngrid, 1e-7, 2))
    for _ in range(warmup):
        sim.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.step()
    dt = time.perf_counter() - t0
    timers = sim.timers()
    sim.close()
    return dict(seconds=dt, particles=len(ids), kind=kind, timers=timers, description=ref.description)


def reference_replicas(ngrid, steps, warmup, nproc):
    """nproc independent single-threaded replicas of the reference cycle, one per host core, started together: the
    throughput the reference's own MPI decomposition could reach at best on these cores (its MPI build cannot be
    produced here: mpic++ / LATfield2 / FFTW are absent)"""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--replica", "--steps", str(steps), "--warmup", str(warmup), "--ngrid-ref", str(ngrid)]
    t0 = time.perf_counter()
    procs = [subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for _ in range(nproc)]
    outs = [p.communicate()[0] for p in procs]
    wall = time.perf_counter() - t0
    res = [json.loads(o.strip().splitlines()[-1]) for o, p in zip(outs, procs) if p.returncode == 0 and o.strip()]
    if not res:
        raise RuntimeError("no reference replica finished")
    seconds = max(r["seconds"] for r in res)                 # all replicas run concurrently: the slowest one bounds the rate
    return dict(seconds=seconds, particles=sum(r["particles"] for r in res), replicas=len(res), wall=wall, kind=res[0]["kind"])


def run_reference(args, rank):
    if args.replica:
        r = cpu_reference_run(args.ngrid_ref, args.steps, args.warmup)
        print(json.dumps({"seconds": r["seconds"], "particles": r["particles"], "kind": r["kind"]}), flush=True)
        return
    if rank != 0:
        return
    ngrid = env_int("GEVB_REF_NGRID", 128)
    cores = env_int("GEVB_REF_PROCS", os.cpu_count() or 1)
    r = reference_replicas(ngrid, args.steps, args.warmup, cores)
    value = r["particles"] * args.steps / r["seconds"]
    sample = (f"{args.steps} cycles of the reference main loop (gevolution.hpp compiled against the single-rank LATfield2 shim; "
              f"MPI/LATfield2/FFTW absent) at {ngrid}^3 grid / {ngrid}^3 particles, same GR parabolic path; {r['replicas']} independent "
              f"single-threaded replicas run concurrently, one per host core (upper bound of what the MPI build could reach on these cores)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["seconds"] / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "GR 512^3 grid / 512^3 particles (parabolic B); CPU arm runs a bounded sample", "sample_ngrid": ngrid, "replicas": r["replicas"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["replicas"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def clustered_particles_local(N, a, seed):
    """z ~ 0 like occupancy at full size: 70 % of the particles in Gaussian blobs (sigma 4..40 cells), 30 % Poisson floor"""
    rng = np.random.default_rng(seed)
    n = N ** 3
    nblob = 4096
    centres = rng.random((nblob, 3))
    widths = 10 ** rng.uniform(np.log10(4.0 / N), np.log10(40.0 / N), nblob)
    which = rng.integers(0, nblob, n)
    pos = rng.standard_normal((n, 3))
    pos *= widths[which, None]
    pos += centres[which]
    floor = rng.random(n) < 0.3
    pos[floor] = rng.random((int(floor.sum()), 3))
    pos -= np.floor(pos)
    pos[pos >= 1.0] = 0.0
    vel = rng.standard_normal((n, 3))
    vel *= 1e-3 * a
    return np.arange(n, dtype=np.int64), pos, vel


def particle_regime(gevb, ctx, common, N, label, ds, cosmo, mass, phi, chi):
    """deposit / kick+drift / re-bin times of 2 cycles on a fresh particle state (single rank)"""
    a0 = 1.0 / (1.0 + ds[3])
    if label == "lattice":
        ids, pos, vel = local_particles(N, 0, N, a0, 43)
        vel *= 0.03                                           # cold (30 km/s): positions stay at one particle per cell
        what = "one particle per cell (sigma 0.05 cell), q ~ N(0,(3e-5 a)^2): occupancy stays exactly 1"
    else:
        ids, pos, vel = clustered_particles_local(N, a0, 44)
        what = "70 % of the particles in 4096 Gaussian blobs (sigma 4..40 cells) + 30 % Poisson floor"
    sim = gevb.Sim(ctx, 1, 0, ds, cosmo)
    sim.set_particles(0, ids, pos, vel, mass)
    sim.set_field("phi", phi); sim.set_field("chi", chi)
    counts = sim.pcls(0).cell_counts()
    occ = {"max_per_cell": int(counts.max()), "empty_cells_frac": float((counts == 0).mean()), "particles_in_shared_cells_frac": float(counts[counts > 1].sum() / counts.sum())}
    del counts
    sim.step()
    ctx.sync(); ctx.timing(True); ctx.timing_read()
    nst = 2
    for _ in range(nst):
        sim.step()
    per = ctx.timing_read()
    ctx.timing(False)
    sim.close()
    out = {"what": what, "occupancy": occ}
    for k in ("projection_T00_Tij_project", "kick_drift", "rebin_sort"):
        if k in per:
            out[k + "_ms"] = per[k][0] / nst
    return out


# ----------------------------------------------------------------------------- multi-rank parity (checker use of oracle/)
def multi_rank_parity(gevb, dist, rank, world, local_rank):
    """The decomposition-versus-oracle checks of tests/multigpu_worker.py (deposit + fold, halo, kick, drift with slab
    migration, whole cycles of the time loop) at N = 32 on the ranks of this run, before anything is timed: the scaling lease
    is the only multi-GPU box the driver runs, so the parity of the slab-decomposed path is established there.  oracle/ is
    used here strictly as the checker (tier rule 3); a failure aborts the run."""
    import multigpu_worker as mw
    N = 32
    if N % world != 0 or N // world < 2:
        return {"skipped": f"N = {N} does not divide into {world} slabs"}
    chk = None
    if rank == 0:
        import oracle
        chk = oracle.load_ref() or oracle.load_ora()
        if chk is None:
            raise RuntimeError("multi-rank parity: no CPU checker library (oracle/_ref or oracle/libgev_oracle.so) in the snapshot")
    box = [gevb.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx = gevb.Context(N, device=local_rank, rank=rank, nranks=world, nccl_id=box[0])
    failures = []
    mw.RESULTS.clear()
    mw.QUIET = True
    mw.case_fft(ctx, N, failures)
    mw.case_particles(ctx, chk, N, failures)
    mw.case_time_loop(ctx, chk, N, 0, 1, failures, nsteps=2)
    mw.case_time_loop(ctx, chk, N, 1, 1, failures, nsteps=1)
    ctx.close()
    allf = [None] * world
    dist.all_gather_object(allf, failures)
    bad = [f for fl in allf for f in fl]
    out = None
    if rank == 0:
        worst = max(mw.RESULTS, key=lambda r: r[1] / r[2] if r[2] > 0 else (float("inf") if r[1] > 0 else 0.0))
        cells = [r for r in mw.RESULTS if "cells" in r[0] or "counts" in r[0] or "ids" in r[0] or "conservation" in r[0]]
        out = {"checker": chk.description, "ngrid": N, "ranks": world, "checks": len(mw.RESULTS), "failed": len(bad),
               "worst_field": {"name": worst[0], "error": worst[1], "tolerance": worst[2]},
               "cells_mismatch": int(sum(r[1] for r in cells)), "integer_checks": len(cells)}
    if bad:
        raise RuntimeError(f"multi-rank parity check failed: {bad}")
    return out


def bind_near_gpu(local_rank):
    """Bind this rank's host threads (and so the first-touch placement of its pinned buffers) to the cores of the NUMA
    node its GPU hangs off -- what `mpirun --bind-to` / `numactl` do for the reference's ranks.  Read from sysfs through the
    GPU's PCI address; a no-op when the box does not expose it (single node, VM without topology)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(path + "/numa_node").read())
        cpus = set()
        for part in open(path + "/local_cpulist").read().strip().split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus or cpus == os.sched_getaffinity(0):
            return {"numa_node": node, "cpus": len(os.sched_getaffinity(0)), "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus), "bound": True}
    except Exception as e:                                     # placement is an optimisation, never a failure
        return {"bound": False, "why": str(e)[:80]}


# ----------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch                      # first: its bundled NCCL must be the one mapped into the process
    import torch.distributed as dist
    import common
    import gevb

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    binding = bind_near_gpu(local_rank) if env_int("GEVB_BIND", 1) else {"bound": False}
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    parity = None
    if world > 1 and not args.no_parity:
        parity = multi_rank_parity(gevb, dist, rank, world, local_rank)
    if world > 1:
        box = [gevb.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]

    cfg = CONFIGS[args.config]
    N = args.ngrid if args.ngrid else cfg["ngrid"]
    species = cfg["species"]
    ctx = gevb.Context(N, device=local_rank, rank=rank, nranks=world, nccl_id=nccl_id)
    cosmo, ds = common.shipped_cosmology(), common.shipped_settings()
    ncdm = None
    if "ncdm0" in species:
        cosmo, m_ncdm, T_ncdm, Om_ncdm = common.ncdm_model(cosmo)
        ncdm = (m_ncdm, T_ncdm, Om_ncdm)
    a0 = 1.0 / (1.0 + ds[3])
    t_setup = time.perf_counter()
    sim = gevb.Sim(ctx, 1, cfg["vector_flag"], ds, cosmo)
    if ncdm:
        # deposits of the ncdm species on from the start; move limit as the parser's default (Ngrid, clamped to the slab by the library)
        sim.set_ncdm(m_ncdm, T_ncdm, Om_ncdm, [1e4, 1e4], [1e4, 1e4], 1e4, float(N))
    np_local_total, masses = 0, {}
    for sp_index, name in enumerate(species):
        slot = {"cdm": 0, "b": 1, "ncdm0": 2, "ncdm1": 3}[name]
        ids, pos, vel = local_particles(N, ctx.z0, ctx.nzl, a0, 42 + 17 * sp_index)
        if name.startswith("ncdm"):
            k = int(name[-1])
            # q/m scale of the species: (Omega_g h^2 / C_PLANCK_LAW)^(1/4) T_ncdm k_B / m_ncdm  (ic_basic.hpp:2218)
            T_over_m = (cosmo[7] * cosmo[10] ** 2 / 4.48147e-7) ** 0.25 * T_ncdm[k] * 8.61733e-5 / m_ncdm[k]
            vel = fermi_dirac_momenta(len(ids), T_over_m, 4242 + 1000 * ctx.z0 + k)
            mass = Om_ncdm[k] / N ** 3
        else:
            mass = {"cdm": cosmo[0] if "b" in species else cosmo[0] + cosmo[1], "b": cosmo[1]}[name] / N ** 3
        masses[slot] = mass
        sim.set_particles(slot, ids, pos, vel, mass)
        np_local_total += len(ids)
        if sp_index == 0:
            np_local = len(ids)
        del ids, pos, vel
    if ncdm:
        # the first cycle's sub-stepping needs the largest ncdm velocity (the IC generator returns it, main.cpp:330-340)
        vmax = []
        for k in range(2):
            _, _, v = sim.pcls(2 + k).download()
            q = float(np.sqrt((v * v).sum(axis=1).max())) if len(v) else 0.0
            vmax.append(q / a0 / np.sqrt((q / a0) ** 2 + 1.0))
            del v
        vmax = ctx.parallel_max(vmax)
        sim.set_ncdm_maxvel(vmax)
    np_total = N ** 3 * len(species)
    mass = masses[0]
    phi = analytic_field(N, ctx.z0, ctx.nzl, 1e-5, 1)
    chi = analytic_field(N, ctx.z0, ctx.nzl, 1e-7, 2)
    sim.set_field("phi", phi)
    sim.set_field("chi", chi)
    t_setup = time.perf_counter() - t_setup

    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    clocks = ClockSampler(local_rank)
    for w in range(args.warmup):
        if rank == 0 and w == args.warmup - 1:
            clocks.start()                                      # streaming before the timed region begins
        sim.step()
    ctx.sync()
    ctx.timing(True)
    ctx.timing_read()
    launches0 = ctx.launches
    if rank == 0 and clocks.proc is None:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); ctx.sync(); torch.cuda.synchronize()
    clocks.mark()
    ev0.record(stream)
    for _ in range(args.steps):
        sim.step()
    ev1.record(stream)
    ctx.sync(); torch.cuda.synchronize(); barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    launches = ctx.launches - launches0
    per_class = ctx.timing_read()
    ctx.timing(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = np_total * args.steps / (ms * 1e-3)
    state = sim.state()
    # size-independent invariants of the timed run (full benchmark size): no particle lost in re-binning / migration,
    # mass conservation of the deposit (T00hom = Omega_m up to the O(phi) metric and O(q^2) kinetic corrections)
    n_now = torch.tensor([float(sum(sim.pcls({"cdm": 0, "b": 1, "ncdm0": 2, "ncdm1": 3}[nm]).count() for nm in species))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(n_now)
    Omega_dep = cosmo[0] + cosmo[1] + (float(np.sum(ncdm[2])) if ncdm else 0.0)
    invariants = {"particles_after": int(n_now.item()), "particles_conserved": int(n_now.item()) == np_total,
                  "T00hom_over_Omega_m_minus_1": state["T00hom"] / Omega_dep - 1.0}
    # the relativistic ncdm species carry kinetic energy: T00hom exceeds Omega_m by <q^2>/2a^2-ish for them (a few per cent of Omega_ncdm)
    tol_T00 = 1e-3 if not ncdm else 2e-2
    if not invariants["particles_conserved"] or abs(invariants["T00hom_over_Omega_m_minus_1"]) > tol_T00:
        raise RuntimeError(f"benchmark run violates an invariant: {invariants}")
    mem_gb = torch.cuda.mem_get_info(local_rank)
    mem_used_gb = (mem_gb[1] - mem_gb[0]) / 2 ** 30
    ncdm_steps = None
    if ncdm:
        ncdm_steps = [int(x) for x in sim.ncdm_state()[1][:2]]

    # ---- config 5: the hij spectrum call of output.hpp:1961-1981 once, outside the timed region (Tij deposit, 6 FFTs, TT projection, binning)
    hij_ms = None
    if cfg["hij"]:
        import tempfile
        tmpd = tempfile.mkdtemp()
        ctx.sync(); barrier()
        t0 = time.perf_counter()
        sim.write_spectra(os.path.join(tmpd, "pk"), 0, 1024, 128)
        ctx.sync(); barrier()
        hij_ms = 1e3 * (time.perf_counter() - t0)
        if rank == 0:
            pk = np.loadtxt(os.path.join(tmpd, "pk000_hij.dat"))
            invariants["hij_spectrum_bins"] = int(len(pk))
            invariants["hij_spectrum_finite"] = bool(np.all(np.isfinite(pk)))

    # ---- the FFT exchange by itself: two more cycles with the component pipeline off, so that the time of the pushes is
    #      not hidden behind the local transforms (the timed run above keeps the overlap on)
    exchange_ms = None
    if world > 1:
        gevb.tuning("fft_overlap", 0)
        sim.step()
        ctx.sync(); ctx.timing(True); ctx.timing_read()
        for _ in range(2):
            sim.step()
        per_x = ctx.timing_read(); ctx.timing(False)
        gevb.tuning("fft_overlap", 2)
        if "fft_alltoall" in per_x:
            exchange_ms = per_x["fft_alltoall"][0] / 2

    # ---- optional ablation of kernel variants (stderr; not part of the bench line) ------------------------------
    if args.ablate:
        for spec in args.ablate.split(","):
            knob, vals = spec.split("=")
            for v in vals.split(":"):
                gevb.tuning(knob, int(v))
                sim.step()
                ctx.sync(); ctx.timing(True); ctx.timing_read()
                for _ in range(2):
                    sim.step()
                per = ctx.timing_read(); ctx.timing(False)
                if rank == 0:
                    print(json.dumps({"ablate": knob, "value": int(v), "ms": {k: round(t / 2, 3) for k, (t, n) in per.items() if t / 2 > 0.1}}), file=sys.stderr, flush=True)
            gevb.tuning(knob, int(vals.split(":")[0]))

    # ---- end to end through the C ABI with host buffers (species 0 and the metric state; config 3 is the quoted one) ----
    e2e_value, h2d, d2h, e2e_parts = None, 0, 0, None
    e2e_steps = max(1, min(args.steps, env_int("GEVB_E2E_STEPS", 2)))
    if len(species) == 1 and not args.no_e2e:
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
        gid, gpos, gvel = sim.pcls(0).download()
        h_phi, h_chi, h_bft = pin(sim.get_field("phi")), pin(sim.get_field("chi")), pin(sim.get_field("BiFT"))
        # slab counts change as particles migrate; host buffers are sized once with head-room
        cap = len(gid) if world == 1 else int(np_local * 1.1) + 1024
        h_id, h_pos, h_vel = pin(np.resize(gid, cap)), pin(np.resize(gpos, (cap, 3))), pin(np.resize(gvel, (cap, 3)))
        nloc = [len(gid)]
        del gid, gpos, gvel
        fbytes = h_phi.nbytes + h_chi.nbytes + h_bft.nbytes

        e2e_parts = {}

        def e2e_step(parts=None):
            def mark(name, t=[0.0]):                            # untimed warm-up only: where an end-to-end step spends its time
                if parts is not None:
                    ctx.sync(); now = time.perf_counter()
                    if name: parts[name] = round((now - t[0]) * 1e3, 2)
                    t[0] = now
            mark(None)
            n = nloc[0]
            sim.set_particles(0, h_id[:n], h_pos[:n], h_vel[:n], mass)          # host -> device: particle state
            mark("h2d_particles_ms")
            sim.set_field("phi", h_phi); sim.set_field("chi", h_chi); sim.set_field("BiFT", h_bft)
            mark("h2d_fields_ms")
            sim.step()
            mark("step_ms")
            p = sim.pcls(0)
            m = p.count()
            L = gevb.lib()
            gevb._ck(L.gevb_pcls_download(p.h, gevb._ptr(h_id[:m]), gevb._ptr(h_pos[:m]), gevb._ptr(h_vel[:m])), "download")
            mark("d2h_particles_ms")
            for name, buf in (("phi", h_phi), ("chi", h_chi), ("BiFT", h_bft)):
                gevb._ck(L.gevb_field_download(sim.field(name).h, gevb._ptr(buf)), "download")
            mark("d2h_fields_ms")
            nloc[0] = m
            return 56 * n + fbytes, 56 * m + fbytes

        e2e_step()                                              # warm-up (allocations)
        e2e_step(e2e_parts)                                     # second warm-up, with a synchronisation after every part
        ctx.sync(); barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            up, down = e2e_step()
            h2d += up; d2h += down
        ctx.sync(); barrier()
        e2e_s = time.perf_counter() - t0
        h2d //= e2e_steps; d2h //= e2e_steps
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
            hb = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
            dist.all_reduce(hb)
            h2d, d2h = int(hb[0].item()), int(hb[1].item())
        e2e_value = np_total * e2e_steps / e2e_s

    # ---- roofline of the dominant own kernel ---------------------------------------------------
    peak, peak_src = measured_peak()
    bytes_per = algorithmic_bytes(N, ctx.nzl, np_local, world)
    kernels = {}
    NESTED = ("fft_alltoall", "fft_transpose")                 # timed inside fft_forward / fft_backward
    total_ms = sum(v[0] for k, v in per_class.items() if k not in NESTED) or 1.0
    for name, (kms, cnt) in per_class.items():
        b = bytes_per.get(name)
        ncomp_calls = cnt
        if name in ("fft_forward", "fft_backward"):
            # one call transforms all components of its field: bytes are counted per call below
            ncomp_calls = None
        entry = {"ms_per_step": kms / args.steps, "calls_per_step": cnt / args.steps, "share": kms / total_ms}
        if b is not None and ncomp_calls is not None:
            gbs = b * cnt / (kms * 1e-3) / 1e9
            entry.update({"bytes_per_launch": b, "achieved_gbs": gbs, "frac": gbs / peak})
        kernels[name] = entry
    # FFT: 12 component transforms per step (parabolic): 1 + 6 forward, 1 + 1 + 3 backward; the elliptic method adds 3 forward
    nfwd = 7 + (3 if cfg["vector_flag"] else 0)
    for name, ncomp_per_step in (("fft_forward", nfwd), ("fft_backward", 5)):
        if name in kernels and per_class[name][0] > 0:
            gbs = 16 * ctx.nzl * N * N * ncomp_per_step * args.steps / (per_class[name][0] * 1e-3) / 1e9
            kernels[name].update({"bytes_per_launch": 16 * ctx.nzl * N * N, "achieved_gbs": gbs, "frac": gbs / peak, "note": "cuFFT; ideal 16 B per site and component"})
    nvlink = None
    if world > 1 and "fft_alltoall" in per_class and per_class["fft_alltoall"][0] > 0:
        # bytes one rank puts on NVLink per step: component transforms x 16 B x local k-sites x (P-1)/P
        sent = (nfwd + 5) * 16 * (N // 2 + 1) * N * (N // world) * (world - 1) / world
        exposed_ms = per_class["fft_alltoall"][0] / args.steps
        a2a_ms = exchange_ms if exchange_ms else exposed_ms
        nvlink = {"bound": "nvlink", "kernel": "k_push_fwd / k_push_bwd (FFT transposes stored straight into peer memory over NVLink) + barrier", "achieved": sent / (a2a_ms * 1e-3) / 1e9, "peak": 770.0, "unit": "GB/s per direction per GPU",
                  "frac": sent / (a2a_ms * 1e-3) / 1e9 / 770.0, "peak_source": "measured peer copy, B200_PROFILING.md (900 nominal)", "bytes_sent_per_rank_per_step": int(sent), "ms_per_step": a2a_ms,
                  "exposed_ms_per_step_in_timed_run": exposed_ms, "note": "ms_per_step: exchange measured with the component pipeline off; exposed: what the main stream still waited for in the timed run (pushes overlap the local transforms)"}
    own = {k: v for k, v in kernels.items() if not k.startswith("fft_") and "frac" in v}
    top = max(own, key=lambda k: own[k]["ms_per_step"]) if own else None
    roofline = None
    if top:
        # DRAM traffic of that kernel from the committed ncu --set full capture of the SAME regime (the timed run's: synthetic ICs
        # evolved by >= 10 cycles) and configuration (512^3, one species); scaled by this rank's share of the work, else null
        traffic, traffic_src = None, None
        tj = ncu_traffic().get(top)
        if isinstance(tj, dict) and args.config == 3 and N == 512:
            traffic = tj["dram_bytes_per_launch"] * np_local / float(N ** 3)
            traffic_src = tj.get("capture")
        roofline = {"kernel": top, "bound": "hbm", "achieved": own[top]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": own[top]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "ms_per_launch": own[top]["ms_per_step"] / max(own[top]["calls_per_step"], 1e-9), "share_of_step": own[top]["share"]}

    # ---- the particle kernels in the other occupancy regimes (report lattice-like and clustered) ---
    regimes = None
    if world == 1 and not args.no_regimes and len(species) == 1:
        regimes = {"timed_run": "quasi-uniform ICs evolved by warmup+steps cycles (hot synthetic velocities: cell occupancy drifts from exactly 1 towards Poisson)"}
        for label in ("lattice", "clustered"):
            regimes[label] = particle_regime(gevb, ctx, common, N, label, ds, cosmo, mass, phi, chi)

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) -----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ngrid_cpu, nst = env_int("GEVB_REF_NGRID", 128), env_int("GEVB_REF_STEPS", 6)
        r = reference_replicas(ngrid_cpu, nst, 1, env_int("GEVB_REF_PROCS", os.cpu_count() or 1))
        cpu = {"value": r["particles"] * nst / r["seconds"], "unit": UNIT, "cores": r["replicas"], "kind": r["kind"],
               "sample": f"{nst} cycles of the reference main loop (gevolution.hpp over the single-rank LATfield2 shim; the reference's MPI build cannot be produced: mpic++ / LATfield2 / FFTW absent) at {ngrid_cpu}^3 grid / {ngrid_cpu}^3 "
                         f"particles in {r['replicas']} independent single-threaded replicas run concurrently, one per host core"}

    if rank == 0:
        line = {
            "metric": METRIC if args.config == 3 and N == 512 else f"particle-steps/s (config {args.config}, {N}^3)", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BASELINE config {args.config}: {cfg['what']}; one cycle of main.cpp:372-879 per step (fused T00+Tij deposit, fused kick+drift, {nfwd + 5} FFTs)",
                       "ngrid": N, "particles": np_total, "species": list(species), "slabs": world, "parallelism": f"z-slab x{world}",
                       "l2_policy": "inputs larger than L2: every pass streams >= 1 GB per rank (126 MB L2)",
                       "e2e_steps": e2e_steps, "z": 1.0 / state["a"] - 1.0, "setup_s": t_setup, "device_memory_used_gb_rank0": mem_used_gb, "host_binding_rank0": binding, "e2e_parts_rank0": e2e_parts,
                       "ncdm_substeps": ncdm_steps, "hij_spectrum_call_ms": hij_ms},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "invariants": invariants,
            "parity_check": parity,
            "roofline": roofline,
            "nvlink": nvlink,
            "regimes": regimes,
            "cpu_baseline": cpu,
            "kernels": kernels,
        }
        print(json.dumps(line), flush=True)
    sim.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ngrid", type=int, default=env_int("GEVB_BENCH_NGRID", 0), help="override the configuration's lattice size (parity / smoke runs)")
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS), help="BASELINE.json configuration (SURVEY 8d numbering); the bench line is quoted on 3")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-rank parity check that precedes the timed region when WORLD_SIZE > 1")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replica", action="store_true", help="(internal) one replica of the CPU reference arm")
    ap.add_argument("--ngrid-ref", type=int, default=128)
    ap.add_argument("--ablate", default="", help="e.g. geodesic_variant=6:1:0,fft_decomposed=1:0 -- times 2 cycles per setting (stderr), first value is restored")
    ap.add_argument("--no-regimes", action="store_true", help="skip the extra lattice / clustered measurements of the particle kernels")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if args.gpus == 1 and world == 1:
            pass
        elif world == 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus} (WORLD_SIZE is 1)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

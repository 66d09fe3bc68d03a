#!/usr/bin/env python
"""bench.py -- particle-updates/s of the per-chunk PIC step (push + deposit + sort + halo + migration).

    python bench.py --gpus N --steps K --warmup W            # the B200 path (libnixb200.so)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle/_ref)

Workload (BASELINE.json configs[1]): 128^3 cells per GPU, 128 particles per cell (electrons + ions,
64 each), 2nd-order shape, fp64, periodic thermal plasma, as 8^3 chunks of 16^3 cells (the
reference's int-addressed chunk serialisation caps one chunk at < 2 GiB).
One step = clear J, gather + Boris push + move + Esirkepov deposit, J halo, E/B halo, particle
migration and per-cell count/sort, for every chunk and species.

Prints ONE JSON line (rank 0).  See DESIGN.md section 6 for how every field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries ONE JSON line: everything else that writes to file descriptor 1 (NCCL's version banner
# comes from C code) is sent to stderr; emit() writes the line to the real stdout
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-updates/s (push+deposit+sort)"
# Algorithmic bytes per particle (SURVEY.md 8(d) counts 116 B for a fused push+deposit+count):
#   k_push     reads x y z ux uy uz (48 B), writes them back (48 B), the old position to the temporary
#              array (24 B, the reference's xv[0:3] = xu[0:3]) and the sort key (4 B)
#   k_deposit  reads the old and the new position (48 B)
ALGO_BYTES = {"k_push": 124.0, "k_deposit": 48.0}


CONFIGS = {
    # BASELINE.json configs[1] -- the contract line
    "cfg2": dict(cdims=(8, 8, 8), dims=(16, 16, 16), order=2, ppc=64, ns=2, vth=(0.1, 0.02), drift=None,
                 label="128^3 cells per GPU, 128 ppc (electrons+ions, 64 each), order 2, fp64, periodic thermal plasma"),
    # configs[2]: 256^3 on one GPU as 512 chunks of 32^3 in the reference's Gilbert order (ppc is not given by
    # BASELINE.json; SURVEY.md 8d: 16 per species, 2 species)
    "cfg3": dict(cdims=(8, 8, 8), dims=(32, 32, 32), order=2, ppc=16, ns=2, vth=(0.1, 0.02), drift=None,
                 label="256^3 cells per GPU as 512 chunks of 32^3 (Gilbert order), 32 ppc (2 species x 16), order 2, fp64"),
    # configs[3]: Weibel set-up, two counter-streaming electron species u_z = +-0.5 c, order 3, nb 3; 512^3 over
    # 8 GPUs = 256^3 per GPU.  24 per species instead of 32: 32 needs 186 GB with the double-buffered store
    "cfg4": dict(cdims=(8, 8, 8), dims=(32, 32, 32), order=3, ppc=24, ns=2, vth=(0.03, 0.03), drift=(0.5, -0.5),
                 q=(-1.0, -1.0), m=(1.0, 1.0),
                 label="Weibel: 256^3 cells per GPU as 512 chunks of 32^3, two counter-streaming electron species "
                       "u_z = +-0.5 c, 48 ppc (2 x 24), order 3, nb 3, fp64"),
    # configs[4]: weak scaling with a non-uniform density n(x) = 1 + 0.8 sin(2 pi x / Lx) and SFC rebalancing
    # (Balancer::assign, balancer.cpp:8-72,126-132; chunks shipped GPU to GPU by nixb200_domain_rebalance).
    # 16 ppc on average instead of 32: the rank in the dense half starts with 1.5x the mean and a rebalance
    # holds the old and the new arrays at once
    "cfg5": dict(cdims=(8, 8, 8), dims=(32, 32, 32), order=2, ppc=8, ns=2, vth=(0.1, 0.02), drift=None, sine=0.8,
                 label="256^3 cells per GPU as 512 chunks of 32^3, density 1 + 0.8 sin(2 pi x / Lx), 16 ppc on average "
                       "(2 species x 8), order 2, fp64, rank boundaries rebalanced along the Gilbert curve"),
}


def workload(args):
    w = dict(CONFIGS[args.config])
    small = os.environ.get("NIXB200_BENCH_SMALL", "") == "1" or args.small
    if small:
        w["cdims"] = (2, 2, 2)
    if args.cdims:
        w["cdims"] = tuple(int(v) for v in args.cdims.split(","))
    if args.chunk:
        w["dims"] = tuple(int(v) for v in args.chunk.split(","))
    if args.order:
        w["order"] = args.order
    if args.ppc:
        w["ppc"] = args.ppc
    w["name"] = args.config
    return w


def make_problem(w, seed=2024, cdims=None, coord=None):
    from nix_b200.synth import Problem
    return Problem(cdims or w["cdims"], w["dims"], w["order"], ppc=w["ppc"], ns=w["ns"], seed=seed,
                   vth=w["vth"], coord=coord, q=w.get("q", (-1.0, 1.0)), m=w.get("m", (1.0, 25.0)))


def global_box(cd, world, first_axis=0):
    """Weak scaling: the per-GPU block of cd chunks is repeated world times (z first, then y, then x).  Chunk
    ids follow the reference's Gilbert curve over the WHOLE box (sfc.cpp:97-169 via nix_b200/sfc.py) and every
    rank owns one contiguous segment of it (Balancer::assign_initial with uniform loads, balancer.cpp:101-124)."""
    from nix_b200.sfc import chunk_coords
    rep = [1, 1, 1]
    a, n = 0, world
    while n > 1:
        if n % 2:
            raise SystemExit("bench.py: --gpus must be a power of two")
        rep[(first_axis + a) % 3 if first_axis == 0 else (first_axis - a) % 3] *= 2
        n //= 2
        a += 1
    gcd = tuple(cd[i] * rep[i] for i in range(3))
    return gcd, chunk_coords(gcd)


def chunk_counts(prob, w, ids):
    """particles per chunk and species: uniform, or following 1 + a sin(2 pi x / Lx) at the chunk's centre"""
    n = prob.ncell() * prob.ppc
    if not w.get("sine"):
        return np.full(len(ids), n, dtype=np.int64)
    xc = np.array([(prob.coord[k][2] + 0.5) / prob.cdims[2] for k in ids])
    return np.maximum(1, np.rint(n * (1.0 + w["sine"] * np.sin(2 * np.pi * xc)))).astype(np.int64)


def device_particles(torch, prob, w, ids, s, seed):
    """Synthetic particles of species s for the chunks `ids`, generated on the device (the large configs would
    spend minutes in numpy): uniform positions inside each chunk, Gaussian momenta, optional drift along z."""
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed * 1000 + s)
    dims = torch.tensor(prob.dims, dtype=torch.float64, device="cuda")
    counts = chunk_counts(prob, w, ids)
    first = np.concatenate([[0], np.cumsum(counts)])
    out = torch.empty((int(first[-1]), 7), dtype=torch.float64, device="cuda")
    for j, k in enumerate(ids):
        n = int(counts[j])
        o = out[int(first[j]):int(first[j + 1])]
        lo = torch.tensor([float(prob.coord[k][a] * prob.dims[a]) for a in range(3)], dtype=torch.float64, device="cuda")
        u = torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=gen) * (1.0 - 1e-12)
        o[:, 0] = lo[2] + u[:, 0] * dims[2]
        o[:, 1] = lo[1] + u[:, 1] * dims[1]
        o[:, 2] = lo[0] + u[:, 2] * dims[0]
        o[:, 3:6] = torch.randn((n, 3), dtype=torch.float64, device="cuda", generator=gen) * w["vth"][s]
        if w.get("drift"):
            o[:, 5] += w["drift"][s]
        idv = torch.arange(n, dtype=torch.int64, device="cuda") + (int(k) << 32) + (s << 56)
        o[:, 6] = idv.view(torch.float64)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

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
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(w, steps, warmup, budget_s=None):
    """Time the reference's own CPU implementation (oracle/_ref when it was built from the reference
    sources, else the plain-C port) on a bounded sample of the workload with all host threads."""
    from oracle import nixoracle as no
    name = no.best_timing_backend()
    lib = no.load(name)
    nthreads = os.cpu_count() or 1
    lib.nixo_set_num_threads(nthreads)
    nthreads = lib.nixo_get_num_threads()
    # bounded sample: same chunk shape / ppc / order, fewer chunks (>= 1 chunk per thread and phase)
    if nthreads <= 8:
        cd = (2, 2, 2)
    elif nthreads <= 32:
        cd = (2, 4, 4)
    elif nthreads <= 128:
        cd = (4, 4, 4)
    else:
        cd = (4, 4, 8)
    prob = make_problem(w, cdims=cd)
    dom = no.Domain(lib, prob.cdims, prob.dims, prob.nb, prob.order, prob.ns, prob.q, prob.m, prob.coord,
                    prob.ncell() * prob.ppc)
    for k, c in enumerate(dom.chunks):
        c.uf[...] = prob.field(k)
        for s in range(prob.ns):
            c.set_particles(s, prob.particles(k, s))
    dom.exchange(no.MODE_FIELD)
    dom.sort_only()
    simd = name.startswith("ref")  # the reference's vectorised sorted path (xsimd batches)
    npart = dom.total_particles()
    for _ in range(warmup):
        dom.step(0.5, 1.0, simd)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        dom.step(0.5, 1.0, simd)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    kind = "reference" if name.startswith("ref") else "port"
    # the reference's scalar templates on the same sample, one step (both paths reported)
    scalar = None
    if simd:
        t1 = time.perf_counter()
        dom.step(0.5, 1.0, False)
        scalar = npart / (time.perf_counter() - t1)
    return {
        "value": npart * done / dt, "unit": "particle-updates/s", "cores": nthreads, "kind": kind,
        "backend": name, "simd_lanes": lib.nixo_simd_lanes() if simd else 1, "scalar_path_value": scalar,
        "sample": f"{cd[0]}x{cd[1]}x{cd[2]} chunks of {'x'.join(map(str, w['dims']))} cells, {w['ppc']} ppc x {w['ns']} species "
                  f"({npart} particles), order {w['order']}, {done} steps after {warmup} warm-up",
        "ms_per_step": 1e3 * dt / done, "steps": done,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs on rank 0 only
    w = workload(args)
    r = cpu_reference_run(w, args.steps, args.warmup)
    out = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "particle-updates/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s (bounded CPU sample of the same chunk shape, ppc and order: %s)"
                               % (w["name"], w["label"], r["sample"])},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "backend", "simd_lanes",
                                           "scalar_path_value")},
        "e2e": {"value": r["value"], "unit": "particle-updates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def fp64_peak():
    """Measured fp64 FMA peak of this GPU (tools/micro/dfma_peak, a register-resident DFMA kernel)."""
    exe = os.path.join(ROOT, "tools", "micro", "dfma_peak")
    if not os.path.exists(exe):
        return None
    try:
        r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        return None


def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from nix_b200 import core

    w = workload(args)
    peak64 = fp64_peak() if rank == 0 else None  # before the timed regions, GPU otherwise idle
    # (cfg5: the box grows along x first, the axis the density varies along, so that already two ranks are imbalanced)
    gcd, gcoord = global_box(w["cdims"], world, first_axis=2 if w.get("sine") else 0)
    prob = make_problem(w, seed=2024, cdims=gcd, coord=gcoord)
    bd = core.uniform_boundary(prob.nchunk, world)
    ids = list(range(int(bd[rank]), int(bd[rank + 1])))
    stream = torch.cuda.current_stream()
    dom = core.Domain(prob.cdims, prob.dims, prob.nb, prob.order, prob.q, prob.m, coord=prob.coord, device=local,
                      id_range=(ids[0], ids[-1] + 1), strict_fp=bool(args.strict), capacity_factor=1.12,
                      stream=stream.cuda_stream)
    if world > 1:
        dom.set_ranks(bd, rank)
        dom.comm_init_torch()
    nchunk = dom.nchunk
    cells = int(np.prod(dom.M))
    # pinned host mirrors of the grid arrays (initial condition in, diagnostic snapshots out)
    uf_host = torch.empty((nchunk, cells, 6), dtype=torch.float64, pin_memory=True)
    icells = prob.ncell()
    ufi_host = torch.empty((nchunk, icells, 6), dtype=torch.float64, pin_memory=True)  # interior cells only
    uji_host = torch.empty((nchunk, icells, 4), dtype=torch.float64, pin_memory=True)
    ufn = uf_host.numpy()
    for k in range(nchunk):
        ufn[k] = prob.field(ids[k]).reshape(cells, 6)
    dom.field_upload_async(core.FIELD_UF, uf_host.data_ptr())
    dom.exchange_field()
    dom.field_download_async(core.FIELD_UF, uf_host.data_ptr())  # ghosts consistent on the host too
    dom.synchronize()
    nbh = prob.nb
    ufi_host.numpy()[...] = ufn.reshape((nchunk,) + tuple(prob.M) + (6,))[
        :, nbh:nbh + prob.dims[0], nbh:nbh + prob.dims[1], nbh:nbh + prob.dims[2]].reshape(nchunk, icells, 6)
    npc = prob.ncell() * prob.ppc

    def load_particles(dm):
        for s in range(prob.ns):
            if w["name"] == "cfg2":  # the contract line keeps the numpy inputs the parity tests use
                flat = np.empty((nchunk * npc, 7), dtype=np.float64)
                for k in range(nchunk):
                    flat[k * npc:(k + 1) * npc] = prob.particles(ids[k], s)
                dm.set_particles_flat(s, flat, np.full(nchunk, npc, dtype=np.int64))
                del flat
            else:
                t = device_particles(torch, prob, w, ids, s, seed=2024)
                torch.cuda.synchronize()
                dm.set_particles_ptr(s, t.data_ptr(), chunk_counts(prob, w, ids))
                del t
                torch.cuda.empty_cache()
        dm.sort()
        dm.synchronize()

    load_particles(dom)
    ntot = dom.total_particles()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dt = 0.5
    # coupling constant of the device-side field update: E -= cfj dt J.  With unit charges and ppc particles per
    # cell the electron plasma frequency is sqrt(cfj ppc); cfj = 0.16 / ppc puts it at 0.4 (wpe dt = 0.2)
    cfj = 0.16 / w["ppc"]
    for _ in range(max(args.warmup, 3)):
        dom.step(dt)
    barrier()
    err = dom.check()
    if err:
        raise SystemExit(f"bench.py: device error bits {err} during warm-up")

    rebal = None
    skip_e2e = False
    if w["name"] == "cfg5":
        args.no_fp32 = True

    def timed_steps(nsteps, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(nsteps):
            fn()
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    if w["name"] == "cfg5" and world > 1:
        # imbalanced start (uniform boundaries over a non-uniform density), then rounds of
        # Balancer::assign on the per-chunk particle counts + device-to-device chunk shipping
        from nix_b200 import balancer
        ms_before = timed_steps(args.steps, lambda: dom.step(dt))
        n_before = dom.total_particles()
        bnow = [int(v) for v in bd]
        rounds, moved = 0, 0
        tr0 = time.perf_counter()
        for _ in range(12):
            # Chunk::load (chunk.hpp:177-192) of a GPU chunk: particles + a per-cell term.  Measured on this
            # workload (profiles/r02p_cfg5_n2.json): a bin costs the kernels about as much as 12 particles
            # (per-bin reduction and flush of the deposit, the 8 keys per cell of the sort)
            mine = torch.zeros(prob.nchunk, dtype=torch.float64, device="cuda")
            mine[bnow[rank]:bnow[rank + 1]] = torch.tensor(
                sum(dom.get_np(s) for s in range(prob.ns)) + 12.0 * prob.ncell() * prob.ns, dtype=torch.float64)
            dist.all_reduce(mine)
            bnew = balancer.assign(mine.cpu().numpy(), bnow)
            if bnew == bnow:
                break
            moved += sum(abs(a - b) for a, b in zip(bnew, bnow))
            dom.rebalance(bnew, rank)
            bnow = bnew
            rounds += 1
        torch.cuda.synchronize()
        t_rebal = time.perf_counter() - tr0
        nchunk = dom.nchunk
        ids = list(range(bnow[rank], bnow[rank + 1]))
        for _ in range(2):
            dom.step(dt)
        nmax = torch.tensor([float(n_before), float(dom.total_particles())], dtype=torch.float64, device="cuda")
        nmin = nmax.clone()
        dist.all_reduce(nmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(nmin, op=dist.ReduceOp.MIN)
        tb = torch.tensor([ms_before], dtype=torch.float64, device="cuda")
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        rebal = {"rounds": rounds, "chunks_moved": int(moved), "seconds": t_rebal, "boundary": bnow,
                 "ms_per_step_before": float(tb[0]) / args.steps,
                 "particles_per_rank_before": [float(nmin[0]), float(nmax[0])],
                 "particles_per_rank_after": [float(nmin[1]), float(nmax[1])]}
        ntot = dom.total_particles()
        # the pinned host mirrors follow the new chunk count; they start from the fields the device holds
        ufi_host = torch.empty((nchunk, icells, 6), dtype=torch.float64, pin_memory=True)
        uji_host = torch.empty((nchunk, icells, 4), dtype=torch.float64, pin_memory=True)
        dom.interior_download_overlapped(core.FIELD_UF, ufi_host.data_ptr())
        dom.copy_synchronize()
        skip_e2e = True  # (= do not reset the host mirror from the initial condition)

    # ---- timed region: K device-resident steps ----
    try:
        gpu_sel = "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        gpu_sel = str(local)
    sampler = ClockSampler(gpu_sel)
    sampler.start()
    dom.set_profiling(True)
    dom.phase_ms()
    l0 = core.launch_count()
    ms = timed_steps(args.steps, lambda: dom.step(dt))
    launches = core.launch_count() - l0
    phases = dom.phase_ms()
    dom.set_profiling(False)
    clocks = sampler.stop()

    # ---- the same K steps in the OTHER arithmetic mode (strict = no FMA contraction = bit-exact with the
    #      reference's scalar templates; the headline is the contracted mode unless --strict) ----
    dom.set_strict_fp(not args.strict)
    dom.step(dt)
    ms_other = timed_steps(args.steps, lambda: dom.step(dt))
    dom.set_strict_fp(bool(args.strict))

    # ---- end to end through the C ABI with HOST buffers: one diagnostic interval of a nix application whose
    #      fields live on the device (field solver = nixb200_domain_step_em): the initial E/B of every chunk comes
    #      up from pinned host memory, every step reads back the history diagnostics (field energies and particle
    #      counts per chunk -- what HistoryDiag and the balancer consume), and at the end of the interval the
    #      interior J and E/B of every chunk go down as a snapshot.  Bytes are counted from what is copied. ----
    energies_t = torch.zeros((args.steps, nchunk, 2), dtype=torch.float64, pin_memory=True)
    counts_t = torch.zeros((args.steps, prob.ns, nchunk), dtype=torch.int64, pin_memory=True)
    energies = energies_t.numpy()
    # warm-up of this leg's own calls (first use loads the field-solver kernels and allocates the staging buffers)
    dom.step_em(dt, cfj)
    dom.history_async(energies_t[0].data_ptr(), counts_t[0].data_ptr())
    dom.interior_download_overlapped(core.FIELD_UJ, uji_host.data_ptr())
    dom.interior_download_overlapped(core.FIELD_UF, ufi_host.data_ptr())
    dom.copy_synchronize()
    dom.set_profiling(True)
    dom.phase_ms()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    dom.interior_upload_overlapped(core.FIELD_UF, ufi_host.data_ptr())
    dom.exchange_field()
    for k in range(args.steps):
        dom.step_em(dt, cfj)
        # history diagnostics of this step (field energies, particle counts per chunk = Chunk::load's input) into
        # pinned host memory, in stream order: read after the interval, the device never waits for the host
        dom.history_async(energies_t[k].data_ptr(), counts_t[k].data_ptr())
    dom.interior_download_overlapped(core.FIELD_UJ, uji_host.data_ptr())
    dom.interior_download_overlapped(core.FIELD_UF, ufi_host.data_ptr())
    dom.copy_synchronize()
    e3.record(stream)
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    phases_e2e = dom.phase_ms()
    dom.set_profiling(False)
    h2d_step = ufi_host.numel() * 8 / args.steps
    d2h_step = (ufi_host.numel() + uji_host.numel()) * 8 / args.steps + nchunk * (16 + 8 * prob.ns)
    if int(counts_t[-1].sum()) != dom.total_particles():
        raise SystemExit("bench.py: the particle counts read back during the e2e leg do not add up")
    err = dom.check()
    if err:
        raise SystemExit(f"bench.py: device error bits {err} during the timed region")

    # ---- the round-1 flavour for comparison: HOST-side field solver, interior E/B up and interior J down
    #      every step on a second stream (what the device-side solver removes) ----
    if not skip_e2e:
        ufi_host.numpy()[...] = ufn.reshape((nchunk,) + tuple(prob.M) + (6,))[
            :, nbh:nbh + prob.dims[0], nbh:nbh + prob.dims[1], nbh:nbh + prob.dims[2]].reshape(nchunk, icells, 6)
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record(stream)
    dom.interior_upload_overlapped(core.FIELD_UF, ufi_host.data_ptr())
    for k in range(args.steps):
        dom.exchange_field()  # ghosts of the E/B that has just arrived
        dom.clear_current()
        dom.push_deposit(dt)
        dom.exchange_current()
        dom.interior_download_overlapped(core.FIELD_UJ, uji_host.data_ptr())
        dom.migrate_sort()
        dom.copy_synchronize()  # J is on the host: the field solver would run here
        if k + 1 < args.steps:
            dom.interior_upload_overlapped(core.FIELD_UF, ufi_host.data_ptr())  # its result, for the next step
        dom.get_np(0)
    e5.record(stream)
    barrier()
    ms_e2e_host = e4.elapsed_time(e5)
    err = dom.check()
    if err:
        raise SystemExit(f"bench.py: device error bits {err} during the timed region")
    ntot_end = dom.total_particles()
    traffic = dom.peer_traffic()

    # ---- secondary line: the fp32 mode (north star: "1e-5 (fp32 mode)"; SURVEY 8d: 120 B per particle-update),
    #      same workload, same steps.  The headline stays fp64 = the reference's real type. ----
    fp32 = None
    if not args.no_fp32:
        dom.close()
        torch.cuda.empty_cache()
        d32 = core.Domain(prob.cdims, prob.dims, prob.nb, prob.order, prob.q, prob.m, coord=prob.coord, device=local,
                          id_range=(ids[0], ids[-1] + 1), strict_fp=False, capacity_factor=1.12,
                          stream=stream.cuda_stream, fp32=True)
        if world > 1:
            d32.set_ranks(bd, rank)
            d32.comm_init_torch()
        d32.field_upload_async(core.FIELD_UF, uf_host.data_ptr())
        d32.exchange_field()
        load_particles(d32)
        for _ in range(3):
            d32.step(dt)
        d32.set_profiling(True)
        d32.phase_ms()
        ms32 = timed_steps(args.steps, lambda: d32.step(dt))
        ph32 = d32.phase_ms()
        d32.set_profiling(False)
        ms32_em = timed_steps(args.steps, lambda: d32.step_em(dt, cfj))
        err = d32.check()
        if err:
            raise SystemExit(f"bench.py: device error bits {err} in the fp32 leg")
        n32 = d32.total_particles()
        fp32 = (ms32, ph32, ms32_em, n32)
        dom = d32

    t = torch.tensor([ms, ms_e2e, ms_other, ms_e2e_host, fp32[0] if fp32 else 0.0, fp32[2] if fp32 else 0.0],
                     dtype=torch.float64, device="cuda")
    n = torch.tensor([float(ntot), float(ntot_end), float(traffic["particles_sent"]),
                      float(traffic["halo_cells_sent"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    ms, ms_e2e, ms_other, ms_e2e_host, ms32_max, ms32_em_max = (float(v) for v in t)
    nglobal = float(n[0])

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # per kernel: CUDA events recorded by the library around every launch on the domain's stream
        prof = {}
        tpath = os.path.join(ROOT, "profiles", "push_deposit_traffic.json")
        if os.path.exists(tpath):
            try:
                with open(tpath) as f:
                    prof = json.load(f)
            except (OSError, ValueError):
                prof = {}
        traffic_by_kernel = prof.get("dram_bytes_per_particle", {}) if w["name"] == "cfg2" else {}
        fp64_by_kernel = prof.get("fp64_warp_inst_per_particle", {}) if w["name"] == "cfg2" else {}
        peak_fma = peak64["fp64_gfma_per_s"] * 1e9 if peak64 and "fp64_gfma_per_s" in peak64 else None
        kern = []
        for name in ("k_push", "k_deposit"):
            k_ms, k_calls = phases[name]
            launch_ms = k_ms / max(k_calls, 1)  # one call = one launch = one species
            part = ntot / prob.ns
            ach = ALGO_BYTES[name] * part / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0
            tr = traffic_by_kernel.get(name)
            wi = fp64_by_kernel.get(name)
            kern.append({"kernel": name + "<%d>" % w["order"], "launch_ms": launch_ms, "launches": k_calls,
                         "algorithmic_bytes_per_particle": ALGO_BYTES[name], "particles_per_launch": part,
                         "achieved": ach, "frac": ach / peak if peak else None,
                         "traffic": tr * part if tr else None,
                         # fp64 pipe: warp instructions (ncu) x 32 lanes / launch time / measured DFMA peak
                         "fp64_frac": (wi * 32 * part / (launch_ms * 1e-3) / peak_fma) if (wi and peak_fma and launch_ms > 0) else None})
        dom_k = max(kern, key=lambda k: k["launch_ms"])
        # whole step against SURVEY 8(d)'s 232 B per particle-update (push+deposit+count 116 B, sort 116 B)
        step_gbs = 232.0 * ntot / (ms / args.steps * 1e-3) / 1e9
        value = nglobal * args.steps / (ms * 1e-3)
        other = nglobal * args.steps / (ms_other * 1e-3)
        out = {
            "metric": METRIC, "value": value, "unit": "particle-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "%s: %s; %dx%dx%d chunks of %s per GPU in the reference's Gilbert order"
                            % (w["name"], w["label"], w["cdims"][0], w["cdims"][1], w["cdims"][2],
                               "x".join(map(str, w["dims"]))),
                "particles_per_gpu": ntot, "particles_end_all_ranks": int(n[1]),
                "fp_contract": "off" if args.strict else "fma",
                "l2": "inputs (%.1f GB of particles per GPU) larger than L2" % (ntot * 56 / 1e9),
                "multi_gpu": ("one periodic box of %dx%dx%d chunks, ids along the reference's Gilbert curve, one "
                              "contiguous curve segment per rank (%d ranks); J / E/B halo and particle migration "
                              "between ranks over NCCL send/recv, one message per peer and mode; last step: %d "
                              "particles and %d ghost cells per exchange crossed rank boundaries"
                              % (prob.cdims + (world, int(n[2]), int(n[3])))) if world > 1 else "single GPU",
            },
            "clocks": clocks,
            # the bit-exact (no FMA contraction) mode next to the contracted one, same steps, same data
            "strict_value": value if args.strict else other,
            "strict_ms_per_step": (ms if args.strict else ms_other) / args.steps,
            "fma_value": other if args.strict else value,
            "e2e": {"value": nglobal * args.steps / (ms_e2e * 1e-3), "unit": "particle-updates/s",
                    "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h_step),
                    "ms_per_step": ms_e2e / args.steps,
                    "what": "one diagnostic interval of %d steps through the C ABI with host buffers: interior E/B of "
                            "every chunk up from pinned host memory at the start, then per step "
                            "nixb200_domain_step_em (the PIC step + Yee field update on the device) and the read-back "
                            "of per-chunk field energies and particle counts, and the interior J + E/B snapshot of "
                            "every chunk down to pinned host memory at the end; bytes are totals / steps" % args.steps},
            "e2e_host_fields": {"value": nglobal * args.steps / (ms_e2e_host * 1e-3), "unit": "particle-updates/s",
                                "h2d_bytes_per_step": int(ufi_host.numel() * 8),
                                "d2h_bytes_per_step": int(uji_host.numel() * 8 + nchunk * 8),
                                "what": "round-1 flavour: HOST-side field solver, interior E/B up and interior J down "
                                        "every step on a second stream"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom_k["kernel"], "achieved": dom_k["achieved"], "peak": peak,
                         "unit": "GB/s", "frac": dom_k["frac"], "traffic": dom_k["traffic"],
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_particle": dom_k["algorithmic_bytes_per_particle"],
                         "launch_ms": dom_k["launch_ms"], "kernels": kern,
                         "step": {"algorithmic_bytes_per_update": 232.0, "achieved": step_gbs, "frac": step_gbs / peak},
                         "fp64_peak": peak64,
                         "note": "dominant kernel by time; at order >= 2 / fp64 k_deposit is bound by the fp64 pipe "
                                 "and issue latency and k_push by the shared-memory data pipe, not by HBM "
                                 "(DESIGN.md 3.2, SURVEY.md 8d): fp64_frac = fp64 warp instructions x 32 / time / "
                                 "measured DFMA peak"},
            "rebalance": rebal,
            "phases_ms_per_step": {k: (v[0] / args.steps) for k, v in phases.items()},
            "phases_ms_per_step_e2e": {k: (v[0] / args.steps) for k, v in phases_e2e.items()},
            "field_energy_last": [float(energies[-1, :, 0].sum()), float(energies[-1, :, 1].sum())],
        }
        if fp32:
            ms32, ph32, ms32_em, n32 = fp32
            k32 = []
            for name, ab in (("k_push", 64.0), ("k_deposit", 24.0)):
                k_ms, k_calls = ph32[name]
                lm = k_ms / max(k_calls, 1)
                part = n32 / prob.ns
                ach = ab * part / (lm * 1e-3) / 1e9 if lm > 0 else 0.0
                k32.append({"kernel": name + "<%d,f32>" % w["order"], "launch_ms": lm, "algorithmic_bytes_per_particle": ab,
                            "achieved": ach, "frac": ach / peak})
            g32 = 120.0 * n32 / (ms32_max / args.steps * 1e-3) / 1e9
            out["fp32"] = {
                "dtype": "f32", "value": nglobal * args.steps / (ms32_max * 1e-3), "ms_per_step": ms32_max / args.steps,
                "value_with_field_solver": nglobal * args.steps / (ms32_em_max * 1e-3),
                "phases_ms_per_step": {k: (v[0] / args.steps) for k, v in ph32.items()},
                "kernels": k32,
                "step": {"algorithmic_bytes_per_update": 120.0, "achieved": g32, "frac": g32 / peak},
                "what": "same workload and steps with fp32 particles, E/B and J on the device (positions relative to "
                        "the chunk origin); agrees with the fp64 oracle to 1e-5 (tests/test_gpu_fp32.py); no reference "
                        "exists for this mode, the headline stays fp64"}
        if world == 1 and not args.no_cpu:
            r = cpu_reference_run(w, steps=3, warmup=1, budget_s=25.0)
            out["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "backend", "simd_lanes", "scalar_path_value")}
        emit(out)
    dom.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--fma", dest="strict", action="store_false",
                    help="headline in the contracted (FMA) mode; the default headline is the bit-exact mode (no FMA "
                         "contraction: particles, counts and sort permutations identical to the reference's scalar "
                         "templates).  Both modes are timed and reported either way (strict_value / fma_value)")
    ap.add_argument("--strict", dest="strict", action="store_true", help="(default) bit-exact mode as the headline")
    ap.set_defaults(strict=True)
    ap.add_argument("--small", action="store_true", help="2x2x2 chunks (smoke / profiling)")
    ap.add_argument("--cdims", default="", help="override chunks per axis, e.g. 4,4,4")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-fp32", action="store_true", help="skip the secondary fp32 line")
    # exploration only (the contract line is the default: order 2, 16^3-cell chunks, 64 ppc per species)
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS), help="BASELINE.json config (default cfg2 = configs[1])")
    ap.add_argument("--order", type=int, default=0, help="override the shape order 1/2/3")
    ap.add_argument("--chunk", default="", help="override cells per chunk, e.g. 32,32,32")
    ap.add_argument("--ppc", type=int, default=0, help="override particles per cell and species")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the per-tick world step (BASELINE.json: organism-ticks/s at 100k organisms, + HBM roofline fraction).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--organisms M] [--impl ours|reference]

A step is one whole world tick (rules incl. think + dopamine training + reproduction/mutation, Space.step, friction,
post) over one synthetic world of M organisms per GPU (population islands: weak scaling; SURVEY.md section 8(d, e)).
`value` is device-timed with CUDA events, state resident in HBM; `e2e` goes through sapio_tick_host with pinned HOST buffers
(event block up, counters + packed sprite poses down, copies inside the timed region).
`--impl reference` times the CPU restatement of the reference (oracle/, kind "port": pymunk is not installable here, see
DESIGN.md) on a bounded sample of the same workload: the reference is single-threaded, so every host core runs its own world.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "organism_ticks_per_sec"
UNIT = "organism-ticks/s"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop, self.index = [], set(), None, threading.Event(), index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.thread.start(); return self

    def __exit__(self, *a):
        self._stop.set(); self.thread.join(timeout=3)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_world(orc, n_org, seed):
    """The same synthetic recipe as sapiogenesis_b200.synth.build_world, built on the host for the CPU arm."""
    import math
    import numpy as np
    import torch
    from sapiogenesis_b200.synth import REF_AREA
    from sapiogenesis_b200.world import World, default_params
    side = math.sqrt(n_org * REF_AREA)
    n_slots = int(math.ceil(n_org * 1.25 / 64.0) * 64)
    p = default_params(n_org=n_slots, n_dead=max(1024, 8 * n_org), width=side, height=side, seed=seed, population_limit=n_slots - max(64, n_slots // 100),
                       x_cap=max(4096, 4 * n_org))
    w = World(p)
    rng = np.random.default_rng(seed)
    cols = int(math.ceil(math.sqrt(n_org)))
    pitch = (side - 200.0) / cols
    w.t["g_co2"][0] = 30000.0 * (2.0 * side) / (3180.0 + 1800.0)
    for i in range(n_org):
        j = rng.uniform(-0.12, 0.12, 2) * pitch
        assert orc.spawn(w, i, 100.0 + (i % cols + 0.5) * pitch + j[0], 100.0 + (i // cols + 0.5) * pitch + j[1]) == 0
    nf = n_org // 4
    t = w.t
    t["d_state"][:nf] = 1
    t["d_pos"][:nf] = torch.from_numpy(rng.uniform(60.0, side - 60.0, (nf, 2)).astype("float32"))
    t["d_size"][:nf] = 30.0; t["d_mass"][:nf] = 12.0; t["d_mass0"][:nf] = 12.0; t["d_elast"][:nf] = .5; t["d_fric"][:nf] = .5; t["d_owner"][:nf] = -1
    o = torch.arange(n_org, dtype=torch.int32)
    t["org_think_tick"][:n_org] = -(o % p.think_ticks); t["org_dopa_tick"][:n_org] = -(o % p.think_ticks) - 1
    t["org_train_tick"][:n_org] = -((o + 7) % p.train_ticks); t["org_lastbirth_tick"][:n_org] = -(o % p.repro_ticks)
    orc.post(w)
    return w


def _cpu_island(args):
    """One CPU island: its own world on its own core (the reference is single-threaded; independent worlds are how it uses a
    multi-core host). Returns (organism-ticks done, seconds) per timed step."""
    n_org, seed, warm, steps, ticks, barrier = args
    from oracle import oracle as orc
    w = cpu_world(orc, n_org, seed)
    orc.tick(w, 127, max(warm, 3))                        # warm-up: contacts and adjacency exist
    barrier.wait(600)
    out = []
    for _ in range(steps):
        before = w.stats()["org_ticks"]
        t0 = time.perf_counter()
        orc.tick(w, 127, ticks)
        out.append((w.stats()["org_ticks"] - before, time.perf_counter() - t0))
    return out


def cpu_islands(n_org, warm, steps, ticks, cores=None):
    """The restated reference on every host core: one independent world per core, started together.
    Returns (organism-ticks/s over all cores, per-step wall seconds (region time of the slowest core / steps), cores)."""
    import multiprocessing as mp
    from oracle import oracle as orc
    orc.build()
    cores = cores or max(1, min(os.cpu_count() or 1, 64))
    ctx = mp.get_context("fork")
    with ctx.Manager() as mgr:
        barrier = mgr.Barrier(cores)
        with ctx.Pool(cores) as pool:
            res = pool.map_async(_cpu_island, [(n_org, seed, warm, steps, ticks, barrier) for seed in range(cores)]).get(timeout=1200)
    units = sum(u for r in res for u, _ in r)
    total = max(sum(t for _, t in r) for r in res)         # the timed region of K steps, max over cores (as for the GPU ranks)
    step_s = [total / steps] * steps
    return units / total, step_s, cores


def cpu_baseline(n_org=1000, ticks=24):
    """The CPU arm as a bounded sample, in a fresh interpreter (forking worker processes out of a process that holds a CUDA
    context is not safe)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "3",
           "--ref-organisms", str(n_org), "--ref-ticks", str(ticks)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "0", "WORLD_SIZE": "1"})
        line = json.loads(r.stdout.strip().splitlines()[-1])
        out = line["cpu_baseline"]
        out["host_cores_available"] = os.cpu_count()
        return out
    except Exception as e:                                   # the GPU numbers do not depend on it
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"CPU arm failed: {e!r}"[:300]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_org, ticks = args.ref_organisms, args.ref_ticks or 2
    v, step_s, cores = cpu_islands(n_org, args.warmup, args.steps, ticks)
    total_t = sum(step_s)
    sample = (f"{cores} independent worlds (one per host core) x {n_org} organisms x {ticks} ticks per step, oracle/ C restatement of the reference with indexed sensor candidates and per-victim contact lists "
              f"(pymunk/Chipmunk not installable: DESIGN.md)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args, args.organisms), "organisms_per_gpu": args.organisms,
                   "sample": f"bounded sample of that workload: {cores} x {n_org} organisms x {ticks} ticks per step, young worlds (the port's cost per organism-tick "
                             f"does not depend on the world size: broadphase grid and per-organism loops are linear)", "organisms": cores * n_org},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def world_shape(w, eng):
    """Realised sizes of the world, the inputs of the algorithmic-byte formulas, split into organisms without
    and with cross contacts (the organism solver and the coupled pass share the formula, on their own organisms)."""
    import torch
    from sapiogenesis_b200.world import MAXC
    t = w.t
    org = t["org_state"] == 1
    alive = (t["c_alive"] == 1).view(-1, MAXC) & org[:, None]
    cells_per = alive.sum(1).to(torch.int64)
    types = t["c_type"].view(-1, MAXC)
    sens_per = (alive & ((types == 2) | (types == 3))).sum(1).to(torch.int64)
    off = t["xa_off"].to(torch.int64)
    deg = (off[1:] - off[:-1])
    NC = w.p.n_org * MAXC
    org_deg = deg[:NC].view(-1, MAXC).sum(1)
    coupled = org & (org_deg > 0)
    free = org & ~coupled
    icn = t["org_ic_n"].to(torch.int64)
    dead_live = t["d_state"] > 0
    dead_coupled = int((dead_live & (deg[NC:NC + w.p.n_dead] > 0)).sum())
    def part(m):
        c = cells_per[m]
        return dict(organisms=int(m.sum()), cells=int(c.sum()), sensors=int(sens_per[m].sum()), pairs=int((c * (c - 1) // 2).sum()), intra_contacts=int(icn[m].sum()))
    return dict(free=part(free), coupled=part(coupled), all=part(org), dead=int(dead_live.sum()), dead_coupled=dead_coupled,
                cross_contacts=int(t["g_x_n"].item()), cells_hist=torch.bincount(cells_per[free], minlength=MAXC + 1).cpu().tolist(),
                icn_by_cells=(torch.zeros(MAXC + 1, dtype=torch.float64, device=cells_per.device).index_add_(0, cells_per[free], icn[free].double())).cpu().tolist(),
                mean_memory_rows=float(t["org_mem_n"][org].float().mean()) if int(org.sum()) else 0.0)


def solver_bytes(part):
    return 74 * part["cells"] + 32 * part["sensors"] + 8 * part["pairs"] + 20 * part["intra_contacts"]


def chain_bound(shape, info, ms, mhz, iterations=10, ideal_cycles=70.0):
    """The yardstick that applies to the organism solver (it is not memory-bound): a sweep over an organism of n living cells is a
    dependent chain of 2n-3 pin wavefronts + the sensor-pin step + its contact level steps; organisms in flight per SM are what the
    shared-memory footprint of a size class admits. ideal = steps / (SMs x lane groups in flight) x 70 cycles (one shared-memory load,
    seven dependent fp32 operations, the store and the warp barrier) at the clock of the run; achieved cycles per step next to it."""
    hist, icn = shape["cells_hist"], shape["icn_by_cells"]
    ideal_ms = 0.0; steps_total = 0.0; slots = 0.0
    for n in range(1, len(hist)):
        if not hist[n]:
            continue
        cls = 0 if n <= 8 else (n - 1) >> 3
        i = info[cls]
        groups = max(1, i["resident_ctas"] // max(i["sms"], 1)) * i["warps_per_cta"] * i["orgs_per_warp"]
        k = icn[n] / hist[n]
        contact_steps = max(1.0, 2.0 * k / n) if k > 0 else 0.0          # levels of a tree-like contact set, lanes >= the widest level
        steps = hist[n] * iterations * (max(2 * n - 3, 0) + 1 + contact_steps)
        steps_total += steps; slots += steps / (i["sms"] * groups)
    ideal_ms = slots * ideal_cycles / (mhz * 1e3)
    return {"model": "sum over size classes of wavefront steps / (SMs x organisms in flight per SM) x cycles per step; classes run back to back",
            "wavefront_steps_per_tick": steps_total, "serial_steps_per_lane_group": slots, "ideal_cycles_per_step": ideal_cycles,
            "ideal_ms": ideal_ms, "achieved_ms": ms, "frac": ideal_ms / ms if ms else None,
            "achieved_cycles_per_step": ms * mhz * 1e3 / slots if slots else None, "sm_mhz": mhz}


def ncu_traffic():
    """DRAM bytes per tick of the solver kernels from the committed ncu capture (tools/ncu_capture.sh -> profiles/ncu_traffic.json),
    only if it was taken on exactly these kernel sources."""
    try:
        from sapiogenesis_b200.build import source_hash
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            d = json.load(f)
        return d if d.get("source_sha") == source_hash() else None
    except Exception:
        return None


def timed_ticks(isl, eng, w, n, barrier, dist, world_size, dev, profile=False):
    """n ticks, device-timed (CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks)."""
    import torch
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ticks0 = w.stats()["org_ticks"]; launches0 = eng.launch_count()
    if profile:
        eng.profile(True)
    barrier()
    ev0.record()
    isl.step(n)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    prof = None
    if profile:
        prof = eng.profile_read(reset=True); eng.profile(False)
    units = w.stats()["org_ticks"] - ticks0
    t = torch.tensor([ms, float(units)], dtype=torch.float64, device=dev)
    if world_size > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, units = float(tmax[0]), float(tsum[1])
    return ms, units, prof, eng.launch_count() - launches0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from sapiogenesis_b200.synth import build_world
    from sapiogenesis_b200.world import PH_ALL, PH_REPRODUCE, PH_TRAIN

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the world step has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    clk = ClockSampler(local_rank)
    clk.__enter__()                                        # sampled over the whole run (ageing included): the timed region is a fraction of a second
    M = args.organisms
    w, eng = build_world(M, device=dev, seed=rank)
    from sapiogenesis_b200.islands import Island
    if args.forward_only:                                # the same as maladaptive=True for everybody and population_limit=0
        w.t["org_flags"] |= 2
        w.p.population_limit = 0
        eng.refresh()
    # N > 1: population islands with ring migration of whole organisms (the only exchange step of the path; NCCL p2p). The period is
    # chosen so that the exchange fires INSIDE the timed region whatever --steps is.
    every = min(args.migrate_every, max(args.steps // 2, 1)) if (world_size > 1 and args.migrate_every) else 0
    isl = Island(w, eng, every=every, count=args.migrate_count, seed=1234)
    warm = max(args.warmup, 3)
    isl.step(warm)
    if every:
        isl.migrate()                                   # warm-up of the exchange step too (NCCL p2p channels are set up lazily)
    eng.synchronize()

    # ---- the young world (round-1 headline, kept as a secondary figure): ticks warm .. warm + 20 -----------------------------------
    y_ms, y_units, _, _ = timed_ticks(isl, eng, w, 20, barrier, dist, world_size, dev)
    young = {"value": y_units / (y_ms / 1e3), "ms_per_step": y_ms / 20, "world_age_ticks": warm + 20,
             "note": "the same world before ageing: memories empty, nobody reproduces, hardly any contact between organisms"}

    # ---- ageing prelude, untimed: the measured world is a mature one ---------------------------------------------------------------
    age = 0 if args.forward_only else max(args.age - (warm + 20), 0)
    t_age = time.perf_counter()
    done = 0
    while done < age:                                   # in chunks: the launch queue stays short and a hung tick shows up
        run = min(500, age - done)
        isl.step(run); eng.synchronize(); done += run
    t_age = time.perf_counter() - t_age
    c0 = isl.counts() if every else {}
    migr0 = (c0.get("total_out", 0), isl.collect_migration_ms() if every else 0.0, isl.migrations, c0.get("total_in", 0), c0.get("total_dropped", 0))

    # ---- device-resident throughput -------------------------------------------------------------------------------------
    w.t["g_stats"].view(-1)[14] = 0                      # rare-path bits: evidence of the timed region only
    st0 = w.stats()
    region0 = len(clk.samples)
    ms, units, prof, launches = timed_ticks(isl, eng, w, args.steps, barrier, dist, world_size, dev, profile=True)
    region_samples = len(clk.samples) - region0
    st1 = w.stats()
    value = units / (ms / 1e3)
    migr = None
    if every:
        c1 = isl.counts()
        migr = {"migrations_in_region": isl.migrations - migr0[2], "organisms_sent_by_rank0": c1["total_out"] - migr0[0],
                "organisms_placed_on_rank0": c1["total_in"] - migr0[3], "arrivals_dropped_on_rank0": c1["total_dropped"] - migr0[4],
                "migration_ms_in_region": round(isl.collect_migration_ms() - migr0[1], 3), "buffer_bytes_per_migration": isl.last_bytes,
                "record_format": "ragged (include/sapio_migrate.h): extents in use only, ~40 kB per organism"}

    # ---- end to end through the C-ABI with host buffers ---------------------------------------------------------------------
    e2e_steps = max(3, min(args.steps, 20))
    eng.tick_host(1, PH_ALL)                               # pinned buffers, copy stream and the size guess exist after these
    eng.tick_host_begin(1, PH_ALL); eng.tick_host_end()
    eng.tick_host_begin(1, PH_ALL); eng.tick_host_end()
    ticks1 = w.stats()["org_ticks"]
    h2d = d2h = 0
    barrier()
    t0 = time.perf_counter()
    eng.tick_host_begin(1, PH_ALL)                         # frame k downloads (copy stream, pinned host memory) while tick k + 1 computes;
    for _ in range(e2e_steps - 1):                         # every step uploads its event block and receives its own frame
        eng.tick_host_begin(1, PH_ALL)
        hdr, cells, dead, (bi, bo) = eng.tick_host_end()
        h2d += bi; d2h += bo
    hdr, cells, dead, (bi, bo) = eng.tick_host_end()
    h2d += bi; d2h += bo
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    e2e_units = w.stats()["org_ticks"] - ticks1
    t = torch.tensor([e2e_s, float(e2e_units)], dtype=torch.float64, device=dev)
    if world_size > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        e2e_s, e2e_units = float(tmax[0]), float(tsum[1])
    e2e_value = e2e_units / e2e_s
    clk.__exit__()

    if rank == 0:
        shape = world_shape(w, eng)
        s = shape["all"]
        info = eng.solver_info()
        state_gib = w.nbytes() / 2 ** 30
        # ---- BASELINE configs[1] as a sub-record: 10k organisms, physics + sensors + brain forward only --------------------------
        fwd = None
        if world_size == 1 and not args.forward_only and not args.no_forward_record:
            del isl
            w2, eng2 = build_world(10_000, device=dev, seed=rank)
            w2.t["org_flags"] |= 2; w2.p.population_limit = 0; eng2.refresh(); eng2.exclusive = True
            ph = PH_ALL & ~(PH_TRAIN | PH_REPRODUCE)
            eng2.tick(200, ph); eng2.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            u0 = w2.stats()["org_ticks"]
            e0.record(); eng2.tick(200, ph); e1.record(); eng2.synchronize()
            fms = e0.elapsed_time(e1)
            fwd = {"workload": "BASELINE configs[1]: 10k organisms, one world, physics + sensors + brain forward only (no dopamine training, no reproduction)",
                   "value": (w2.stats()["org_ticks"] - u0) / (fms / 1e3), "unit": UNIT, "ms_per_step": fms / 200, "steps": 200, "world_age_ticks": 400}
            del w2, eng2
        # ---- roofline of the dominant kernel (algorithmic bytes: DESIGN.md "Kernels") -----------------------------------
        peak, peak_src = peaks()
        stage_ms = {k: v[0] / max(v[1], 1) for k, v in prof.items() if v[1] > 0}
        dom = max(stage_ms, key=stage_ms.get) if stage_ms else "org_solve"
        traffic = ncu_traffic()
        clocks = clk.summary()
        mhz = clocks["sm_mhz"] or 1965.0
        kernels = {
            "org_solve": ("k_org_solve<8|16|24|32|40|48|56|64> + k_component<1|2|4|16> (Space.step's solver: per organism its intra contacts, sensor pins and "
                          "all-pairs pin joints, exact warm start + 10 Gauss-Seidel sweeps with everything resident in shared memory; organisms without cross "
                          "contacts by size class, organisms and dead bodies that touch each other one connected component per CTA; 15 concurrent launches)",
                          solver_bytes(shape["free"]) + solver_bytes(shape["coupled"]) + 64 * shape["cross_contacts"] + 36 * shape["dead_coupled"]),
            "coupled": ("k_coupled_simple (cooperative fallback for components beyond the largest CTA tier)", 64 * shape["cross_contacts"]),
        }
        roof = None
        rk = dom if dom in kernels else "org_solve"
        if rk in stage_ms:
            name, alg = kernels[rk]
            ach = alg / (stage_ms[rk] / 1e3) / 1e9
            tr = traffic["kernels"].get(rk) if (traffic and traffic.get("organisms") == M and traffic.get("world_age_ticks") == args.age) else None
            roof = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": tr, "traffic_source": (f"profiles/ncu_traffic.json (ncu --set full, kernel sources {traffic['source_sha']})" if tr else
                                                      "null: no ncu capture of exactly these kernel sources and this workload is committed"),
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "avg_ms": stage_ms[rk], "share_of_step": stage_ms[rk] / (ms / args.steps),
                    "dominant_stage": dom,
                    "note": "not HBM-bound: each sweep is a dependent chain of 2n-3 pin wavefronts per organism; the yardstick that applies is chain_bound below"}
            if "org_solve" in stage_ms:
                roof["chain_bound"] = chain_bound(shape, info, stage_ms["org_solve"], mhz)
        # the streaming stages against the same peak (algorithmic bytes: DESIGN.md section 4)
        bodies = s["cells"] + shape["dead"]
        stream_bytes = {
            "friction": 24 * bodies, "broadphase": 36 * s["cells"] + 24 * s["sensors"] + 36 * shape["dead"] + 16 * bodies,
            "settle": 12 * bodies, "rules_main": 60 * s["cells"] + 128 * s["organisms"], "rules_begin": 60 * s["cells"] + 128 * s["organisms"],
        }
        stage_roof = {k: {"bytes": int(b), "ms": round(stage_ms[k], 4), "GB/s": round(b / (stage_ms[k] / 1e3) / 1e9, 1),
                          "frac": round(b / (stage_ms[k] / 1e3) / 1e9 / peak, 4)} for k, b in stream_bytes.items() if stage_ms.get(k)}
        cpu = cpu_baseline(args.ref_organisms, args.ref_ticks or 40) if (not args.no_cpu and world_size == 1) else None
        K = args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args, M),
                       "organisms_per_gpu": M, "world_age_ticks": int(w.tick) - e2e_steps - 3, "ageing_prelude_s": round(t_age, 1),
                       "living": s["organisms"], "births_per_tick": (st1["births"] - st0["births"]) / K, "deaths_per_tick": (st1["deaths"] - st0["deaths"]) / K,
                       "cell_deaths_per_tick": (st1["cell_deaths"] - st0["cell_deaths"]) / K, "trains_per_tick": (st1["trains"] - st0["trains"]) / K,
                       "thinks_per_tick": (st1["thinks"] - st0["thinks"]) / K, "mean_memory_rows": round(shape["mean_memory_rows"], 1),
                       "coupled_organisms": shape["coupled"]["organisms"], "rare_paths_in_region": st1["rare_paths"], "cells": s["cells"], "pin_joints": s["pairs"],
                       "dead_cells": shape["dead"], "cross_contacts": shape["cross_contacts"], "intra_contacts": s["intra_contacts"],
                       "state_gib_per_gpu": round(state_gib, 1), "l2_policy": "inputs larger than L2",
                       "parallelism": f"{world_size} island(s), one per GPU" + (f"; ring migration of {args.migrate_count} whole organisms every {every} ticks over NCCL p2p, "
                                      f"arrivals placed by the reference's placement search" if every else ""),
                       "migration": migr},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps,
                    "steps": e2e_steps, "api": "sapio_tick_host_begin / _end (pinned host buffers; per tick: event block up, counters + packed sprite poses down; the download of frame k overlaps tick k + 1)"},
            "gpu_launches": int(launches),
            "stage_ms": {k: round(v, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1])},
            "clocks": {**clocks, "samples_in_timed_region": region_samples, "sampled": "whole run incl. the ageing prelude"},
            "young_world": young,
            "world_stats": st1,
        }
        if fwd:
            line["forward_only_10k"] = fwd
        if roof:
            line["roofline"] = roof
            line["stage_roofline"] = stage_roof
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


def run_slabs(args):
    """BASELINE configs[4]: ONE world of N x M organisms cut along x into N slabs, one per GPU; every tick the slabs exchange the organisms
    and dead cells within the halo of their common boundaries (NCCL p2p) and compose the gas maps / refunds / dispersion / population in
    rank order (small allgathers). M organisms per GPU: weak scaling in the driver's sense (1M organisms at N = 8 with --organisms 125000)."""
    import torch
    import torch.distributed as dist
    from sapiogenesis_b200.slabs import DistComm, LocalComm, SlabDriver, build_slab

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the world step has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    clk = ClockSampler(local_rank); clk.__enter__()
    M = args.organisms
    slab = build_slab(M * world_size, rank, world_size, dev, seed=7)
    drv = SlabDriver([slab], DistComm() if world_size > 1 else LocalComm(1))
    warm = max(args.warmup, 3)
    drv.tick(warm); drv.synchronize()
    age = max(args.age - warm, 0)
    t_age = time.perf_counter()
    done = 0
    while done < age:
        run = min(250, age - done)
        drv.tick(run); drv.synchronize(); done += run
    t_age = time.perf_counter() - t_age
    w, eng = slab.world, slab.engine
    st0 = w.stats(); halo0 = slab.halo_bytes; launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.profile(True)
    barrier()
    ev0.record()
    drv.tick(args.steps)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    prof = eng.profile_read(reset=True); eng.profile(False)
    st1 = w.stats()
    units = st1["org_ticks"] - st0["org_ticks"]                     # owned organisms only: a ghost is its owner's organism
    owned = int(slab.owned().sum()); ghosts = int((w.t["org_state"] == 1).sum()) - owned
    t = torch.tensor([ms, float(units), float(owned), float(ghosts), float(slab.halo_bytes - halo0)], dtype=torch.float64, device=dev)
    if world_size > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, units, owned_all, ghosts_all, halo_all = float(tmax[0]), float(tsum[1]), float(tsum[2]), float(tsum[3]), float(tsum[4])
    else:
        owned_all, ghosts_all, halo_all = float(owned), float(ghosts), float(t[4])
    clk.__exit__()
    if rank == 0:
        stage_ms = {k: v[0] / max(v[1], 1) for k, v in prof.items() if v[1] > 0}
        line = {
            "metric": METRIC, "value": units / (ms / 1e3), "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"full world tick of ONE world of {M * world_size} organisms at spawn cut into {world_size} spatial slab(s) along x "
                                   f"(BASELINE configs[4]), aged {args.age} ticks before the timed region",
                       "mode": "slabs", "organisms_per_gpu": M, "world_age_ticks": int(w.tick), "ageing_prelude_s": round(t_age, 1),
                       "living_owned": int(owned_all), "ghost_copies": int(ghosts_all), "halo_px": slab.halo,
                       "halo_bytes_per_tick_all_ranks": int(halo_all / args.steps),
                       "halo_buffer_bytes_per_side": int(slab.send[0][0].numel() + slab.send[0][1].numel()),
                       "exchange": "per tick: ragged organism records + dead cells within the halo to each neighbour (NCCL send/recv), allgathers of the "
                                   "gas map (16 B), refund (8 B), dispersion (8 B) and owned population (4 B) composed in rank order",
                       "l2_policy": "inputs larger than L2",
                       "parallelism": f"{world_size} slab(s) of one world, one per GPU"},
            "e2e": None, "e2e_note": "the slab mode has no host-buffer entry point: frames are downloaded per slab with the island API (sapio_tick_host)",
            "gpu_launches": int(eng.launch_count() - launches0),
            "stage_ms": {k: round(v, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1])},
            "clocks": clk.summary(), "world_stats_rank0": st1,
        }
        print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


def workload_name(args, M):
    if args.forward_only:
        return f"forward-only tick (rules + think + Space.step + friction; no dopamine training, no reproduction: BASELINE configs[1]), {M} organisms per GPU island"
    return (f"full world tick (rules + think + dopamine training + reproduction/mutation + Space.step + friction), {M} organisms per GPU island at spawn "
            f"(BASELINE configs[2]), world aged {args.age} ticks before the timed region")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--organisms", type=int, default=100_000, help="organisms per GPU (island)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-organisms", type=int, default=1000, help="organisms per host core of the CPU arm's bounded sample")
    ap.add_argument("--ref-ticks", type=int, default=0, help="ticks per CPU step (default: 40 for cpu_baseline, 2 per step for --impl reference)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--forward-only", action="store_true",
                    help="BASELINE configs[1]: physics + sensors + brain forward only (no dopamine training, no reproduction); use with --organisms 10000")
    ap.add_argument("--age", type=int, default=3000, help="world age (ticks) at the start of the timed region: untimed ageing prelude on the device")
    ap.add_argument("--no-forward-record", action="store_true", help="skip the 10k forward-only sub-record (BASELINE configs[1])")
    ap.add_argument("--migrate-every", type=int, default=50, help="N > 1: ticks between ring migrations (0 = never)")
    ap.add_argument("--migrate-count", type=int, default=64, help="organisms each island sends per migration")
    ap.add_argument("--mode", default="islands", choices=["islands", "slabs"],
                    help="multi-GPU form: independent population islands with ring migration (BASELINE configs[3], default) or spatial slabs of "
                         "one world of gpus x organisms with halo exchange (configs[4])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "slabs":
        run_slabs(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

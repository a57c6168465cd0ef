"""Headless runner: the wiring of Sapiogenesis.py:671-700 without Qt.

    python -m sapiogenesis_b200.run --organisms 12 --ticks 1000 --dump out.npz      # the reference's default world
    python -m sapiogenesis_b200.run --organisms 100000 --ticks 200 --every 50       # a synthetic island, stats every 50 ticks

The default world is the reference's (3180 x 1800, co2 30000, population limit 12) with `--organisms` creatures added the way
add_creature_clicked does (Sapiogenesis.py:297-323); larger worlds use the synthetic recipe of sapiogenesis_b200.synth.
`--dump` writes the last frame (what the GUI would draw: pose and owner of every living cell, pose and mass of every dead cell)
and the per-report counters.
"""
from __future__ import annotations

import argparse
import json
import time

import numpy as np
import torch


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--organisms", type=int, default=12)
    ap.add_argument("--ticks", type=int, default=1000)
    ap.add_argument("--every", type=int, default=100, help="ticks between reports")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--drought", type=float, default=0.0, help="severity of the drought event (Sapiogenesis.py:37-41)")
    ap.add_argument("--poison", type=float, default=0.0, help="severity of the poison event (Sapiogenesis.py:42-47)")
    ap.add_argument("--random-cells", type=float, default=0.0, help="severity of the random dead cell event (Sapiogenesis.py:48-53)")
    ap.add_argument("--dump", default=None, help="write the last frame and the report series to this .npz")
    args = ap.parse_args(argv)

    from . import physics, sprites
    from .synth import build_world
    if args.organisms <= 64:
        env = physics.Environment(None, None, 3180, 1800, capacity=max(64, 2 * args.organisms), seed=args.seed,
                                  population_limit=max(12, args.organisms))
        rng = np.random.default_rng(args.seed)
        for _ in range(args.organisms):
            env.add_creature((rng.uniform(300, 2880), rng.uniform(300, 1500)))
        env.preUpdateEvent = lambda: sprites.update_organisms(env)
        world, engine = env.world, env.engine
    else:
        world, engine = build_world(args.organisms, seed=args.seed)
        env = None
    engine.exclusive = env is None
    dt = world.p.dt_wall
    series, done, t0 = [], 0, time.perf_counter()
    while done < args.ticks:
        run = min(args.every, args.ticks - done)
        ev = dict(drought=args.drought * dt, poison=args.poison * dt, random_cells=args.random_cells * dt)
        if env is not None and not any(ev.values()):
            for _ in range(run):
                env.update(None)
            frame = None
        else:
            for _ in range(run - 1):
                engine.tick_host(1, **ev) if any(ev.values()) else engine.tick(1)
            frame = engine.tick_host(1, **ev)
        done += run
        engine.synchronize()
        st = world.stats()
        rep = dict(tick=world.tick, population=int((world.t["org_state"] == 1).sum()), cells=int((world.t["c_alive"] == 1).sum()),
                   dead_cells=int((world.t["d_state"] > 0).sum()), co2=float(world.t["g_co2"][0]), oxygen=float(world.t["g_o2"][0]),
                   births=st["births"], deaths=st["deaths"], thinks=st["thinks"], trains=st["trains"], overflow=st["overflow"],
                   ticks_per_s=round(done / (time.perf_counter() - t0), 1))
        series.append(rep)
        print(json.dumps(rep), flush=True)
    if args.dump:
        hdr, cells, dead, _ = engine.tick_host(0)
        np.savez_compressed(args.dump, cells=cells.numpy(), dead=dead.numpy(), header=json.dumps(hdr), series=json.dumps(series))
        print(f"wrote {args.dump}: {len(cells)} cells, {len(dead)} dead cells")


if __name__ == "__main__":
    main()

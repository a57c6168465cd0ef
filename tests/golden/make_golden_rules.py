"""Golden vectors for the rules phase: the REFERENCE's own sprites.update_organisms / Organism.update / _update_cell /
add_energy / take_energy / convert_co2 / convert_o2 / kill_cell / gradual_dispersion (sprites.py) run headless behind the
stand-ins of refshim.py on single organisms (no cross-organism interaction, so list order cannot matter), thinking and
training switched off. Output: tests/golden/rules_cases.npz. Run in the build container only."""
from __future__ import annotations

import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402
from make_golden import UI, build, feed, round_dna_to_f32  # noqa: E402

N_TICKS = 40
DT = 0.030


def f32(x):
    return float(np.float32(x))


def gen_rules(n_cases=15, seed=23):
    ns = refshim.load()
    out = {"meta": np.asarray([n_cases, seed, N_TICKS], dtype=np.int32)}
    none = lambda: refshim.UniformFeed([], "none")
    for ci in range(n_cases):
        tag = f"rules-{seed}-{ci}"
        rng = random.Random(tag + "-state")
        u, ui = feed(tag + "-u", 30000, "u"), feed(tag + "-ui", 600000, "init")
        refshim.set_feeds(ns, u, ui, none())
        dna = ns.sprites.DNA().randomize(**UI)
        used = [u.pos, ui.pos]
        round_dna_to_f32(dna)
        ub = feed(tag + "-ub", 600000, "build")
        refshim.set_feeds(ns, refshim.UniformFeed([0.5] * 100000, "gate"), ub, none())
        env = refshim.FakeEnv(ns)
        env.info["co2"] = f32(rng.uniform(5000, 30000))
        env.info["oxygen"] = f32(rng.uniform(0, 8000))
        env.info["population"] = 1
        env.info["population_limit"] = 0          # reproduction gate closed, energy halving still happens (sprites.py:1664-1678)
        pos = [f32(rng.uniform(400, 2700)), f32(rng.uniform(400, 1400))]
        org = build(ns, env, dna, pos)
        used.append(ub.pos)
        org.maladaptive = True                    # no training
        org.brain.lastUpdated = 1e18              # never thinks
        # bodies carry fp32 positions (what the world arrays hold)
        order = list(org.cells.keys())
        for cid in order:
            b = org.cells[cid].body
            b.position = (f32(b.position[0]), f32(b.position[1]))
        org.pos = [f32(org.cells[order[0]].body.position[0]), f32(org.cells[order[0]].body.position[1])]
        # state perturbation
        mode = ci % 5
        energy, health = [], []
        for k, cid in enumerate(order):
            sp = org.cells[cid]
            ci_ = sp.cell_info
            if mode == 0:      # spawn state
                e, h = ci_["energy_storage"] / 2, ci_["max_health"]
            elif mode == 1:    # nearly full: saturating add_energy, reproduction-gate halving
                e, h = ci_["energy_storage"] * rng.uniform(0.985, 1.0), ci_["max_health"] * rng.uniform(0.5, 1.0)
            elif mode == 2:    # starving
                e, h = ci_["energy_storage"] * rng.uniform(0.0, 0.05), ci_["max_health"] * rng.uniform(0.0, 0.12)
            elif mode == 3:    # some cells about to die -> cascades, break damage, dead bodies
                e, h = ci_["energy_storage"] * rng.uniform(0.1, 0.9), ci_["max_health"] * rng.uniform(0.2, 1.0)
                if k > 0 and rng.random() < 0.4:
                    h = -1.0 if rng.random() < 0.5 else 0.0
            else:
                e, h = ci_["energy_storage"] * rng.uniform(0.0, 1.0), ci_["max_health"] * rng.uniform(0.0, 1.0)
            sp.info["energy"], sp.info["health"] = f32(e), f32(h)
            energy.append(sp.info["energy"]); health.append(sp.info["health"])
        move = [f32(rng.uniform(-1, 1)) for _ in range(4)]
        if ci % 3 == 0:
            move[1] = f32(rng.uniform(-0.09, 0.09))          # |speed| < .1: push cells idle
        org.movement = {"rotation": move[0], "speed": move[1], "direction": [move[2], move[3]]}
        org.energy_consumed = f32(rng.uniform(0, 300) if ci % 2 else 0.0)
        org.lastHealth = f32(org.health_percent()); org.lastEnergy = f32(org.energy_percent())
        if mode == 1:
            org.lastBirth = ns.clock.now - 100.0             # reproduction_limit already satisfied
        out[f"c{ci}.used"] = np.asarray(used, dtype=np.int32)
        out[f"c{ci}.pos"] = np.asarray(pos, dtype=np.float64)
        out[f"c{ci}.env0"] = np.asarray([env.info["co2"], env.info["oxygen"]], dtype=np.float64)
        out[f"c{ci}.energy0"] = np.asarray(energy, dtype=np.float32)
        out[f"c{ci}.health0"] = np.asarray(health, dtype=np.float32)
        out[f"c{ci}.move"] = np.asarray(move, dtype=np.float32)
        out[f"c{ci}.org0"] = np.asarray([org.energy_consumed, org.lastHealth, org.lastEnergy, 1.0 if mode == 1 else 0.0], dtype=np.float64)
        nc = len(order)
        rec = dict(health=[], energy=[], alive=[], inuse=[], vel=[], angvel=[], org=[], env=[], dead=[])
        for t in range(N_TICKS):
            ns.clock.now += DT
            ns.sprites.update_organisms(env)
            # the world arrays are fp32: round the carried state like the oracle's store does
            for cid in order:
                sp = org.cells[cid]
                sp.info["energy"], sp.info["health"] = f32(sp.info["energy"]), f32(sp.info["health"])
                sp.body.velocity = (f32(sp.body.velocity[0]), f32(sp.body.velocity[1]))
                sp.body.angular_velocity = f32(sp.body.angular_velocity)
            org.energy_consumed = f32(org.energy_consumed); org.lastHealth = f32(org.lastHealth); org.lastEnergy = f32(org.lastEnergy)
            for s in env.sprites:
                if not s.alive:
                    s.body._set_mass(f32(s.body.mass))
            rec["health"].append([org.cells[c].info["health"] for c in order])
            rec["energy"].append([org.cells[c].info["energy"] for c in order])
            rec["alive"].append([1 if org.cells[c].alive else 0 for c in order])
            rec["inuse"].append([1 if org.cells[c].info["in_use"] else 0 for c in order])
            rec["vel"].append([[org.cells[c].body.velocity[0], org.cells[c].body.velocity[1]] for c in order])
            rec["angvel"].append([org.cells[c].body.angular_velocity for c in order])
            rec["org"].append([org.energy_consumed, org.lastHealth, org.lastEnergy, 1.0 if org.alive() else 0.0])
            rec["env"].append([env.info["co2"], env.info["oxygen"]])
            dead = sorted((s.body.mass, s.body.position[0], s.body.position[1]) for s in env.sprites if not s.alive)
            rec["dead"].append(dead + [(0.0, 0.0, 0.0)] * (nc - len(dead)))
        for k, v in rec.items():
            out[f"c{ci}.{k}"] = np.asarray(v, dtype=np.float64 if k in ("env", "org") else np.float32)
        n_dead = sum(1 for s in env.sprites if not s.alive)
        print(f"rules case {ci}: mode {mode}, cells {nc}, alive at end {sum(rec['alive'][-1])}, dead bodies {n_dead}, "
              f"co2 {env.info['co2']:.2f} o2 {env.info['oxygen']:.2f}")
    np.savez_compressed(os.path.join(HERE, "rules_cases.npz"), **out)


if __name__ == "__main__":
    gen_rules()

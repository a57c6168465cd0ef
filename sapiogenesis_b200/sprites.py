"""Reference-named rules phase: `update_organisms(environment)` and the `Organism` / `DNA` views.

Mirrors the surface of the reference's sprites.py that the world step goes through (sprites.py:1809-1856 and what it
calls). The module-level knobs have the reference's names, units and defaults (sprites.py:37-65); assigning to them
(`sprites.reproduction_limit = 4`, Sapiogenesis.py:82-84) takes effect at the next `update_organisms`, as it does there.
"""
from __future__ import annotations

import math
from typing import Iterator

import torch

from .world import CELL_TYPES, F_IMMORTAL, F_MALADAPTIVE, MAXC, PH_REPRODUCE, PH_RULES, PH_THINK, PH_TRAIN

# -- knobs (sprites.py:37-65), seconds / percent as in the reference ---------------------------------------------------
co2_conversion_ratio = 3
cell_dispersion_rate = 6
reproduction_limit = 6
offspring_amount = 1
heal_rate = 4
mass_cutoff = 1
break_damage = 5
starvation_rate = 6
neural_update_delay = .5
learning_update_delay = .5
training_epochs = 1
max_push_speed = 250
max_rotation_speed = 6
age_limit = 200
eyesight_multiplier = 10
olfactory_multiplier = 15

cell_types = list(CELL_TYPES[:8])                      # sprites.py:22


def _ticks(seconds: float, dt: float) -> int:
    """Virtual-clock form of `time.time() - t0 >= seconds` with one tick == dt seconds."""
    return max(1, math.ceil(seconds / dt - 1e-4))          # dt is stored as fp32


def _sync_knobs(environment) -> None:
    p = environment.world.p
    g = globals()
    want = dict(co2_ratio=float(g["co2_conversion_ratio"]), dispersion_rate=float(g["cell_dispersion_rate"]),
                heal_rate=float(g["heal_rate"]), mass_cutoff=float(g["mass_cutoff"]), break_damage=float(g["break_damage"]),
                starvation_rate=float(g["starvation_rate"]), max_push=float(g["max_push_speed"]),
                max_rot=float(g["max_rotation_speed"]), eyesight_mult=float(g["eyesight_multiplier"]),
                olfactory_mult=float(g["olfactory_multiplier"]), offspring_amount=int(g["offspring_amount"]),
                repro_ticks=_ticks(g["reproduction_limit"], p.dt_wall), age_ticks=_ticks(g["age_limit"], p.dt_wall),
                think_ticks=_ticks(g["neural_update_delay"], p.dt_wall), train_ticks=_ticks(g["learning_update_delay"], p.dt_wall))
    changed = False
    for k, v in want.items():
        if getattr(p, k) != v:
            setattr(p, k, v); changed = True
    if g["training_epochs"] != 1:
        raise NotImplementedError("training_epochs != 1 (the reference never changes it: sprites.py:57)")
    if changed:
        environment.engine.refresh()


def update_organisms(environment) -> None:
    """sprites.py:1809-1856: ageing, Organism.update for every living organism (metabolism, gases, damage, death cascade,
    brain activation, outputs), dopamine training, reproduction, dead-cell dispersion, out-of-bounds handling.
    One call == the rules phase of one tick, on the device."""
    if environment.info["paused"]:
        return
    _sync_knobs(environment)
    e = environment.engine
    e.rules(PH_RULES | PH_THINK | PH_TRAIN | PH_REPRODUCE)
    e.train()
    e.settle()
    e.reproduce()


class DNA:
    """View of an organism's genome as materialised in the world (sprites.py:329-408): the cell table and brain structure."""

    def __init__(self, organism: "Organism"):
        self._o = organism

    @property
    def cells(self) -> dict:
        """cell_id -> cell_info subset (type, size, mass, first parent, relative position), sprites.py:238-253."""
        t = self._o.environment.world.t
        base = self._o.slot * MAXC
        out = {}
        rows = torch.nonzero(t["c_alive"][base:base + MAXC] == 1).view(-1).tolist()
        ids = t["c_id"][base:base + MAXC].tolist()
        for k in rows:
            r = base + k
            par = int(t["c_parent"][r])
            out[ids[k]] = dict(type=CELL_TYPES[int(t["c_type"][r])], size=float(t["c_size"][r]), mass=float(t["c_mass"][r]),
                               first=(ids[par] if par >= 0 else None), relative_pos=t["c_rel"][r].tolist(),
                               mirror_self=(ids[int(t["c_mirror"][r])] if int(t["c_mirror"][r]) >= 0 else None))
        return out

    @property
    def generation(self) -> int:
        return int(self._o.environment.world.t["org_generation"][self._o.slot])

    @property
    def brain_structure(self) -> list:
        t = self._o.environment.world.t
        n = int(t["org_nlayers"][self._o.slot])
        return t["org_ldim"][self._o.slot][:n + 1].tolist()

    @property
    def trainingInput(self):
        t = self._o.environment.world.t
        o = self._o.slot
        n, nin = int(t["org_mem_n"][o]), int(t["org_ldim"][o][0])
        return t["org_mem_in"][o].view(232, -1)[:n, :nin]

    @property
    def curiosity(self) -> float:
        return float(self._o.environment.world.t["org_curiosity"][self._o.slot])

    def export(self) -> dict:
        """The attribute dict of the reference's `sprites.DNA` for this organism (sapiogenesis_b200.dna_io)."""
        from .dna_io import export_dna
        return export_dna(self._o.environment.world, self._o.slot)

    def save(self, path: str) -> None:
        """save_organism_clicked (Sapiogenesis.py:243-254): a pickle the reference's "Import organism" opens."""
        from .dna_io import dumps_reference_pickle
        with open(path, "wb") as f:
            f.write(dumps_reference_pickle(self.export()))


class Organism:
    """View of one organism slot (sprites.py:1089-1125)."""

    def __init__(self, environment, slot: int):
        self.environment, self.slot = environment, int(slot)
        self.dna = DNA(self)

    def _t(self, name):
        return self.environment.world.t[name][self.slot]

    @property
    def brain(self):
        from .neural import Network
        return Network(self)

    def alive(self) -> bool:                             # sprites.py:1127-1135
        return int(self._t("org_state")) == 1

    @property
    def uid(self) -> int:
        return int(self._t("org_uid"))

    @property
    def immortal(self) -> bool:
        return bool(int(self._t("org_flags")) & F_IMMORTAL)

    @immortal.setter
    def immortal(self, v: bool):
        f = self.environment.world.t["org_flags"]
        f[self.slot] = (int(f[self.slot]) | F_IMMORTAL) if v else (int(f[self.slot]) & ~F_IMMORTAL)

    @property
    def maladaptive(self) -> bool:
        return bool(int(self._t("org_flags")) & F_MALADAPTIVE)

    @maladaptive.setter
    def maladaptive(self, v: bool):
        f = self.environment.world.t["org_flags"]
        f[self.slot] = (int(f[self.slot]) | F_MALADAPTIVE) if v else (int(f[self.slot]) & ~F_MALADAPTIVE)

    @property
    def cells(self) -> dict:
        """cell_id -> Sprite for the living cells (sprites.py:1698-1729)."""
        from .physics import Sprite
        t = self.environment.world.t
        base = self.slot * MAXC
        rows = torch.nonzero(t["c_alive"][base:base + MAXC] == 1).view(-1).tolist()
        ids = t["c_id"][base:base + MAXC].tolist()
        return {ids[k]: Sprite(self.environment, "cell", base + k) for k in rows}

    def _living(self):
        t = self.environment.world.t
        base = self.slot * MAXC
        return base, (t["c_alive"][base:base + MAXC] == 1)

    def total_energy(self) -> float:                     # sprites.py:1356-1361
        base, m = self._living()
        return float(self.environment.world.t["c_energy"][base:base + MAXC][m].sum())

    def max_energy(self) -> float:
        base, m = self._living()
        return float(self.environment.world.t["c_storage"][base:base + MAXC][m].sum())

    def health(self) -> float:
        base, m = self._living()
        return float(self.environment.world.t["c_health"][base:base + MAXC][m].sum())

    def max_health(self) -> float:
        base, m = self._living()
        return float(self.environment.world.t["c_maxhealth"][base:base + MAXC][m].sum())

    def health_percent(self) -> float:                   # sprites.py:1145-1156
        mh = self.max_health()
        return 100.0 * self.health() / mh if mh else 0.0

    def energy_percent(self) -> float:
        me = self.max_energy()
        return 100.0 * self.total_energy() / me if me else 0.0

    def position(self):
        """Mean position of the living cells (the quantity neural.activate derives its speed / direction inputs from)."""
        base, m = self._living()
        return self.environment.world.t["c_pos"][base:base + MAXC][m].mean(0).tolist()

    @property
    def movement(self) -> dict:
        """self.movement (sprites.py:1113-1118): the decoded brain outputs the cells act on."""
        mv = self._t("org_move").tolist()
        return dict(rotation=mv[0], speed=mv[1], direction=[mv[2], mv[3]])

    def kill(self):
        """sprites.py:1362-1365: kill the first cell; the cascade takes the rest. Here the heart's health is zeroed and the
        cascade runs in the next update_organisms (kill_cell, sprites.py:1326-1354)."""
        self.environment.world.t["c_health"][self.slot * MAXC] = 0.0

    # -- manual controls of the selected organism (Sapiogenesis.py:209-242, 534-557) -----------------------------------------
    def heal(self):                                      # heal_btn_clicked: every living cell back to max_health
        base, m = self._living()
        t = self.environment.world.t
        t["c_health"][base:base + MAXC][m] = t["c_maxhealth"][base:base + MAXC][m]

    def feed(self):                                      # feed_btn_clicked: every living cell's energy to its storage
        base, m = self._living()
        t = self.environment.world.t
        t["c_energy"][base:base + MAXC][m] = t["c_storage"][base:base + MAXC][m]

    def hurt(self):                                      # hurt_btn_clicked: 10 % of max_health off every living cell, floor 0
        base, m = self._living()
        t = self.environment.world.t
        h = t["c_health"][base:base + MAXC]
        h[m] = torch.clamp(h[m] - 0.10 * t["c_maxhealth"][base:base + MAXC][m], min=0.0)

    @property
    def finite_memory(self) -> bool:
        from .world import F_FINITE_MEMORY
        return bool(int(self._t("org_flags")) & F_FINITE_MEMORY)

    @finite_memory.setter
    def finite_memory(self, v: bool):
        from .world import F_FINITE_MEMORY
        f = self.environment.world.t["org_flags"]
        f[self.slot] = (int(f[self.slot]) | F_FINITE_MEMORY) if v else (int(f[self.slot]) & ~F_FINITE_MEMORY)

    def update(self):
        raise NotImplementedError("organisms are updated together by update_organisms(environment): the canonical tick "
                                  "(DESIGN.md section 2) has no per-organism order")

    def __eq__(self, o):
        return isinstance(o, Organism) and o.environment is self.environment and o.slot == self.slot

    def __hash__(self):
        return hash((id(self.environment), self.slot))


class OrganismList:
    """environment.info["organism_list"] (physics.py:296): the living organisms in slot order."""

    def __init__(self, environment):
        self._env = environment

    def _slots(self):
        return torch.nonzero(self._env.world.t["org_state"] == 1).view(-1).tolist()

    def __len__(self):
        return int((self._env.world.t["org_state"] == 1).sum())

    def __iter__(self) -> Iterator[Organism]:
        for s in self._slots():
            yield Organism(self._env, s)

    def __getitem__(self, i):
        s = self._slots()[i]
        return [Organism(self._env, k) for k in s] if isinstance(s, list) else Organism(self._env, s)


def disperse_all_dead(environment) -> None:
    """sprites.py:1747-1750: remove every dead cell (disperse_cell returns no gas: sprites.py:1733-1745)."""
    environment.world.t["d_state"].zero_()


# -- world-wide controls (Sapiogenesis.py:396-412) -----------------------------------------------------------------------
def feed_all(environment) -> None:
    """feed_all_clicked: every living cell of every organism gets its storage full."""
    t = environment.world.t
    live = (t["c_alive"] == 1) & (t["org_state"].repeat_interleave(MAXC) == 1)
    t["c_energy"][live] = t["c_storage"][live]


def kill_all(environment) -> None:
    """kill_all_clicked: org.kill() for every organism (the cascades run in the next update_organisms)."""
    t = environment.world.t
    hearts = torch.nonzero(t["org_state"] == 1).view(-1) * MAXC
    t["c_health"][hearts] = 0.0


def reset(environment) -> None:
    """reset_clicked: kill everything, clear the dead cells, gases back to 30000 / 0."""
    kill_all(environment)
    update_organisms(environment)                        # the cascades turn every organism into dead cells ...
    disperse_all_dead(environment)                       # ... which are removed
    environment.info["co2"] = 30000
    environment.info["oxygen"] = 0

"""Reference-named host side of the world step: `Environment`, `Space`, `Sprite` views.

Mirrors the simulation-relevant surface of the reference's physics.py so that the code which drives a world in the
reference (Sapiogenesis.py:676-700) reads the same here:

    env = physics.Environment(None, None, 3180, 1800)          # physics.py:280
    env.preUpdateEvent = lambda: sprites.update_organisms(env)  # Sapiogenesis.py:687
    env.update(None)                                            # physics.py:425-442, one timer event

Same names, same argument meaning, same order of effects. The state lives in HBM (`World`), the work is done by
libsapio_b200.so; the Qt arguments (`worldView`, `scene`) are accepted and ignored. Differences forced by the canonical
tick: time is the virtual clock, `update` is synchronous, and `sprites` / `organism_list` are views
over device arrays, not Python object lists.
"""
from __future__ import annotations

import math
from typing import Iterator

import numpy as np
import torch

from .engine import Engine
from .world import MAXC, PH_FRICTION, PH_POST, World, default_params


class Space:
    """The `pymunk.Space` the reference steps (physics.py:287, 432). `step(dt)` is sapio_space_step."""

    def __init__(self, env: "Environment"):
        self._env = env
        self.gravity = (0, 0)
        self.iterations = int(env.world.p.iterations)

    def step(self, dt: float):
        env = self._env
        if dt == 0.0:
            return                                     # paused world: physics.py:428-430 steps by 0.0 (nothing moves)
        if abs(env.world.p.dt_phys - dt) > 0.0 or env.world.p.iterations != self.iterations:
            env.world.p.dt_phys = float(dt)
            env.world.p.iterations = int(self.iterations)
            env.engine.refresh()
        env.engine.space_step()


class Sprite:
    """Read/write view of one body (a living cell row or a free dead cell), physics.py:125-275."""

    __slots__ = ("environment", "_kind", "_i")

    def __init__(self, environment: "Environment", kind: str, index: int):
        self.environment, self._kind, self._i = environment, kind, int(index)

    def _f(self, cell_name, dead_name):
        t = self.environment.world.t
        return t[cell_name if self._kind == "cell" else dead_name]

    @property
    def alive(self) -> bool:
        return self._kind == "cell"

    @property
    def radius(self) -> float:
        return float(self._f("c_size", "d_size")[self._i]) / 2.0

    def getPos(self):                                   # physics.py:203 (body position)
        p = self._f("c_pos", "d_pos")[self._i].tolist()
        return [p[0], p[1]]

    def getCenter(self):                                # physics.py:194
        return self.getPos()

    def getWidth(self):
        return 2.0 * self.radius

    def getHeight(self):
        return 2.0 * self.radius

    @property
    def velocity(self):
        return tuple(self._f("c_vel", "d_vel")[self._i].tolist())

    @velocity.setter
    def velocity(self, v):
        self._f("c_vel", "d_vel")[self._i] = torch.tensor([float(v[0]), float(v[1])])

    @property
    def angle(self) -> float:
        return float(self._f("c_ang", "d_ang")[self._i])

    @property
    def mass(self) -> float:
        return float(self._f("c_mass", "d_mass")[self._i])

    @property
    def info(self) -> dict:
        """The simulation entries of sprite.info (sprites.py:255-327)."""
        t = self.environment.world.t
        i = self._i
        if self._kind == "dead":
            return dict(type="dead", size=float(t["d_size"][i]), mass=float(t["d_mass"][i]), health=0.0)
        from .world import CELL_TYPES
        return dict(type=CELL_TYPES[int(t["c_type"][i])], size=float(t["c_size"][i]), mass=float(t["c_mass"][i]),
                    health=float(t["c_health"][i]), energy=float(t["c_energy"][i]), max_health=float(t["c_maxhealth"][i]),
                    energy_storage=float(t["c_storage"][i]), in_use=bool(t["c_inuse"][i]), cell_id=int(t["c_id"][i]))

    def bump(self, direction, speed, invert=False):     # physics.py:211-217
        d = math.radians(direction)
        k = -1.0 if invert else 1.0
        vel = self._f("c_vel", "d_vel")
        vel[self._i] += torch.tensor([k * math.cos(d) * speed, k * math.sin(d) * speed], device=vel.device)

    # -- collisions and joints (physics.py:219-248) -----------------------------------------------------------------------
    def collision(self, item):                          # physics.py:219-220: a hook the reference leaves empty
        pass

    @property
    def _organism(self):
        return self._i // MAXC if self._kind == "cell" else -1

    @property
    def connections(self) -> dict:
        """physics.py:143 / :222-248: the pin joints of this sprite. In the world step the joints of an organism are implicit -- every
        pair of its LIVING cells is pinned (sprites.py:1721-1726), a joint dies with either of its cells (sprites.py:1340-1343) and a free
        dead cell has none -- so the dict is derived from the living-cell set."""
        if self._kind != "cell":
            return {}
        t = self.environment.world.t
        o, base = self._organism, self._organism * MAXC
        if int(t["c_alive"][self._i]) != 1:
            return {}
        rows = torch.nonzero(t["c_alive"][base: base + int(t["org_ncells"][o])] == 1).view(-1).tolist()
        return {Sprite(self.environment, "cell", base + k): "PinJoint" for k in rows if base + k != self._i}

    def connectTo(self, sprite, rigid=True):            # physics.py:222-241
        """Two living cells of one organism are pinned already (a no-op, like the reference's early return for an existing connection,
        physics.py:223-224). Any other joint is not part of the per-tick world step this library implements: it fails loudly instead of
        being dropped silently."""
        if self._kind == "cell" and sprite._kind == "cell" and self._organism == sprite._organism:
            return
        raise NotImplementedError("pin joints exist between the living cells of one organism only (sprites.py:1721-1726); "
                                  "a joint across organisms or to a free body is outside the world step (SURVEY.md section 8)")

    def disconnectFrom(self, sprite):                   # physics.py:243-248
        """The reference removes a joint when one of its cells dies (sprites.py:1340-1343); the world step does the same inside the death
        cascade. Disconnecting two living cells of an organism by hand is not representable (the joint set is the living-cell set)."""
        if sprite not in self.connections:
            return                                      # physics.py:244: nothing to do
        raise NotImplementedError("the joints of an organism follow its living cells; kill the cell (removeSprite) to remove its joints")

    def __eq__(self, o):
        return isinstance(o, Sprite) and o.environment is self.environment and (o._kind, o._i) == (self._kind, self._i)

    def __hash__(self):
        return hash((self._kind, self._i))

    def updateSprite(self):
        """Per-sprite form of the friction pass (physics.py:250-275); the batch form is Environment.update."""
        f = 1.0 - self.environment.friction / 100.0 * self.environment.world.p.dt_wall
        self._f("c_vel", "d_vel")[self._i] *= f
        self._f("c_angvel", "d_angvel")[self._i] *= f


class _SpriteList:
    """`Environment.sprites` (physics.py:285): every body of the world, living cells first."""

    def __init__(self, env: "Environment"):
        self._env = env

    def _indices(self):
        t = self._env.world.t
        cells = torch.nonzero(t["c_alive"] == 1).view(-1).tolist()
        dead = torch.nonzero(t["d_state"] > 0).view(-1).tolist()
        return cells, dead

    def __len__(self):
        t = self._env.world.t
        return int((t["c_alive"] == 1).sum()) + int((t["d_state"] > 0).sum())

    def __iter__(self) -> Iterator[Sprite]:
        cells, dead = self._indices()
        for i in cells:
            yield Sprite(self._env, "cell", i)
        for i in dead:
            yield Sprite(self._env, "dead", i)


class _Info(dict):
    """`Environment.info` (physics.py:290-311). The world-state keys read through to the device."""

    _LIVE = ("co2", "oxygen", "population", "organism_list")

    def __init__(self, env: "Environment", **kw):
        super().__init__(**kw)
        self._env = env

    def __getitem__(self, k):
        w = self._env.world
        if k == "co2":
            return float(w.t["g_co2"][0])
        if k == "oxygen":
            return float(w.t["g_o2"][0])
        if k == "population":
            return int((w.t["org_state"] == 1).sum())
        if k == "organism_list":
            from .sprites import OrganismList
            return OrganismList(self._env)
        return super().__getitem__(k)

    def __setitem__(self, k, v):
        w = self._env.world
        p = w.p
        if k == "co2":
            w.t["g_co2"][0] = float(v)
        elif k == "oxygen":
            w.t["g_o2"][0] = float(v)
        elif k in ("mutation_severity", "brain_mutation_severity", "learning_rate"):
            setattr(p, k, float(v)); self._env.engine.refresh(); super().__setitem__(k, v)
        elif k == "population_limit":
            p.population_limit = int(v); self._env.engine.refresh(); super().__setitem__(k, v)
        elif k == "reproduction_limit":
            p.offspring_amount = int(v); self._env.engine.refresh(); super().__setitem__(k, v)
        elif k in ("use_rnn", "hidden_rnn_size"):
            # physics.py:305,308, read by build_brain (sprites.py:1683-1690): brains built from now on (spawn, birth, import) are
            # networks.RNNetwork with that hidden size; living organisms keep the brain they have. The size is fixed once set: the
            # reference cannot change it either while RNN organisms live (torch.tensor of ragged trainingHidden rows raises, neural.py:412)
            super().__setitem__(k, v)
            h = int(dict.get(self, "hidden_rnn_size", 6)) if dict.get(self, "use_rnn", False) else 0
            if h:
                w.reserve_rnn(h)
            p.rnn_hidden = h; self._env.engine.refresh()
        elif k == "weight_persistence":
            if not v:
                raise NotImplementedError("weight_persistence=False crashes the reference (SURVEY C-12); not reproduced")
            super().__setitem__(k, v)
        else:
            super().__setitem__(k, v)

    def __contains__(self, k):
        return k in self._LIVE or super().__contains__(k)


class Environment:
    """physics.py:279-471. `capacity` (organism slots) and `dead_capacity` replace the unbounded Python lists."""

    def __init__(self, worldView=None, scene=None, width=3180, height=1800, friction=60.0, gravity=(0, 0), frameWidth=10.0,
                 *, capacity: int = 64, dead_capacity: int | None = None, device="cuda", seed: int = 0, **caps):
        if tuple(gravity) != (0, 0):
            raise NotImplementedError("the reference only ever runs with gravity (0, 0) (physics.py:473)")
        self.worldView, self.scene = worldView, scene
        p = default_params(n_org=int(capacity), n_dead=int(dead_capacity or max(1024, 64 * capacity)), width=float(width),
                           height=float(height), seed=seed, population_limit=caps.pop("population_limit", 12), **caps)
        p.friction_pct = float(friction)
        p.frame_width = float(frameWidth)
        self.world = World(p, device=device)
        self.engine = Engine(self.world)
        self.friction = float(friction)
        self.width, self.height = width, height
        self.default_world_speed = 0.06
        self.worldSpeed = 0.06
        self.updateSpeed = 30                                            # environment_settings.json "update_speed" (ms)
        self.space = Space(self)
        self.sprites = _SpriteList(self)
        self.info = _Info(self, startTime=0.0, selected=None, lastPosition=[0, 0], copied=None, mutation_severity=0.5,
                          brain_mutation_severity=0.5, reproduction_limit=6, population_limit=int(p.population_limit),
                          weight_persistence=True, learning_rate=0.02, use_rnn=False, paused=False, paused_time=0.0,
                          hidden_rnn_size=6, finite_memory=True, last_event_update=0.0)
        self.world.t["g_co2"][0] = 30000.0
        self.lastUpdated = 0.0
        self.mousePressFunc = None
        self.mouseReleaseFunc = None

    # hooks the application assigns (physics.py:402-408)
    def preUpdateEvent(self):
        pass

    def postUpdateEvent(self):
        pass

    def pause_world(self):                               # physics.py:410-416: velocities are kept where they are
        self.info["paused"] = True

    def resume_world(self):
        self.info["paused"] = False

    def update(self, event=None):
        """One timer event (physics.py:425-442): preUpdateEvent -> Space.step(worldSpeed) -> updateSprite for every sprite
        -> postUpdateEvent. Closes the tick (contact bookkeeping for the next tick's rules, tick counter)."""
        if self.info["paused"]:
            self.postUpdateEvent()
            self.space.step(0.0)
            return
        if abs(self.friction - self.world.p.friction_pct) > 0.0:
            self.world.p.friction_pct = float(self.friction); self.engine.refresh()
        self.preUpdateEvent()
        self.space.step(self.worldSpeed)
        self.engine.tick(1, PH_FRICTION | PH_POST)        # every Sprite.updateSprite + the tick's bookkeeping, clock += 1
        self.postUpdateEvent()

    def step(self, n: int = 1):
        """n whole ticks in one call (the fused path; equals n x update() with update_organisms as preUpdateEvent)."""
        self.engine.tick(n)

    # -- world editing (Sapiogenesis.py:297-323, 490-507) ------------------------------------------------------------
    def add_ball_sprite(self, xy, width, height, mass=10, friction=.3, elasticity=.5, image=None, parent=None):
        """physics.py:458-464: a free circle; in this world that is a dead cell (food). Returns its Sprite."""
        t = self.world.t
        free = torch.nonzero(t["d_state"] == 0).view(-1)
        if free.numel() == 0:
            raise RuntimeError("dead-cell capacity exhausted")
        d = int(free[0])
        t["d_state"][d] = 1
        t["d_pos"][d] = torch.tensor([float(xy[0]), float(xy[1])])
        t["d_vel"][d] = 0.0; t["d_ang"][d] = 0.0; t["d_angvel"][d] = 0.0; t["d_vbias"][d] = 0.0; t["d_wbias"][d] = 0.0
        t["d_size"][d] = float(width); t["d_mass"][d] = float(mass); t["d_mass0"][d] = float(mass)
        t["d_elast"][d] = float(elasticity); t["d_fric"][d] = float(friction); t["d_owner"][d] = -1; t["d_eaten"][d] = 0.0
        return Sprite(self, "dead", d)

    def add_creature(self, pos):
        """add_creature_clicked (Sapiogenesis.py:297-323): a random genome at `pos`. Returns the Organism view."""
        from .sprites import Organism
        free = torch.nonzero(self.world.t["org_state"] == 0).view(-1)
        if free.numel() == 0:
            raise RuntimeError("organism capacity exhausted")
        slot = int(free[0])
        self.engine.spawn(np.array([slot], dtype=np.int32), np.array([[float(pos[0]), float(pos[1])]]))
        return Organism(self, slot)

    def import_organism(self, path_or_dna, pos):
        """import_organism_clicked + paste (Sapiogenesis.py:413-424, 325-335): a `sprites.DNA` pickle (or its attribute dict)
        becomes a new organism at `pos`. Returns the Organism view."""
        from .dna_io import import_dna, loads_reference_pickle
        from .sprites import Organism
        dna = path_or_dna
        if not isinstance(dna, dict):
            with open(path_or_dna, "rb") as f:
                dna = loads_reference_pickle(f.read())
        free = torch.nonzero(self.world.t["org_state"] == 0).view(-1)
        if free.numel() == 0:
            raise RuntimeError("organism capacity exhausted")
        slot = int(free[0])
        import_dna(self.world, slot, pos, dna)
        self.engine._grid_valid = False
        return Organism(self, slot)

    def add_sprite(self, sprite: Sprite):                # physics.py:453-456
        """Puts a sprite (back) into the space. Bodies live in the world's tables, so a view made by add_ball_sprite is in the space
        already; a free body taken out with removeSprite comes back with the state its row still holds."""
        if sprite.environment is not self:
            raise ValueError("the sprite belongs to another environment")
        if sprite._kind == "dead":
            if int(self.world.t["d_state"][sprite._i]) == 0:
                self.world.t["d_state"][sprite._i] = 1
                self.engine._grid_valid = False
            return sprite
        if int(self.world.t["c_alive"][sprite._i]) != 1:
            raise NotImplementedError("a cell that died cannot be re-added: it has left through the death cascade (sprites.py:1326-1354)")
        return sprite

    def collision(self, arbiter=None, space=None, data=None):   # physics.py:379-386
        """The reference's collision handler forwards every cell / cell arbiter to the two sprites' (empty) collision hooks; the contact
        adjacency the rules use (sprite.colliding) is built by Space.step on the device, so there is nothing to forward here."""
        return True

    def removeSprite(self, sprite: Sprite):              # physics.py:444-451
        t = self.world.t
        if sprite._kind == "dead":
            t["d_state"][sprite._i] = 0
        else:
            t["c_health"][sprite._i] = 0.0               # a living cell leaves through the death cascade

    def stop(self):
        self.engine.synchronize()


def setupEnvironment(worldView=None, scene=None, friction=60.0, gravity=(0, 0), **kw):
    """physics.py:473-480."""
    width = kw.pop("width", 3180); height = kw.pop("height", 1800)
    return Environment(worldView, scene, width, height, friction=friction, gravity=gravity, **kw)

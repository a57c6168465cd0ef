"""Harness that imports the REFERENCE's own physics.py / sprites.py / neural.py / networks.py headless.

Used only by tests/golden/make_golden.py, in the build container where /root/reference exists; the GPU box
never runs this. PyQt5 and pymunk are not installed anywhere in this image, so they are replaced by inert
stand-ins: Qt classes swallow every call, pymunk bodies/shapes/joints are plain attribute holders and
Space.step does nothing. Everything that runs through these stand-ins is therefore the reference's OWN
Python arithmetic (genome construction, metabolism, gas exchange, death cascade, sensor encoding, brain
forward / training); nothing of Chipmunk's solver is pinned this way.

Every random draw of the reference is redirected to a caller-supplied list of uniforms (UniformFeed), so the
oracle can be fed the identical uniforms. time.time() is a virtual clock.
"""
from __future__ import annotations

import math
import os
import sys
import types

REF = os.environ.get("SAPIO_REFERENCE", "/root/reference")


class UniformFeed:
    def __init__(self, values=None, name="u"):
        self.values = list(values) if values is not None else []
        self.pos = 0
        self.name = name

    def next(self) -> float:
        if self.pos >= len(self.values):
            raise RuntimeError(f"uniform feed {self.name} exhausted at {self.pos}")
        v = float(self.values[self.pos])
        self.pos += 1
        return v


class RandomShim:
    """random.random / randrange / choice / triangular on one uniform each (the oracle's u_* helpers)."""

    def __init__(self, feed: UniformFeed):
        self.feed = feed

    def random(self):
        return self.feed.next()

    def randrange(self, a, b=None):
        if b is None:
            a, b = 0, a
        if b <= a:
            raise ValueError("empty range")
        return a + int(math.floor(self.feed.next() * (b - a)))

    def choice(self, seq):
        return seq[int(math.floor(self.feed.next() * len(seq)))]

    def triangular(self, low, mode, high):
        u = self.feed.next()
        if high == low:
            return low
        c = (mode - low) / (high - low)
        if u > c:
            u, c, low, high = 1.0 - u, 1.0 - c, high, low
        return low + (high - low) * math.sqrt(u * c)


class Clock:
    def __init__(self):
        self.now = 1000.0

    def time(self):
        return self.now


class _Inert:
    """Qt stand-in: accepts any construction, any attribute, any call."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: _Inert()

    def __call__(self, *a, **k):
        return _Inert()


class Vec2d:
    def __init__(self, x=0.0, y=None):
        if y is None:
            x, y = x[0], x[1]
        self.x, self.y = float(x), float(y)

    def __getitem__(self, i): return (self.x, self.y)[i]
    def __iter__(self): return iter((self.x, self.y))
    def __len__(self): return 2
    def __add__(self, o): return Vec2d(self.x + o[0], self.y + o[1])
    __radd__ = __add__
    def __sub__(self, o): return Vec2d(self.x - o[0], self.y - o[1])
    def __repr__(self): return f"Vec2d({self.x}, {self.y})"


class Body:
    def __init__(self, mass=0.0, moment=0.0):
        self.mass, self.moment = mass, moment
        self._position = Vec2d(0, 0)
        self._velocity = Vec2d(0, 0)
        self.angular_velocity = 0.0
        self.force = Vec2d(0, 0)
        self.angle = 0.0
        self.space = None
        self.velocity_func = None

    position = property(lambda s: s._position, lambda s, v: setattr(s, "_position", Vec2d(v[0], v[1])))
    velocity = property(lambda s: s._velocity, lambda s, v: setattr(s, "_velocity", Vec2d(v[0], v[1])))

    def _set_mass(self, m):
        self.mass = m

    @staticmethod
    def update_velocity(body, gravity, damping, dt):
        pass


class _Points:
    def __init__(self, pts):
        self.points = pts


class Circle:
    def __init__(self, body, radius, offset=(0, 0)):
        self.body, self.radius = body, radius
        self.friction = self.elasticity = 0.0
        self.sensor = False
        self.collision_type = 0
        self.sprite = None

    def shapes_collide(self, other):
        # Chipmunk's strict circle/circle test (SURVEY A-3)
        dx = other.body.position[0] - self.body.position[0]
        dy = other.body.position[1] - self.body.position[1]
        md = self.radius + other.radius
        return _Points([1] if dx * dx + dy * dy < md * md else [])


class Segment:
    def __init__(self, body, a, b, radius):
        self.body, self.a, self.b, self.radius = body, a, b, radius
        self.friction = self.elasticity = 0.0


class PinJoint:
    def __init__(self, a, b, anchor_a=(0, 0), anchor_b=(0, 0)):
        self.a, self.b = a, b


class _Handler:
    begin = post_solve = None


class Space:
    def __init__(self):
        self.gravity = (0, 0)
        self.static_body = Body()
        self.items = []

    def add(self, *items):
        for it in items:
            self.items.append(it)
            if isinstance(it, Body):
                it.space = self

    def remove(self, *items):
        for it in items:
            if it in self.items:
                self.items.remove(it)

    def add_collision_handler(self, a, b):
        return _Handler()

    def step(self, dt):
        pass


def _moment_for_circle(mass, inner, outer, offset=(0, 0)):
    return mass * (0.5 * (inner * inner + outer * outer) + offset[0] ** 2 + offset[1] ** 2)


_loaded = None


def load():
    """Import the reference modules behind the stand-ins. Returns a namespace with the modules and the knobs."""
    global _loaded
    if _loaded is not None:
        return _loaded
    import numpy
    import torch

    qt = types.ModuleType("PyQt5")
    for sub in ("QtCore", "QtGui", "QtWidgets"):
        m = types.ModuleType("PyQt5." + sub)
        def _mod_getattr(name, _m=sub):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Inert if name[0].isupper() else _Inert()
        m.__getattr__ = _mod_getattr
        setattr(qt, sub, m)
        sys.modules["PyQt5." + sub] = m
    sys.modules["PyQt5"] = qt
    sys.modules["PyQt5.QtWidgets"].QGraphicsPixmapItem = _Inert
    sys.modules["PyQt5.QtCore"].Qt = _Inert()

    pm = types.ModuleType("pymunk")
    pm.Vec2d, pm.Body, pm.Circle, pm.Segment, pm.PinJoint, pm.Space = Vec2d, Body, Circle, Segment, PinJoint, Space
    pm.moment_for_circle = _moment_for_circle
    sys.modules["pymunk"] = pm
    ui = types.ModuleType("userInterface")
    ui.Ui_MainWindow = object
    sys.modules["userInterface"] = ui

    sys.path.insert(0, REF)
    import networks, neural, physics, sprites   # noqa: E401  (the reference's own files)
    sys.path.pop(0)

    ns = types.SimpleNamespace(networks=networks, neural=neural, physics=physics, sprites=sprites,
                               torch=torch, numpy=numpy, clock=Clock())
    # virtual clock
    tmod = types.SimpleNamespace(time=ns.clock.time)
    for mod in (physics, sprites, neural):
        mod.time = tmod
    sprites.bubble_interval = 1e18          # cosmetic bubbles off (they only consume `random`)
    sprites.last_bubble_creation = ns.clock.now
    _loaded = ns
    return ns


def set_feeds(ns, u: UniformFeed, init: UniformFeed | None = None, fill: UniformFeed | None = None):
    """Route every RNG the reference touches to uniform feeds:
    random.* and numpy.random.choice -> u ; nn.Linear init -> init ; torch.rand -> fill ; torch.randn -> fill (Box-Muller)."""
    torch, numpy = ns.torch, ns.numpy
    shim = RandomShim(u)
    ns.sprites.random = shim
    ns.neural.random = shim
    ns.physics.random = shim

    class _NpRandom:
        @staticmethod
        def choice(a, p=None):
            x = u.next()
            cdf = numpy.cumsum(p)
            return a[int(numpy.searchsorted(cdf / cdf[-1], x, side="right"))]

    ns.sprites.numpy = types.SimpleNamespace(random=_NpRandom, arange=numpy.arange, linspace=numpy.linspace)

    if init is not None:
        def reset_parameters(self):
            fan_in = self.weight.shape[1]
            bound = 1.0 / math.sqrt(fan_in)
            with torch.no_grad():
                w = [(2.0 * init.next() - 1.0) * bound for _ in range(self.weight.numel())]
                self.weight.copy_(torch.tensor(w, dtype=torch.float64).reshape(self.weight.shape).float())
                if self.bias is not None:
                    b = [(2.0 * init.next() - 1.0) * bound for _ in range(self.bias.numel())]
                    self.bias.copy_(torch.tensor(b, dtype=torch.float64).float())
        torch.nn.Linear.reset_parameters = reset_parameters
    if fill is not None:
        real_tensor = torch.tensor

        def rand(*shape):
            shape = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else shape
            n = int(numpy.prod(shape)) if len(shape) else 1
            return real_tensor([fill.next() for _ in range(n)], dtype=torch.float32).reshape(shape)
        ns.neural.torch = _TorchProxy(torch, rand=rand)


class _TorchProxy:
    def __init__(self, torch, **over):
        self._t, self._o = torch, over

    def __getattr__(self, name):
        if name in self._o:
            return self._o[name]
        return getattr(self._t, name)


class FakeEnv:
    """physics.Environment built without Qt: same attributes, same add/remove bookkeeping (physics.py:279-469)."""

    def __new__(cls, ns, width=3180.0, height=1800.0):
        cwd = os.getcwd()
        os.chdir(REF)
        try:
            env = ns.physics.Environment(_Inert(), _Inert(), width, height)
        finally:
            os.chdir(cwd)
        env.lastUpdated = ns.clock.now
        return env

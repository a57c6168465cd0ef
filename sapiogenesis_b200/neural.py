"""Reference-named brain entry points: `setup_network`, `activate`, `train_network` (neural.py:91, 369, 426).

In the reference these are called per organism from `Organism.update` / `update_organisms`. The world step calls them in
batch (sprites.update_organisms -> sapio_rules / sapio_train); the functions here are the same operations for one organism
on demand -- the call a tool or a test makes -- and run on the device through sapio_activate / sapio_train_network.
"""
from __future__ import annotations

import torch

from .world import MAXC


class Network:
    """View of an organism's brain (networks.py:137-230): ragged Linear stack stored as one parameter blob."""

    def __init__(self, organism):
        self.organism = organism

    @property
    def structure(self) -> list:
        return self.organism.dna.brain_structure

    def layers(self):
        """[(weight [out, in], bias [out])] views into HBM, input layer first (networks.py:226-230)."""
        o = self.organism
        blob = o.environment.world.t["org_brain"][o.slot]
        dims = self.structure
        out, off = [], 0
        for l in range(len(dims) - 1):
            fin, fout = dims[l], dims[l + 1]
            out.append((blob[off:off + fout * fin].view(fout, fin), blob[off + fout * fin:off + fout * fin + fout]))
            off += fout * fin + fout
        return out

    @property
    def lastUpdated(self) -> int:
        return int(self.organism.environment.world.t["org_think_tick"][self.organism.slot])

    @property
    def lastTrained(self) -> int:
        return int(self.organism.environment.world.t["org_train_tick"][self.organism.slot])

    @property
    def loss(self) -> float:
        return float(self.organism.environment.world.t["org_loss"][self.organism.slot])


def setup_network(dna, learning_rate=0.02, rnn=False, hiddenSize=6) -> Network:
    """neural.py:426-454. The parameters of a materialised organism already exist (build_brain runs inside sapio_spawn /
    sapio_reproduce with the reference's initialisation); this returns the view. `rnn=True` is not built: the reference's
    RNN path is unreachable from the UI default and crashes on `activate` (SURVEY C-11)."""
    if rnn:
        raise NotImplementedError("use_rnn=True is outside the world step (SURVEY C-11)")
    env = dna._o.environment
    if abs(env.world.p.learning_rate - learning_rate) > 0:
        env.world.p.learning_rate = float(learning_rate); env.engine.refresh()
    return Network(dna._o)


def activate(network: Network, environment, organism, uDiff=None) -> list:
    """neural.py:91-366: read the sensors, run the network, apply the boredom gate, record history / short-term memory and
    stimulation, store the movement outputs. Returns the four outputs [rotation, speed, x_direction, y_direction]
    the organism will act on (clamped as neural.py:320-349 stores them)."""
    environment.engine.activate([organism.slot])
    return environment.world.t["org_move"][organism.slot].tolist()


def train_network(organism, epochs=1) -> None:
    """neural.py:369-423: fold the short-term memories into the curated training set (1-decimal keys, better reward
    wins, finite-memory trim) and run one Adam epoch on the input and output layers."""
    if epochs != 1:
        raise NotImplementedError("epochs != 1 (sprites.training_epochs is 1 in the reference)")
    organism.environment.engine.train_network([organism.slot])


def calculateStimulation(dna) -> float:
    """neural.py:60-75 as last evaluated by activate."""
    o = dna._o
    return float(o.environment.world.t["org_stim_hist"][o.slot][9])

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

    @property
    def hiddenSize(self) -> int:
        """RNNetwork.hiddenSize (networks.py:119); 0 for a feed-forward networks.Network"""
        return int(self.organism.environment.world.t["org_rnn_h"][self.organism.slot])

    @property
    def is_rnn(self) -> bool:
        return self.hiddenSize > 0

    @property
    def lastHidden(self):
        """RNNetwork.lastHidden (networks.py:149,174): the hidden state the next forward pass is fed"""
        o = self.organism
        return o.environment.world.t["org_last_hidden"][o.slot][: self.hiddenSize]

    def layers(self):
        """[(weight [out, in], bias [out])] views into HBM, input layer first (networks.py:226-230); for an RNN brain the input layer
        has hiddenSize more columns and the hidden head outputLayerH comes last (RNNetwork.layers, networks.py:127-134)."""
        o = self.organism
        blob = o.environment.world.t["org_brain"][o.slot]
        dims, H = self.structure, self.hiddenSize
        shapes = [(dims[l] + (H if l == 0 else 0), dims[l + 1]) for l in range(len(dims) - 1)] + ([(dims[-2], H)] if H else [])
        out, off = [], 0
        for fin, fout in shapes:
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
    sapio_reproduce with the reference's initialisation, as an RNNetwork when Environment.info["use_rnn"] was set at that moment);
    this returns the view. `rnn` / `hiddenSize` must describe the brain the organism has."""
    env = dna._o.environment
    have = int(env.world.t["org_rnn_h"][dna._o.slot])
    if bool(rnn) != (have > 0) or (rnn and int(hiddenSize) != have):
        raise ValueError(f"organism {dna._o.slot} was built with {'an RNN brain of hidden size %d' % have if have else 'a feed-forward brain'}: "
                         "set Environment.info['use_rnn'] / ['hidden_rnn_size'] before it is spawned or born")
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

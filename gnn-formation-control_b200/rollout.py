"""Batched rollout front-end (BASELINE config 4).

The reference steps a simulation like this (scene.py:385-394 -> robot.py:654 -> suhaas_agent.py:30-57): at every
step each of the N robots calls ``Scene.readADjMatrix`` (an O(N^2) Python loop, scene.py:140-154), ships its own copy
of the — identical — adjacency to the GPU, and runs the whole policy at batch 1 to keep ONE column of the output
(``outs[index]``).  That is N GSO builds + N batch-1 forwards per scene and step, for one distinct result.

``Rollout`` replaces that caller pattern for the graph-filter stage: positions of B parallel scenes stay on the
device, and one ``step(pos, x)`` = ONE pass over the L stacked filter layers for all B scenes and all N robots, with
the GSO rebuilt on chip from the positions inside each layer's kernel (never materialised) and layer l+1 reading
layer l's node-major output in place.  ``out[b, :, i]`` is what robot i of scene b computes in the reference.

For fixed shapes the step is captured once into a CUDA graph (``graph=True``) and replayed: a 16384-scene step is a
few hundred microseconds of GPU work, so Python / launch overhead would otherwise be visible.
"""
import torch
import torch.nn as nn

from . import _cabi as C
from .graph_filter import GraphFilterBatch

_ACT_OF = {nn.LeakyReLU: "leaky_relu", nn.ReLU: "relu"}


class Rollout(nn.Module):
    """``Rollout(layers, radius=2.0, mode="binary_le")`` — inference front-end over stacked graph-filter layers.

    ``layers``: ``gnnfc.GraphFilterBatch`` modules (activation fused or ``None``), or use :meth:`from_gfl` to adopt a
    reference-style ``nn.Sequential`` of ``GraphFilterBatch`` + ``LeakyReLU`` / ``ReLU`` pairs (suhaas_model.py:112-123,
    decentralplanner.py:215-221)."""

    def __init__(self, layers, radius=2.0, mode="binary_le", graph=False):
        super().__init__()
        assert mode in C.GSO_MODES
        self.layers = nn.ModuleList(layers)
        for a, b in zip(self.layers[:-1], self.layers[1:]):
            assert a.F == b.G, "layer widths do not chain: %d -> %d" % (a.F, b.G)
        self.radius, self.mode = float(radius), mode
        self.use_graph = bool(graph)
        self._g = None          # (key, CUDAGraph, static pos, static x, static out)

    @classmethod
    def from_gfl(cls, gfl, radius=2.0, mode="binary_le", graph=False, device=None):
        """adopt a GFL ``nn.Sequential`` (reference ``gml.GraphFilterBatch`` or ``gnnfc.GraphFilterBatch`` modules, each
        optionally followed by ``nn.LeakyReLU`` / ``nn.ReLU``): taps are copied, the activation moves into the
        filter's epilogue."""
        mods = list(gfl)
        layers, i = [], 0
        while i < len(mods):
            m = mods[i]
            assert hasattr(m, "weight") and m.weight.dim() == 4, "expected a graph-filter layer at position %d" % i
            F_, E, K, G = m.weight.shape
            act, slope = getattr(m, "activation", None), getattr(m, "negative_slope", 0.01)
            if i + 1 < len(mods) and type(mods[i + 1]) in _ACT_OF:
                assert act is None, "layer %d already has a fused activation" % i
                act = _ACT_OF[type(mods[i + 1])]
                slope = getattr(mods[i + 1], "negative_slope", 0.0)
                i += 1
            lay = GraphFilterBatch(G, F_, K, E, bias=m.bias is not None, activation=act, negative_slope=slope)
            with torch.no_grad():
                lay.weight.copy_(m.weight)
                if m.bias is not None:
                    lay.bias.copy_(m.bias)
            layers.append(lay)
            i += 1
        r = cls(layers, radius, mode, graph)
        return r.to(device) if device is not None else r

    # ------------------------------------------------------------------------------------------
    def _run(self, pos, x):
        y = x
        for lay in self.layers:
            lay.addPositions(pos, self.radius, self.mode)
            y = lay(y)
        return y

    @torch.no_grad()
    def step(self, pos, x):
        """pos [B,N,2] (device-resident positions of B scenes), x [B,G,N] -> [B,F,N] (view over node-major memory)."""
        assert pos.dim() == 3 and pos.shape[2] == 2 and x.dim() == 3 and x.shape[0] == pos.shape[0]
        assert x.shape[2] == pos.shape[1] and x.shape[1] == self.layers[0].G
        if not self.use_graph:
            return self._run(pos, x)
        key = (tuple(pos.shape), tuple(x.shape), x.device)
        if self._g is None or self._g[0] != key:
            spos = torch.empty(pos.shape, dtype=torch.float32, device=x.device)
            sx = torch.empty(x.shape, dtype=torch.float32, device=x.device)
            spos.copy_(pos); sx.copy_(x)
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._run(spos, sx)
            torch.cuda.current_stream(x.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._run(spos, sx)
            self._g = (key, g, spos, sx, out)
        _, g, spos, sx, out = self._g
        spos.copy_(pos, non_blocking=True)
        sx.copy_(x, non_blocking=True)
        g.replay()
        return out

    def rollout(self, pos0, x_fn, move_fn, steps):
        """convenience loop: ``y_t = step(pos_t, x_fn(t, pos_t)); pos_{t+1} = move_fn(t, pos_t, y_t)``; everything stays
        on the device.  Returns the final positions and the last output."""
        pos, y = pos0, None
        for t_ in range(steps):
            y = self.step(pos, x_fn(t_, pos))
            pos = move_fn(t_, pos, y)
        return pos, y

"""Recorded-data formats feeding the filter path.

The reference records, per robot, one row per time step and concatenates the robots' recordings
(``data.py:43,150-151``: key ``graph`` = ``[T, N*N]`` float32 per robot; ``RobotDataset`` receives the
robot-major concatenation ``[N*T, N*N]``).  ``RobotDataset.__init__`` (``custom_dataset.py:15-63``) re-lays this
out with Python loops into ``[T, N, ...]``; the agent then takes robot 0's view of the graph
(``suhaas_agent.py:117``) and the model adds the edge-feature axis (``suhaas_model.py:155``).  These helpers do the
same re-layout as one strided view / one copy, on whatever device the data lives on."""
import numpy as np
import torch


def robot_major_to_batch(rows, nA, item_shape=None):
    """``[nA*T, ...]`` robot-major rows (row ``j*T + i`` = robot j at step i) -> ``[T, nA, *item_shape]``
    (``c[i, j] = rows[j*T + i].reshape(item_shape)``, custom_dataset.py:15-63)."""
    t = torch.as_tensor(rows)
    assert t.shape[0] % nA == 0, "row count must be a multiple of the number of robots"
    T = t.shape[0] // nA
    tail = tuple(item_shape) if item_shape is not None else tuple(t.shape[1:])
    return t.reshape((nA, T) + tail).transpose(0, 1)


def graphs_from_recording(graph_rows, nA):
    """the ``graph`` recordings ``[nA*T, nA*nA]`` -> ``[T, nA (view of robot j), nA, nA]`` (custom_dataset.py:38-45)"""
    return robot_major_to_batch(graph_rows, nA, (nA, nA))


def gso_batch_from_recording(graph_rows, nA, view=0, device=None, dtype=torch.float32):
    """GSO batch for ``GraphFilterBatch.addGSO``: robot ``view``'s adjacency at every step as a contiguous
    ``[T, 1, nA, nA]`` tensor (suhaas_agent.py:117 ``[:, 0, :, :]`` + suhaas_model.py:155 ``unsqueeze(1)``)."""
    g = graphs_from_recording(graph_rows, nA)[:, view]
    g = g.to(device=device, dtype=dtype) if device is not None else g.to(dtype)
    return g.unsqueeze(1).contiguous()


def positions_from_recording(position_list, nA):
    """``positionList_*.npy`` episodes ``[E, steps*nA, 2]`` with index ``t*nA + r`` (test5_more_robots.py:176-178)
    -> positions ``[E*steps, nA, 2]`` float32, ready for ``addPositions`` / ``gfc_gso_build``."""
    p = torch.as_tensor(np.asarray(position_list))
    E, M, two = p.shape
    assert two == 2 and M % nA == 0
    return p.reshape(E * (M // nA), nA, 2).to(torch.float32).contiguous()


class RecordingLoader:
    """Device-side replacement for ``RobotDataset`` + ``DataLoader(trainset, batch_size=16, shuffle=True,
    drop_last=True)`` as the agent uses them (custom_dataset.py:15-82, suhaas_agent.py:110-121).

    The reference re-lays every recording out with per-element Python loops into fresh float64 arrays on the host and
    then ships each batch to the GPU piece by piece.  Here the robot-major recordings are uploaded ONCE (float32, as
    recorded — data.py:43 declares them float32), the ``[T, nA, ...]`` layout is a strided view, and a batch is one
    ``index_select`` per key on the device; nothing touches the host inside the epoch.  Batches carry the reference's
    keys and shapes: ``data [b,nA,inW,inH]``, ``graphs [b,nA,nA,nA]``, ``actions [b,nA,2]``, ``refs [b,nA,1]``,
    ``alphas [b,nA,1]`` in ``dtype`` (float64 like ``transform=ToTensor`` + ``.double()``), plus ``S`` =
    ``graphs[:, view]`` — the ``[b,nA,nA]`` batch ``model.addGSO`` takes (suhaas_agent.py:117,121)."""

    KEYS = ("data", "actions", "graphs", "refs", "alphas")

    def __init__(self, observations, actions, graphs, refs, alphas, nA, inW=100, inH=100, batch_size=16,
                 shuffle=True, drop_last=True, device="cuda", dtype=torch.float64, view=0, generator=None):
        self.nA, self.batch_size, self.shuffle, self.drop_last = int(nA), int(batch_size), shuffle, drop_last
        self.device, self.dtype, self.view = torch.device(device), dtype, int(view)
        self.generator = generator

        def up(a):
            t = torch.as_tensor(np.asarray(a))
            return t.to(device=self.device, dtype=torch.float32 if t.is_floating_point() else t.dtype)

        self.cols = dict(
            data=robot_major_to_batch(up(observations), nA, (inW, inH)),
            actions=robot_major_to_batch(up(actions), nA, (2,)),
            graphs=robot_major_to_batch(up(graphs), nA, (nA, nA)),
            refs=robot_major_to_batch(up(refs), nA, (1,)),
            alphas=robot_major_to_batch(up(alphas), nA, (1,)))
        self.T = self.cols["data"].shape[0]
        for k in self.KEYS:
            assert self.cols[k].shape[0] == self.T, "recordings of different lengths"

    @classmethod
    def from_recordings(cls, per_robot, nA=None, **kw):
        """``per_robot``: one mapping per robot with the keys ``data.py`` records (``observations``, ``actions``,
        ``graph``, ``obs2`` — columns 1 and 2 of ``obs2`` are the reference angle and alpha, suhaas_agent.py:103-104),
        e.g. ``np.load('data/data001_0.npz')``; robots are concatenated robot-major like ``Data.append`` does."""
        nA = len(per_robot) if nA is None else nA
        cat = lambda key: np.concatenate([np.asarray(d[key]) for d in per_robot], axis=0)   # noqa: E731
        obs2 = cat("obs2")
        return cls(cat("observations"), cat("actions"), cat("graph"), obs2[:, 1], obs2[:, 2], nA, **kw)

    @classmethod
    def from_npz(cls, paths, **kw):
        return cls.from_recordings([np.load(p) for p in paths], **kw)

    def __len__(self):
        n = self.T // self.batch_size
        return n if self.drop_last or self.T % self.batch_size == 0 else n + 1

    def batch(self, idx):
        idx = torch.as_tensor(idx, device=self.device, dtype=torch.long)
        out = {k: self.cols[k].index_select(0, idx).to(self.dtype) for k in self.KEYS}
        out["S"] = out["graphs"][:, self.view]
        return out

    def __iter__(self):
        if self.shuffle:
            order = torch.randperm(self.T, generator=self.generator).to(self.device)
        else:
            order = torch.arange(self.T, device=self.device)
        for i in range(len(self)):
            yield self.batch(order[i * self.batch_size:(i + 1) * self.batch_size])

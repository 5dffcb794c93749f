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

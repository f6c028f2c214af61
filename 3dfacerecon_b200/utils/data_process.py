"""Label wire format of the reference: one text file per sample, ``label_dim`` (= 235) lines of ``%.6f``
(written by ``prepare_data/script_generate_dataset.m:114-125``, read by ``utils/data_process.py:38-60``), batched as
``[batch, 1, 1, label_dim]`` -- the shape ``FaceRecNet`` feeds to ``vertices_transform`` / ``get_loss``.
Only the parameter-label side of the data pipeline is mirrored; images and list files are outside the hot path.
"""
from __future__ import annotations

import os

import numpy as np


def prepare_input_label(label_files, batch_size, label_dim):
    """``utils/data_process.py:38-60``: same arguments, same ``[batch_size, 1, 1, label_dim]`` float64 result, same
    exceptions (``FileNotFoundError`` for a missing file, ``IOError`` for an unreadable one or a wrong length)."""
    assert len(label_files) == batch_size
    input_label = np.zeros([batch_size, 1, 1, label_dim])
    for i in range(batch_size):
        if not os.path.exists(label_files[i]):
            raise FileNotFoundError(label_files[i])
        try:
            labels = np.loadtxt(label_files[i])
        except Exception as exc:
            raise IOError("cannot parse %s" % label_files[i]) from exc
        if labels.ndim == 1 and labels.shape[0] == label_dim:
            input_label[i, 0, 0, :] = labels
        else:
            raise IOError("%s holds %s values, expected %d" % (label_files[i], labels.shape, label_dim))
    return input_label


def write_label(path, params):
    """One label file as the MATLAB generator writes it: one ``%.6f`` per line, pose | shape | expression."""
    params = np.asarray(params, np.float64).reshape(-1)
    with open(path, "w") as f:
        for v in params:
            f.write("%.6f\n" % v)

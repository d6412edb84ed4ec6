"""ORACLE (test infrastructure only) -- CPU restatement of the reference's stage-2 heads,
`/root/reference/Code/sage+gat+diffpool/train_triplet.py`:
  * `evaluate` :80-85  -- KNeighborsClassifier(n_neighbors=3) on the embeddings (sklearn: Euclidean, uniform
    weights; predict = scipy.stats.mode of the neighbours' labels => the smallest label among equal counts);
  * `evaluate_mlp` :148-165 -- nn.Sequential(Linear(in,64), LeakyReLU, Linear(64,32), LeakyReLU, Linear(32,2)),
    Adam(lr=1e-3), one embedding per step: forward, cross_entropy, backward, step, zero_grad.
PARITY: the kNN restatement is cross-checked against sklearn itself in tests/test_oracle_heads.py; the MLP loop is
torch's own modules and optimiser (nothing restated but the loop)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def knn_predict(train: np.ndarray, labels: np.ndarray, query: np.ndarray, k: int = 3) -> np.ndarray:
    d = ((query[:, None, :].astype(np.float64) - train[None, :, :].astype(np.float64)) ** 2).sum(-1)
    out = np.zeros(query.shape[0], np.int64)
    for q in range(query.shape[0]):
        order = np.lexsort((np.arange(train.shape[0]), d[q]))[:k]         # distance, then index
        cnt = np.bincount(labels[order])
        out[q] = int(np.argmax(cnt))                                       # first maximum = smallest label
    return out


def mlp1_train(emb: torch.Tensor, labels: torch.Tensor, model: nn.Sequential, lr: float = 1e-3):
    """train_triplet.py:153-165 on CPU tensors.  Returns the per-step losses."""
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    losses = []
    for i in range(emb.size(0)):
        pred = model(emb[i]).unsqueeze(0)
        loss = F.cross_entropy(pred, labels[i:i + 1])
        loss.backward()
        opt.step()
        opt.zero_grad()
        losses.append(float(loss))
    return losses


def make_model(in_feat: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(in_feat, 64), nn.LeakyReLU(), nn.Linear(64, 32), nn.LeakyReLU(), nn.Linear(32, 2))

"""K12 stage-2 heads: kNN predictions identical to the oracle (and to sklearn on tie-free data); the one-sample-per-
step MLP trainer within 1e-5 of torch's own Adam loop (per-step losses and final parameters)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import heads_ref as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,Q,D,classes,k", [(300, 120, 32, 2, 3), (1000, 257, 64, 3, 5), (5, 9, 8, 2, 3), (2, 3, 4, 2, 3)])
def test_knn(cuda, M, Q, D, classes, k):
    from tsg import heads
    rng = np.random.default_rng(M + Q)
    train = rng.normal(size=(M, D)).astype(np.float32)
    labels = rng.integers(0, classes, size=M)
    query = rng.normal(size=(Q, D)).astype(np.float32)
    query[::7] = train[rng.integers(0, M, size=query[::7].shape[0])]            # exact hits (distance 0)
    train[1] = train[0]; labels[1] = 1 - labels[0] if classes == 2 else labels[1]  # an exact distance tie
    ref = H.knn_predict(train, labels, query, min(k, M))
    got = heads.knn_predict(torch.from_numpy(train).to(cuda), torch.from_numpy(labels).to(cuda),
                            torch.from_numpy(query).to(cuda), k=k, num_classes=classes)
    assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.parametrize("N,D", [(64, 32), (300, 192)])
def test_mlp1_trainer(cuda, N, D):
    from tsg import heads
    g = torch.Generator().manual_seed(N)
    emb = torch.randn(N, D, generator=g)
    labels = (torch.rand(N, generator=g) < 0.5).long()
    torch.manual_seed(5)
    ref_model = H.make_model(D)
    torch.manual_seed(5)
    clf = heads.Mlp1Classifier(D, cuda)                   # same RNG stream => same initial weights
    for a, b in zip(clf.views(), ref_model.parameters()):
        assert torch.equal(a.cpu(), b.detach())
    ref_losses = H.mlp1_train(emb, labels, ref_model)
    losses = clf.fit(emb.to(cuda), labels.to(cuda), want_losses=True)
    assert rel_err(losses, torch.tensor(ref_losses)) <= 1e-4
    for a, b in zip(clf.views(), ref_model.parameters()):
        assert rel_err(a, b) <= 2e-4
    # first steps (before rounding differences can be amplified by Adam's normalisation) hold 1e-5
    assert rel_err(losses[:8], torch.tensor(ref_losses[:8])) <= 1e-5
    q = torch.randn(50, D, generator=g)
    assert torch.equal(clf.predict(q.to(cuda)).cpu(), ref_model(q).argmax(1))
    # a second fit continues the Adam step count
    clf.fit(emb.to(cuda)[:5], labels.to(cuda)[:5])
    assert clf.step == N + 5

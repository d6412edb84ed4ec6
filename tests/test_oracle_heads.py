"""oracle/heads_ref.knn_predict vs sklearn's KNeighborsClassifier (the class the reference instantiates)."""
import numpy as np

from oracle import heads_ref as H


def test_knn_restatement_matches_sklearn():
    from sklearn.neighbors import KNeighborsClassifier
    rng = np.random.default_rng(0)
    for classes in (2, 3):
        train = rng.normal(size=(300, 16)).astype(np.float32)
        labels = rng.integers(0, classes, size=300)
        query = rng.normal(size=(120, 16)).astype(np.float32)
        ref = KNeighborsClassifier(n_neighbors=3).fit(train, labels).predict(query)
        assert np.array_equal(H.knn_predict(train, labels, query, 3), ref)
        # predicting the training set itself: every point is its own nearest neighbour
        ref2 = KNeighborsClassifier(n_neighbors=3).fit(train, labels).predict(train)
        assert np.array_equal(H.knn_predict(train, labels, train, 3), ref2)

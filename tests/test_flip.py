"""Sanity of the LDR-FLIP restatement used by the image parity tests (tests/flip.py): the fixed points and orderings the published
measure guarantees by construction."""
import numpy as np

import flip


def test_flip_fixed_points_and_ordering():
    rng = np.random.default_rng(0)
    a = rng.random((48, 64, 3))
    assert flip.mean_flip(a, a) == 0.0
    green, blue = np.tile([0.0, 1.0, 0.0], (48, 64, 1)), np.tile([0.0, 0.0, 1.0], (48, 64, 1))
    assert abs(flip.mean_flip(green, blue) - 1.0) < 1e-9            # the largest colour difference of the measure maps to 1
    bw = flip.mean_flip(np.zeros((48, 64, 3)), np.ones((48, 64, 3)))
    assert 0.95 < bw < 1.0                                           # above the 0.95 knee of the remapping, below green-blue
    small = np.clip(a + rng.normal(0, 0.004, a.shape), 0, 1)
    large = np.clip(a + rng.normal(0, 0.04, a.shape), 0, 1)
    fs, fl = flip.mean_flip(a, small), flip.mean_flip(a, large)
    assert 0.0 < fs < fl < 1.0
    assert abs(flip.mean_flip(a, small) - flip.mean_flip(small, a)) < 1e-12   # symmetric
    m = flip.flip_map(a, large)
    assert m.shape == a.shape[:2] and m.min() >= 0.0 and m.max() <= 1.0


def test_flip_sees_a_shifted_edge_more_than_a_flat_offset_of_the_same_rmse():
    """The feature pipeline amplifies differences that move edges: a one-pixel shift of a step edge scores higher than a uniform
    offset with the same RMSE."""
    img = np.zeros((48, 64, 3))
    img[:, 32:] = 0.8
    shifted = np.roll(img, 1, axis=1)
    rmse = np.sqrt(((img - shifted) ** 2).mean())
    offset = np.clip(img + rmse, 0, 1)
    assert flip.flip_map(img, shifted).max() > flip.flip_map(img, offset).max()

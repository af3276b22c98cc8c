"""App-faithful MNIST pieces around the hot path ([REF] mnist/src/mnist.cpp): the image-index stream of one
evaluation and the IDX dataset format.  Host-only (the C ABI function needs no GPU)."""
import gzip
import os
import struct

import numpy as np
import pytest

import tpgx


@pytest.mark.parametrize("seed", [0, 1, 12345, 2 ** 63 + 7])
@pytest.mark.parametrize("size", [60000, 10000, 7, 1])
def test_index_stream_matches_libstdcxx_rng(oracle_mod, seed, size):
    """TRAINING / VALIDATION: the i-th classified image is the i-th draw of Mutator::RNG::getUnsignedInt64(0, size-1)
    after setSeed(seed) — the oracle draws through std::uniform_int_distribution over std::mt19937_64."""
    n = 501  # maxNbActionsPerEval = 500 actions consume 501 draws ([REF] mnist/params.json:15)
    ref = np.zeros(n, dtype=np.uint64)
    oracle_mod.load().orc_rng_u64(seed, n, 0, size - 1, ref.ctypes.data)
    for mode in (0, 1):
        got = tpgx.mnist_episode_indices(seed, mode, size, n)
        assert np.array_equal(got.astype(np.uint64), ref)


def test_testing_mode_is_sequential():
    assert np.array_equal(tpgx.mnist_episode_indices(99, 2, 7, 20), np.arange(20) % 7)


def test_bad_arguments():
    with pytest.raises(tpgx.TpgxError):
        tpgx.mnist_episode_indices(0, 3, 10, 5)
    with pytest.raises(tpgx.TpgxError):
        tpgx.mnist_episode_indices(0, 0, 0, 5)


def test_idx_loader(tmp_path):
    img, lab = tpgx.synthetic_images(37, seed=4)
    pi, pl = os.path.join(str(tmp_path), "train-images-idx3-ubyte"), os.path.join(str(tmp_path), "train-labels-idx1-ubyte.gz")
    with open(pi, "wb") as f:
        f.write(struct.pack(">BBBBIII", 0, 0, 8, 3, 37, 28, 28) + img.tobytes())
    with gzip.open(pl, "wb") as f:
        f.write(struct.pack(">BBBBI", 0, 0, 8, 1, 37) + lab.tobytes())
    assert np.array_equal(tpgx.load_idx(pi), img) and np.array_equal(tpgx.load_idx(pl), lab)
    with open(pi, "wb") as f:
        f.write(struct.pack(">BBBBIII", 0, 0, 8, 3, 38, 28, 28) + img.tobytes())
    with pytest.raises(ValueError):
        tpgx.load_idx(pi)

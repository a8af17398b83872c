"""Host-side utilities mirroring code/includes/utils.py: Gumbel noise, clustering accuracy, label generators,
dataset loading and the batching iterators.

The batching iterators keep the reference semantics (reshuffle every epoch, tail batch kept) but hold the
data in ONE pinned host array and yield contiguous slices, so the training loop can issue an asynchronous
host->device copy per batch instead of the reference's per-row Python append (utils.py:449-463).
"""
from __future__ import annotations

import math
import os
from typing import Optional

import numpy as np
import torch
from scipy.optimize import linear_sum_assignment


def sample_gumbel(shape, eps=1e-20):
    """includes/utils.py:17-19."""
    U = np.random.uniform(0, 1, shape)
    return -np.log(eps - np.log(U + eps))


def hungarian_accuracy(d: np.ndarray, size: int) -> float:
    """Accuracy under the best cluster->class assignment of contingency matrix d (utils.py:33-34;
    sklearn's removed linear_assignment_ replaced by scipy's equivalent linear_sum_assignment)."""
    r, c = linear_sum_assignment(d.max() - d)
    return float(d[r, c].sum()) / (size * 1.0)


def get_clustering_accuracy(weights, classes):
    """includes/utils.py:22-34.  The contingency matrix is sized to hold every class label too (the reference sizes it
    by the number of clusters only and raises IndexError when there are fewer clusters than classes, e.g. the
    `--n_experts 4` runs of runOur.sh on 10-class data); identical result whenever the reference's succeeds."""
    weights = np.asarray(weights)
    clusters = np.argmax(weights, axis=-1)
    classes = np.asarray(classes)
    n_classes = max(weights.shape[1], int(classes.max()) + 1 if classes.size else 0)
    size = len(clusters)
    d = np.zeros((n_classes, n_classes), dtype=np.int64)
    np.add.at(d, (clusters, np.asarray(classes).astype(np.int64)), 1)
    return hungarian_accuracy(d, size)


def generate_regression_variable(dataset, output_dim):
    """includes/utils.py:37-60."""
    n_experts = dataset.n_classes
    input_dim = dataset.train_data.shape[1]
    biases = np.random.randn(output_dim, n_experts)
    weights = np.random.randn(output_dim, input_dim, n_experts)
    train_labels = np.swapaxes(np.matmul(dataset.train_data, weights), 0, 1) + biases
    train_labels = train_labels[range(len(dataset.train_data)), :, dataset.train_classes]
    test_labels = np.swapaxes(np.matmul(dataset.test_data, weights), 0, 1) + biases
    test_labels = test_labels[range(len(dataset.test_data)), :, dataset.test_classes]
    return train_labels, test_labels


def generate_classification_variables(dataset):
    """includes/utils.py:63-74."""
    n_classes = dataset.n_classes
    test_labels = np.zeros((len(dataset.test_classes), n_classes))
    test_labels[np.arange(0, len(dataset.test_classes)), dataset.test_classes] = 1
    train_labels = np.zeros((len(dataset.train_classes), n_classes))
    train_labels[np.arange(0, len(dataset.train_classes)), dataset.train_classes] = 1
    return train_labels, test_labels


class _Bag:
    pass


def _synthetic(name, n_train, n_test, dim, binarised, n_classes=10, seed=1):
    """X[b,d] = 1{u < 0.1307} (MNIST-shaped) or randint(0,256)/255 (CIFAR-shaped); labels b mod 10."""
    rng = np.random.RandomState(seed)
    ds = _Bag()

    def make(n):
        if binarised:
            return (rng.uniform(size=(n, dim)) < 0.1307).astype(np.float32)
        return (rng.randint(0, 256, size=(n, dim)) / 255.0).astype(np.float32)

    ds.datagroup = name
    ds.train_data, ds.test_data = make(n_train), make(n_test)
    ds.train_classes = (np.arange(n_train) % n_classes).astype(np.int64)
    ds.test_classes = (np.arange(n_test) % n_classes).astype(np.int64)
    ds.n_classes, ds.input_dim, ds.input_type = n_classes, dim, "binary"
    ds.sample_plot = ds.regeneration_plot = None
    return ds


def load_data(datagroup, output_dim=1, classification=True, **args):
    """includes/utils.py:77-375.  The real corpora need downloads (no network here): ``spiral`` is generated
    exactly as the reference does; ``synthetic_mnist`` / ``synthetic_cifar`` are the benchmark shapes;
    the download-backed names raise NotImplementedError with a pointer to the synthetic ones."""
    if datagroup == "spiral":
        N_tr, N_ts, D, K = args.get("N_tr", 5000), args.get("N_ts", 1000), 2, args.get("K", 5)
        ds = _Bag()
        train_data, test_data = np.zeros((N_tr * K, D)), np.zeros((N_ts * K, D))
        for data, N in ((train_data, N_tr), (test_data, N_ts)):
            for j in range(K):
                ix = range(N * j, N * (j + 1))
                r = np.linspace(2.5, 10.0, N)
                t = np.linspace(j * 1.25, (j + 1) * 1.25, N) + np.random.randn(N) * 0.05
                data[ix] = np.c_[r * np.sin(t), r * np.cos(t)]
        ds.datagroup = "spiral"
        ds.test_data, ds.test_classes = test_data, np.arange(K).repeat(N_ts)
        ds.train_data, ds.train_classes = train_data, np.arange(K).repeat(N_tr)
        ds.n_classes, ds.input_dim, ds.input_type = 5, 2, "real"
        ds.sample_plot = ds.regeneration_plot = None
    elif datagroup in ("synthetic_mnist", "synthetic"):
        ds = _synthetic("synthetic_mnist", args.get("n_train", 55000), args.get("n_test", 10000), 784, True)
    elif datagroup == "synthetic_cifar":
        ds = _synthetic("synthetic_cifar", args.get("n_train", 50000), args.get("n_test", 10000), 3072, False)
    elif datagroup in ("mnist", "cifar10"):
        # the reference downloads these (utils.py:122-148, :204-210); offline, DMVAE_DATA_DIR/<name>.npz (arrays
        # train_data, train_classes, test_data, test_classes) is used when present, else the synthetic stand-in of the
        # same shape - so that the reference's default command line (`--dataset mnist`) runs
        import os
        path = os.path.join(os.environ.get("DMVAE_DATA_DIR", ""), datagroup + ".npz")
        if os.environ.get("DMVAE_DATA_DIR") and os.path.exists(path):
            z = np.load(path)
            ds = _Bag()
            ds.datagroup = datagroup
            ds.train_data, ds.test_data = z["train_data"].astype(np.float32), z["test_data"].astype(np.float32)
            ds.train_classes, ds.test_classes = z["train_classes"].astype(np.int64), z["test_classes"].astype(np.int64)
            ds.n_classes, ds.input_dim, ds.input_type = 10, ds.train_data.shape[1], "binary"
            ds.sample_plot = ds.regeneration_plot = None
        else:
            print("note: dataset %r is not on disk (set DMVAE_DATA_DIR); using the synthetic stand-in of the same shape" % datagroup)
            if datagroup == "mnist":
                ds = _synthetic("mnist", args.get("n_train", 55000), args.get("n_test", 10000), 784, True)
            else:
                ds = _synthetic("cifar10", args.get("n_train", 50000), args.get("n_test", 10000), 3072, False)
    elif datagroup in ("hhar", "reuters", "reuters10k"):
        raise NotImplementedError("dataset %r needs a download (no network in this build); use 'synthetic_mnist', "
                                  "'synthetic_cifar' or 'spiral', or pass your own arrays to Dataset" % datagroup)
    else:
        raise NotImplementedError
    if classification:
        ds.train_labels, ds.test_labels = generate_classification_variables(ds)
    else:
        ds.train_labels, ds.test_labels = generate_regression_variable(ds, output_dim)
    return ds


def _pinned(a: np.ndarray, dtype) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=dtype))
    if torch.cuda.is_available():
        try:
            return t.pin_memory()
        except RuntimeError:
            pass
    return t


def _storage_format(data: np.ndarray):
    """(dtype, scale): uint8 with scale 1 when the data is exactly {0,1}-valued, uint8 with scale 1/255 when it is
    8-bit intensities k/255 (the CIFAR pixels of includes/utils.py:204-210) - both lossless up to float32 rounding and
    4x less host->device traffic and device re-reads - else float32 with scale 1."""
    if data.size and np.all((data == 0) | (data == 1)):
        return np.uint8, 1.0
    if data.size and data.min() >= 0 and data.max() <= 1:
        k = np.rint(data * 255.0)
        if np.all(np.abs(data * 255.0 - k) < 1e-3):
            return np.uint8, 1.0 / 255.0
    return np.float32, 1.0


def _pack_bits(data: np.ndarray):
    """{0,1}-valued rows at one bit per element (numpy little bit order), each row padded with zero bytes to a multiple of
    16 bytes: what dmvae_gather_rows_bits expands on the device.  None when the data is not binary or D % 16 != 0."""
    if os.environ.get("DMVAE_PACK_BITS", "1") == "0":          # A/B switch: move uint8 rows instead
        return None
    if not data.size or data.shape[1] % 16 or not np.all((data == 0) | (data == 1)):
        return None
    bits = np.packbits(data.astype(bool), axis=1, bitorder="little")
    pitch = (bits.shape[1] + 15) // 16 * 16
    out = np.zeros((data.shape[0], pitch), np.uint8)
    out[:, :bits.shape[1]] = bits
    return out


def _storage_dtype(data: np.ndarray):
    return _storage_format(data)[0]


def _to_storage(data: np.ndarray, dtype, scale) -> np.ndarray:
    if dtype == np.uint8 and scale != 1.0:
        return np.rint(data * 255.0).astype(np.uint8)
    return data


class Dataset:
    """includes/utils.py:428-466.  ``data`` is (data, classes).

    The reference re-orders ``self.data`` in place at every epoch and builds each batch with a per-row Python append.
    Here the rows stay where they are: an epoch is a fresh PERMUTATION (same ``np.random.permutation`` draw as the
    reference, utils.py:450-454), ``get_batches`` yields ``data[perm[i:i+B]]`` for reference-shaped loops, and the fast
    training path hands the permutation to the device, which gathers each batch straight out of ONE pinned host copy
    (made once, in storage dtype) - no per-epoch host gather, no re-pinning."""

    def __init__(self, data, batch_size=100, shuffle=True, compact=True):
        data, classes = data
        self.data = np.copy(data)
        self.classes = np.copy(classes)
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.data_dim = self.data.shape[1]
        self.epoch_len = int(math.ceil(len(self.data) / batch_size))
        self.len = len(self.data)
        self._host: Optional[torch.Tensor] = None
        self._compact = compact
        self.perm: Optional[np.ndarray] = None
        if shuffle:
            self._permute()

    def _permute(self):
        nxt = getattr(self, "_next_perm", None)
        self._next_perm = None
        self.perm = nxt if nxt is not None else np.random.permutation(len(self.data))

    def prefetch_epoch(self):
        """Draw the NEXT epoch's permutation now (same generator, same order of draws): the fast training path calls
        this while the device is still working through the current epoch, so the draw costs no device time."""
        if self.shuffle and getattr(self, "_next_perm", None) is None:
            self._next_perm = np.random.permutation(len(self.data))

    def host_tensor(self) -> torch.Tensor:
        """Pinned host copy in storage dtype (uint8 for binarised data), rows in their ORIGINAL order; built once."""
        if self._host is None:
            dt, self.host_scale = _storage_format(self.data) if self._compact else (np.float32, 1.0)
            self._host = _pinned(_to_storage(self.data, dt, self.host_scale), dt)
        return self._host

    host_scale = 1.0        # value of one unit of host_tensor() (1/255 for 8-bit intensities stored as uint8)

    def host_bits(self):
        """(pinned [N, pitch] uint8 holding the rows at ONE BIT per element, D) for binarised data, else None; built once.
        The training path moves these 8x fewer bytes per batch and expands them on the device."""
        if not hasattr(self, "_bits"):
            packed = _pack_bits(self.data) if self._compact else None
            self._bits = None if packed is None else (_pinned(packed, np.uint8), self.data.shape[1])
        return self._bits

    def begin_epoch(self):
        """New epoch order, drawn like get_batches does at the start of every epoch (utils.py:450-454)."""
        if self.shuffle:
            self._permute()

    def epoch_classes(self) -> np.ndarray:
        """Classes in the order of the current epoch's batches."""
        return self.classes if self.perm is None else self.classes[self.perm]

    def get_batches(self):
        self.begin_epoch()
        for i in range(0, len(self.data), self.batch_size):
            if self.perm is None:
                yield self.data[i: i + self.batch_size]
            else:
                yield self.data[self.perm[i: i + self.batch_size]]

    def __len__(self):
        return self.epoch_len


class MEDataset:
    """includes/utils.py:378-425.  ``data`` is (data, classes, labels)."""

    def __init__(self, data, batch_size=100, shuffle=True):
        self.data, self.classes, self.labels = [np.asarray(a) for a in data]
        self.shuffle = shuffle
        self.len = len(self.data)
        assert len(self.labels) == self.len and len(self.classes) == self.len
        self.batch_size = batch_size
        self.epoch_len = int(math.ceil(len(self.data) / batch_size))
        self.data_dim = self.data.shape[1]
        self.perm: Optional[np.ndarray] = None
        self._host = None

    def host_tensors(self):
        """Pinned host copies (data in storage dtype, labels float32), rows in their original order; built once."""
        if self._host is None:
            lab = self.labels.reshape(self.len, -1)
            dt, self.host_scale = _storage_format(self.data)
            self._host = (_pinned(_to_storage(self.data, dt, self.host_scale), dt), _pinned(lab, np.float32))
        return self._host

    host_scale = 1.0

    def host_bits(self):
        """Like Dataset.host_bits, for the expert models' inputs."""
        if not hasattr(self, "_bits"):
            packed = _pack_bits(self.data)
            self._bits = None if packed is None else (_pinned(packed, np.uint8), self.data.shape[1])
        return self._bits

    def begin_epoch(self):
        if self.shuffle:
            nxt = getattr(self, "_next_perm", None)
            self._next_perm = None
            self.perm = nxt if nxt is not None else np.random.permutation(len(self.data))

    def prefetch_epoch(self):
        if self.shuffle and getattr(self, "_next_perm", None) is None:
            self._next_perm = np.random.permutation(len(self.data))

    def get_batches(self):
        self.begin_epoch()
        for i in range(0, len(self.data), self.batch_size):
            s = slice(i, i + self.batch_size)
            if self.perm is None:
                yield self.data[s], self.labels[s], self.classes[s]
            else:
                ix = self.perm[s]
                yield self.data[ix], self.labels[ix], self.classes[ix]

    def __len__(self):
        return self.epoch_len

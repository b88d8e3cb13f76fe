"""v1's small-dataset helpers: MNIST / CIFAR loaders (from local files -- nothing is downloaded), one-hot conversion and the image
augmentations its CNN examples use.  (ref: hetu/v1/python/hetu/data.py)"""
from __future__ import annotations

import gzip
import os
import pickle

import numpy as np


def convert_to_one_hot(vals, max_val: int = 0):
    vals = np.asarray(vals).astype(np.int64).reshape(-1)
    n = max(int(max_val), int(vals.max()) + 1 if vals.size else 0)
    out = np.zeros((vals.size, n), np.float32)
    out[np.arange(vals.size), vals] = 1.0
    return out


def mnist(dataset="mnist.pkl.gz", onehot=True):
    """[(train_x, train_y), (valid_x, valid_y), (test_x, test_y)] from the classic pickled archive at `dataset`"""
    if not os.path.exists(dataset):
        raise FileNotFoundError(f"{dataset}: place the MNIST archive there (no download is attempted)")
    with gzip.open(dataset, "rb") as f:
        splits = pickle.load(f, encoding="latin1")
    return [(np.asarray(x, np.float32), convert_to_one_hot(y, 10) if onehot else np.asarray(y)) for x, y in splits]


def _cifar_batches(directory, names, label_key):
    xs, ys = [], []
    for n in names:
        with open(os.path.join(directory, n), "rb") as f:
            d = pickle.load(f, encoding="latin1")
        xs.append(np.asarray(d["data"], np.float32).reshape(-1, 3, 32, 32))
        ys.append(np.asarray(d[label_key]))
    return np.concatenate(xs), np.concatenate(ys)


def cifar10(directory="CIFAR_10", onehot=True):
    tx, ty = _cifar_batches(directory, [f"data_batch_{i}" for i in range(1, 6)], "labels")
    vx, vy = _cifar_batches(directory, ["test_batch"], "labels")
    return (tx, convert_to_one_hot(ty, 10) if onehot else ty, vx, convert_to_one_hot(vy, 10) if onehot else vy)


def cifar100(directory="CIFAR_100", onehot=True):
    tx, ty = _cifar_batches(directory, ["train"], "fine_labels")
    vx, vy = _cifar_batches(directory, ["test"], "fine_labels")
    return (tx, convert_to_one_hot(ty, 100) if onehot else ty, vx, convert_to_one_hot(vy, 100) if onehot else vy)


def normalize_cifar(num_class=10, onehot=True, directory=None):
    """per-channel standardised CIFAR (statistics of the training split)"""
    tx, ty, vx, vy = (cifar10 if num_class == 10 else cifar100)(directory or f"CIFAR_{num_class}", onehot)
    mean, std = tx.mean((0, 2, 3), keepdims=True), tx.std((0, 2, 3), keepdims=True) + 1e-7
    return (tx - mean) / std, ty, (vx - mean) / std, vy


def tf_normalize_cifar(num_class=10, onehot=True, directory=None):
    """per-image standardisation (tf.image.per_image_standardization)"""
    tx, ty, vx, vy = (cifar10 if num_class == 10 else cifar100)(directory or f"CIFAR_{num_class}", onehot)
    return _image_whitening(tx), ty, _image_whitening(vx), vy


def _image_crop(images, shape, rng=None):
    """random crops of `shape` (h, w) after 4-pixel zero padding"""
    rng = rng or np.random
    n, c, h, w = images.shape
    pad = np.pad(images, [(0, 0), (0, 0), (4, 4), (4, 4)])
    out = np.empty((n, c, shape[0], shape[1]), images.dtype)
    for i in range(n):
        y, x = rng.randint(0, h + 8 - shape[0] + 1), rng.randint(0, w + 8 - shape[1] + 1)
        out[i] = pad[i, :, y:y + shape[0], x:x + shape[1]]
    return out


def _image_crop_test(images, shape):
    n, c, h, w = images.shape
    y, x = (h - shape[0]) // 2, (w - shape[1]) // 2
    return images[:, :, y:y + shape[0], x:x + shape[1]]


def _image_flip(images, rng=None):
    rng = rng or np.random
    flip = rng.rand(images.shape[0]) < 0.5
    out = images.copy()
    out[flip] = out[flip][..., ::-1]
    return out


def _image_whitening(images):
    flat = images.reshape(images.shape[0], -1)
    mean = flat.mean(1, keepdims=True)
    std = np.maximum(flat.std(1, keepdims=True), 1.0 / np.sqrt(flat.shape[1]))
    return ((flat - mean) / std).reshape(images.shape).astype(np.float32)


def _image_noise(images, mean=0, std=0.01, rng=None):
    rng = rng or np.random
    return images + rng.normal(mean, std, images.shape).astype(images.dtype)


def data_augmentation(images, mode="train", flip=False, crop=False, crop_shape=(24, 24), whiten=False, noise=False, noise_mean=0,
                      noise_std=0.01, rng=None):
    if crop:
        images = _image_crop(images, crop_shape, rng) if mode == "train" else _image_crop_test(images, crop_shape)
    if flip and mode == "train":
        images = _image_flip(images, rng)
    if whiten:
        images = _image_whitening(images)
    if noise and mode == "train":
        images = _image_noise(images, noise_mean, noise_std, rng)
    return images

"""Image classifiers of the model zoo on a synthetic CIFAR-shaped task (class = which quadrant carries a bright patch).

    python examples/cnn/train_cnn.py --model resnet18 --steps 40
    models: logreg mlp lenet cnn alexnet vgg16 vgg19 resnet18 resnet34 rnn lstm (the last two read the image row by row)

(ref: hetu/v1/examples/cnn/main.py + models/)"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import hetu_b200 as ht
from hetu_b200 import models as M

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="cnn")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--lr", type=float, default=2e-3)
ap.add_argument("--width-div", type=int, default=4, help="shrink VGG / ResNet channel counts (1 = paper sizes)")
a = ap.parse_args()
B, C, HW, K = a.batch, 3, 32, 4
zoo = {"logreg": lambda: M.LogReg(C * HW * HW, K), "mlp": lambda: M.MLP(C * HW * HW, (256, 128), K), "lenet": lambda: M.LeNet(C, K, HW),
       "cnn": lambda: M.CNN3(C, K, HW), "alexnet": lambda: M.AlexNet(C, K, HW), "vgg16": lambda: M.VGG(16, C, K, HW, width_div=a.width_div),
       "vgg19": lambda: M.VGG(19, C, K, HW, width_div=a.width_div), "resnet18": lambda: M.ResNet(18, C, K, width=64 // a.width_div),
       "resnet34": lambda: M.ResNet(34, C, K, width=64 // a.width_div), "rnn": lambda: M.RNN(C * HW, 128, K), "lstm": lambda: M.LSTM(C * HW, 128, K)}
sequence = a.model in ("rnn", "lstm")


def batch(rng):
    x = rng.randn(B, C, HW, HW).astype(np.float32) * 0.3
    y = rng.randint(0, K, B)
    for i, k in enumerate(y):
        r, c = (k // 2) * 16, (k % 2) * 16
        x[i, :, r + 4:r + 12, c + 4:c + 12] += 1.5
    if sequence:
        x = x.transpose(0, 2, 1, 3).reshape(B, HW, C * HW)      # rows as time steps
    return x, y


with ht.graph("define_and_run", create_new=True) as g:
    model = zoo[a.model]()
    X = ht.placeholder("float32", [B, HW, C * HW] if sequence else [B, C, HW, HW], name="x")
    Y = ht.placeholder("int64", [B], name="y")
    loss, logits = model(X, Y)
    train = ht.AdamOptimizer(lr=a.lr).minimize(loss)
rng = np.random.RandomState(0)
for step in range(a.steps):
    x, y = batch(rng)
    out = g.run(loss, [loss, logits, train], {X: torch.as_tensor(x), Y: torch.as_tensor(y)})
    if step % 10 == 0 or step == a.steps - 1:
        print(f"{a.model} step {step} loss {float(out[0]):.4f} acc {float((out[1].float().argmax(1).cpu().numpy() == y).mean()):.2f}", flush=True)

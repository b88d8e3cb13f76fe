"""Image classifiers of the reference's v1 model zoo: logistic regression, MLP, LeNet, a 3-layer CNN, AlexNet, VGG-16/19 and
ResNet-18/34 (CIFAR-style stems), all [N, C, H, W] -> logits.  `model(x, labels)` returns (loss, logits).
(ref: hetu/v1/examples/cnn/models/{LogReg,MLP,LeNet,CNN,AlexNet,VGG,ResNet}.py)"""
from __future__ import annotations

from typing import List, Sequence

from .. import ops
from ..nn import BatchNorm, Conv2d, Dropout, Linear, MaxPool2d, Module, ModuleList


class _Classifier(Module):
    def loss(self, logits, labels):
        return ops.softmax_cross_entropy_sparse(logits, labels, reduction="mean")

    def _out(self, logits, labels):
        return logits if labels is None else (self.loss(logits, labels), logits)


def _flat(x):
    n = 1
    for s in x.shape[1:]:
        n *= s
    return ops.reshape(x, [x.shape[0], n])


class LogReg(_Classifier):
    def __init__(self, in_features=784, num_classes=10):
        super().__init__()
        self.fc = Linear(in_features, num_classes, name="logreg_fc")

    def forward(self, x, labels=None):
        return self._out(self.fc(_flat(x)), labels)


class MLP(_Classifier):
    def __init__(self, in_features=3072, hidden: Sequence[int] = (256, 256), num_classes=10):
        super().__init__()
        dims = [in_features] + list(hidden)
        self.layers = ModuleList([Linear(a, b, name=f"mlp_fc{i}") for i, (a, b) in enumerate(zip(dims[:-1], dims[1:]))])
        self.out = Linear(dims[-1], num_classes, name="mlp_out")

    def forward(self, x, labels=None):
        h = _flat(x)
        for l in self.layers:
            h = l(h, act="relu")
        return self._out(self.out(h), labels)


class LeNet(_Classifier):
    """conv5-pool-conv5-pool-fc120-fc84-fc (28 x 28 or 32 x 32 inputs)"""

    def __init__(self, in_channels=1, num_classes=10, image_size=28):
        super().__init__()
        self.c1 = Conv2d(in_channels, 6, 5, padding=2, name="lenet_c1")
        self.c2 = Conv2d(6, 16, 5, name="lenet_c2")
        self.pool = MaxPool2d(2, 2)
        side = (image_size // 2 - 4) // 2
        self.f1 = Linear(16 * side * side, 120, name="lenet_f1")
        self.f2 = Linear(120, 84, name="lenet_f2")
        self.f3 = Linear(84, num_classes, name="lenet_f3")

    def forward(self, x, labels=None):
        h = self.pool(ops.relu(self.c1(x)))
        h = self.pool(ops.relu(self.c2(h)))
        return self._out(self.f3(self.f2(self.f1(_flat(h), act="relu"), act="relu")), labels)


class CNN3(_Classifier):
    """three conv-bn-relu-pool stages and a linear head (the reference's `CNN` for CIFAR-10)"""

    def __init__(self, in_channels=3, num_classes=10, image_size=32, width=32):
        super().__init__()
        chans = [in_channels, width, 2 * width, 4 * width]
        self.convs = ModuleList([Conv2d(a, b, 3, padding=1, name=f"cnn3_c{i}") for i, (a, b) in enumerate(zip(chans[:-1], chans[1:]))])
        self.bns = ModuleList([BatchNorm(b, name=f"cnn3_bn{i}") for i, b in enumerate(chans[1:])])
        self.pool = MaxPool2d(2, 2)
        self.fc = Linear(chans[-1] * (image_size // 8) ** 2, num_classes, name="cnn3_fc")

    def forward(self, x, labels=None):
        h = x
        for c, b in zip(self.convs, self.bns):
            h = self.pool(ops.relu(b(c(h))))
        return self._out(self.fc(_flat(h)), labels)


class AlexNet(_Classifier):
    """AlexNet with the CIFAR-sized stem (3 x 3 convolutions, three pools)"""

    def __init__(self, in_channels=3, num_classes=10, image_size=32, dropout=0.5):
        super().__init__()
        cfg = [(in_channels, 64), (64, 192), (192, 384), (384, 256), (256, 256)]
        self.convs = ModuleList([Conv2d(a, b, 3, padding=1, name=f"alexnet_c{i}") for i, (a, b) in enumerate(cfg)])
        self.pool_after = {0, 1, 4}
        self.pool = MaxPool2d(2, 2)
        side = image_size // 8
        self.drop = Dropout(dropout)
        self.f1 = Linear(256 * side * side, 1024, name="alexnet_f1")
        self.f2 = Linear(1024, 1024, name="alexnet_f2")
        self.f3 = Linear(1024, num_classes, name="alexnet_f3")

    def forward(self, x, labels=None):
        h = x
        for i, c in enumerate(self.convs):
            h = ops.relu(c(h))
            if i in self.pool_after:
                h = self.pool(h)
        h = self.f1(self.drop(_flat(h)), act="relu")
        h = self.f2(self.drop(h), act="relu")
        return self._out(self.f3(h), labels)


_VGG = {16: [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"],
        19: [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]}


class VGG(_Classifier):
    def __init__(self, depth=16, in_channels=3, num_classes=10, image_size=32, batch_norm=True, width_div=1):
        super().__init__()
        self.plan: List = []
        convs, bns, c = [], [], in_channels
        for v in _VGG[depth]:
            if v == "M":
                self.plan.append("M")
                continue
            o = max(v // width_div, 4)
            convs.append(Conv2d(c, o, 3, padding=1, name=f"vgg{depth}_c{len(convs)}"))
            bns.append(BatchNorm(o, name=f"vgg{depth}_bn{len(bns)}") if batch_norm else None)
            self.plan.append(len(convs) - 1)
            c = o
        self.convs = ModuleList(convs)
        self.bns = ModuleList([b for b in bns if b is not None]) if batch_norm else None
        self.batch_norm = batch_norm
        self.pool = MaxPool2d(2, 2)
        side = max(image_size // 32, 1)
        self.fc = Linear(c * side * side, num_classes, name=f"vgg{depth}_fc")

    def forward(self, x, labels=None):
        h = x
        for step in self.plan:
            if step == "M":
                h = self.pool(h)
            else:
                h = self.convs[step](h)
                if self.batch_norm:
                    h = self.bns[step](h)
                h = ops.relu(h)
        return self._out(self.fc(_flat(h)), labels)


class _BasicBlock(Module):
    def __init__(self, cin, cout, stride, name):
        super().__init__()
        self.c1 = Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False, name=f"{name}_c1")
        self.b1 = BatchNorm(cout, name=f"{name}_b1")
        self.c2 = Conv2d(cout, cout, 3, padding=1, bias=False, name=f"{name}_c2")
        self.b2 = BatchNorm(cout, name=f"{name}_b2")
        self.down = None
        if stride != 1 or cin != cout:
            self.down = Conv2d(cin, cout, 1, stride=stride, bias=False, name=f"{name}_down")
            self.down_bn = BatchNorm(cout, name=f"{name}_down_bn")

    def forward(self, x):
        h = ops.relu(self.b1(self.c1(x)))
        h = self.b2(self.c2(h))
        s = x if self.down is None else self.down_bn(self.down(x))
        return ops.relu(h + s)


class ResNet(_Classifier):
    """ResNet-18 / 34 (basic blocks), CIFAR stem (3 x 3 convolution, no max-pool)"""

    def __init__(self, depth=18, in_channels=3, num_classes=10, width=64):
        super().__init__()
        blocks = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3]}[depth]
        self.stem = Conv2d(in_channels, width, 3, padding=1, bias=False, name=f"resnet{depth}_stem")
        self.stem_bn = BatchNorm(width, name=f"resnet{depth}_stem_bn")
        layers, c = [], width
        for stage, n in enumerate(blocks):
            o = width * 2 ** stage
            for j in range(n):
                layers.append(_BasicBlock(c, o, 2 if (j == 0 and stage > 0) else 1, f"resnet{depth}_s{stage}b{j}"))
                c = o
        self.blocks = ModuleList(layers)
        self.fc = Linear(c, num_classes, name=f"resnet{depth}_fc")

    def forward(self, x, labels=None):
        h = ops.relu(self.stem_bn(self.stem(x)))
        for b in self.blocks:
            h = b(h)
        h = ops.mean(ops.reshape(h, [h.shape[0], h.shape[1], h.shape[2] * h.shape[3]]), [2])          # global average pool
        return self._out(self.fc(h), labels)

"""v1 optimizers (ref: hetu/v1/python/hetu/optimizer.py): `.minimize(loss)` returns the training node."""
from .. import optim


_LIVE = []      # optimizers with update ops in a graph (the Executor applies their lr schedulers)


class _Base:
    """`learning_rate` is a float or an `lr_scheduler` object (its value is applied before, and stepped after, every
    training run of the Executor)"""

    def __init__(self, learning_rate=0.01, l2reg=0.0):
        self.scheduler = learning_rate if hasattr(learning_rate, "get") else None
        self.learning_rate = float(learning_rate.get()) if self.scheduler is not None else learning_rate
        self.l2reg = l2reg

    def minimize(self, loss, var_list=None):
        opt = self._make()
        opt.v1_scheduler = self.scheduler
        self.backend = opt
        node = opt.minimize(loss, var_list) if var_list is not None else opt.minimize(loss)
        # what the parameter-server modes of the Executor need: which loss / variables this training node stands for and how the
        # server should apply a gradient (the local update ops stay in the graph but are not fetched in PS mode)
        opt.v1_loss, opt.v1_var_list, opt.v1_train_node = loss, var_list, node
        opt.v1_server_opt = self._server_opt()
        _LIVE.append(opt)
        return node

    def _server_opt(self):
        return ("sgd", float(self.learning_rate))


class SGDOptimizer(_Base):
    def _make(self):
        return optim.SGDOptimizer(lr=self.learning_rate, weight_decay=self.l2reg)


class MomentumOptimizer(_Base):
    def __init__(self, learning_rate=0.01, momentum=0.9, nesterov=False, l2reg=0.0):
        super().__init__(learning_rate, l2reg)
        self.momentum, self.nesterov = momentum, nesterov

    def _make(self):
        return optim.SGDOptimizer(lr=self.learning_rate, momentum=self.momentum, nesterov=self.nesterov, weight_decay=self.l2reg)

    def _server_opt(self):
        return ("momentum", float(self.learning_rate))


class AdaGradOptimizer(_Base):
    def __init__(self, learning_rate=0.01, initial_accumulator_value=0.0, eps=1e-7, l2reg=0.0):
        super().__init__(learning_rate, l2reg)
        self.eps, self.initial_accumulator_value = eps, initial_accumulator_value

    def _server_opt(self):
        return ("adagrad", float(self.learning_rate))

    def _make(self):
        return optim.AdaGradOptimizer(lr=self.learning_rate, initial_accumulator_value=self.initial_accumulator_value, eps=self.eps,
                                      l2reg=self.l2reg)


class AdamOptimizer(_Base):
    def __init__(self, learning_rate=0.01, beta1=0.9, beta2=0.999, epsilon=1e-7, l2reg=0.0, amsgrad=False):
        super().__init__(learning_rate, l2reg)
        self.beta1, self.beta2, self.epsilon, self.amsgrad = beta1, beta2, epsilon, amsgrad

    def _server_opt(self):
        return ("adam", float(self.learning_rate))

    def _make(self):
        if self.amsgrad:
            return optim.AMSGradOptimizer(lr=self.learning_rate, beta1=self.beta1, beta2=self.beta2, eps=self.epsilon, l2reg=self.l2reg)
        return optim.AdamOptimizer(lr=self.learning_rate, beta1=self.beta1, beta2=self.beta2, eps=self.epsilon, weight_decay=self.l2reg)


class AMSGradOptimizer(AdamOptimizer):
    def __init__(self, learning_rate=0.01, beta1=0.9, beta2=0.999, epsilon=1e-7, l2reg=0.0):
        super().__init__(learning_rate, beta1, beta2, epsilon, l2reg, amsgrad=True)


class AdamWOptimizer(_Base):
    def __init__(self, learning_rate=0.01, beta1=0.9, beta2=0.999, epsilon=1e-7, weight_decay=0.0):
        super().__init__(learning_rate, 0.0)
        self.beta1, self.beta2, self.epsilon, self.weight_decay = beta1, beta2, epsilon, weight_decay

    def _server_opt(self):
        return ("adam", float(self.learning_rate))

    def _make(self):
        return optim.AdamWOptimizer(lr=self.learning_rate, beta1=self.beta1, beta2=self.beta2, eps=self.epsilon, weight_decay=self.weight_decay)


class LambOptimizer(AdamWOptimizer):
    def _make(self):
        return optim.LambOptimizer(lr=self.learning_rate, beta1=self.beta1, beta2=self.beta2, eps=self.epsilon, weight_decay=self.weight_decay)

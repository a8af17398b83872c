"""Layer descriptors mirroring code/includes/layers.py.  A layer here only records its shape; the arithmetic
(x @ W + b, activation) is one fused tcgen05 / SIMT GEMM of the engine."""
from .. import nn


class Layer:
    def __init__(self, name, activation=nn.relu, initializer=nn.xavier_initializer):
        self.name = name
        self.activation = activation
        self.initializer = initializer


class FullyConnected(Layer):
    """includes/layers.py:19-36: W (input_dim, output_dim) and bias (1, output_dim), BOTH xavier-initialised."""

    def __init__(self, name, input_dim, output_dim, activation=nn.relu, initializer=nn.xavier_initializer):
        Layer.__init__(self, name, activation=activation, initializer=initializer)
        nn.check_activation(activation)
        nn.check_initializer(initializer)
        self.input_dim, self.output_dim = input_dim, output_dim


class Convolution(Layer):
    def __init__(self, *a, **k):
        raise NotImplementedError("the CNN encoder (base_models.py:178-216) is outside the accelerated path")


class MaxPooling(Layer):
    def __init__(self, *a, **k):
        raise NotImplementedError("the CNN encoder (base_models.py:178-216) is outside the accelerated path")


class BatchNormalization(Layer):
    def __init__(self, *a, **k):
        raise NotImplementedError("batch normalisation is not on the accelerated path")

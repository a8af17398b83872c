"""code/includes/network.py: sequential container of layer descriptors."""
from .. import nn
from .layers import FullyConnected, Convolution, MaxPooling, BatchNormalization

_layers_id_mapping = {"fc": FullyConnected, "cn": Convolution, "mp": MaxPooling, "bn": BatchNormalization}


class DeepNetwork:
    """includes/network.py:57-82.  ``layers`` is a list of (id, args); unknown ids raise NotImplementedError."""

    def __init__(self, name, layers, activation=nn.relu, initializer=nn.xavier_initializer):
        self.name = name
        self.layers = []
        for index, (layer_id, args) in enumerate(layers):
            lname = "layer_%d" % (index + 1)
            if layer_id not in _layers_id_mapping:
                raise NotImplementedError
            self.layers.append(_layers_id_mapping[layer_id](lname, activation=activation, initializer=initializer, **args))

    def widths(self):
        return tuple(l.output_dim for l in self.layers)

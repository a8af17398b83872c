"""Stand-ins for the TensorFlow symbols the reference passes around (train.py:160-161):
``activation=tf.nn.relu`` and ``initializer=tf.contrib.layers.xavier_initializer``."""


def relu(x=None):
    """Marker for tf.nn.relu: the dense kernels fuse ReLU into the GEMM epilogue."""
    return "relu"


relu.kind = "relu"


class xavier_initializer:
    """Marker for tf.contrib.layers.xavier_initializer (uniform +-sqrt(6/(fan_in+fan_out)))."""
    kind = "xavier"

    def __call__(self):
        return self


def check_activation(activation):
    if activation is None or getattr(activation, "kind", None) == "relu" or activation == "relu":
        return "relu"
    raise NotImplementedError("only ReLU hidden activations are implemented (the reference passes tf.nn.relu)")


def check_initializer(initializer):
    if initializer is None or getattr(initializer, "kind", None) == "xavier" or initializer == "xavier":
        return "xavier"
    raise NotImplementedError("only the xavier initializer is implemented (the reference passes xavier_initializer)")

"""B200-native (sm_100a) message-passing kernels behind the layer call surface of
kaddly/GraphNeuralNetwork's GCN, GAT, GraphSAGE(_Pytorch) and HAN.

    graphneuralnetwork_b200.layers      drop-in nn.Modules with the reference's names
    graphneuralnetwork_b200.functional  autograd functions over the C ABI
    graphneuralnetwork_b200.graph       CSR construction on the device
    graphneuralnetwork_b200.partition   1-D row partition + halo exchange (multi-GPU SpMM)
    include/gnn_b200.h                  the C ABI itself (libgnn_b200.so)

Importing the package does not need a GPU; calling any op does, and raises otherwise.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"

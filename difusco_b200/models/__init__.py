from .gnn_encoder import GNNEncoder  # noqa: F401

"""dualmessagepassing_b200 -- B200-native DMPNN dual message-passing convolution.

Drop-in for the hot path of HKUST-KnowComp/DualMessagePassing:
  DMPLayer / DualGraphConv   same constructor + forward(graph, node_feat, edge_feat[, edge_norm])
  DMPGraph, batch, add_reversed_edges, build_graph_from_triplets   DGL-compatible graph container
The sparse core runs in hand-written sm_100a CUDA behind a C ABI (include/dmp_b200.h,
libdmp_b200.so); there is no CPU fallback.
"""
from . import constants
from .act import map_activation_str_to_layer, supported_act_funcs
from .graph import DMPGraph, add_reversed_edges, batch, build_graph_from_triplets, compute_edgenorm
from .layers import DMPLayer, DMPLRPPoolLayer, DualGraphConv, dual_message_passing
from .models import DMPNNRepNet, relation_mean_pool
from .plan import DMPPlan, get_plan

__all__ = ["DMPLayer", "DMPLRPPoolLayer", "DualGraphConv", "DMPNNRepNet", "DMPGraph", "DMPPlan", "batch", "add_reversed_edges",
           "build_graph_from_triplets", "compute_edgenorm", "dual_message_passing", "get_plan",
           "relation_mean_pool", "map_activation_str_to_layer", "supported_act_funcs", "constants"]

"""Graph-frame key names and scalars that are part of the reference's data contract.

A drop-in layer has to read and write the same `ndata` / `edata` keys as the reference
(SubgraphCountingMatching/constants.py:10-34; the UNC twin hard-codes its own keys,
UnsupervisedNodeClassification/Model/DMPNN/src/model.py:206-238).
"""
LEAKY_RELU_A = 1 / 5.5

# SubgraphCountingMatching frames
REVFLAG = "is_reversed"
OUTDEGREE = "out_deg"
INDEGREE = "in_deg"
NORM = "norm"
NODEID = "id"
EDGEID = "id"
NODELABEL = "label"
EDGELABEL = "label"
NODEFEAT = "node_feat"
EDGEFEAT = "edge_feat"
NODEAGG = "node_agg"
EDGEAGG = "edge_agg"
NODEOUTPUT = "node_out"
EDGEOUTPUT = "edge_out"

# UnsupervisedNodeClassification frames
UNC_FEAT = "h"
UNC_REVFLAG = "is_rev"
UNC_NORM = "norm"
UNC_OUTDEGREE = "out_deg"
UNC_EDGEAGG = "agg"

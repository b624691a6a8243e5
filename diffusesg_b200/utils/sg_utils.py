"""Dataset -> channel-count table (the shape contract of the hot path).

Restates utils/sg_utils.py:348-430 of the reference for the encodings its README runs use ('bits' for node and
edge classes, 4 bbox coordinates appended to the node channels).
"""
from __future__ import annotations

import math


def get_node_adj_num_type(dataset_name: str, flag_sg: bool = True, flag_node_only: bool = False):
    """(num_node_type, num_adj_type, num_allowed_nodes) - utils/sg_utils.py:348-409."""
    if "visual_genome" in dataset_name:
        return 150, 51, 62
    if "coco_stuff" in dataset_name:
        return 171, 7, 33
    raise NotImplementedError(f"unknown scene-graph dataset {dataset_name!r}")


def get_node_adj_model_input_output_channels(config):
    """(in_chans, out_chans_adj, out_chans_node) for bits encodings + bbox (utils/sg_utils.py:412-430)."""
    node_enc, edge_enc = config.train.node_encoding, config.train.edge_encoding
    if node_enc != "bits" or edge_enc != "bits":
        raise NotImplementedError("only the 'bits' encodings of the README runs are built "
                                  f"(got node={node_enc!r}, edge={edge_enc!r})")
    nt, et, _ = get_node_adj_num_type(config.dataset.name)
    out_adj = math.ceil(math.log2(et))
    out_node = math.ceil(math.log2(nt)) + 4
    return out_adj + 2 * out_node, out_adj, out_node

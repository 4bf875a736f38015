"""gcanet_b200 -- B200-native kNN-graph + EdgeConv path of GCANet's DGCNN backbone.

Public surface = the reference's own call signatures (see functional.py / modules.py);
all compute is in libgcanet_b200.so behind the C-ABI of include/gcanet_b200.h.
"""
from .functional import (  # noqa: F401
    KNN,
    affinity_ball_query,
    ball_query,
    compute_batch_adjacency_matrix,
    GroupingOperation,
    edgeconv,
    get_graph_feature,
    get_graph_feature_with_normals,
    get_graph_feature_with_normals_g,
    global_feature,
    group_norm,
    group_norm_relu,
    group_points,
    grouping_operation,
    knn,
    knn_cuda,
    knn_cuda_pair,
    knn_graph,
    knn_point,
    knn_points_normals,
    normal_edgeconv,
    offset_pred,
    splinenet_get_graph_feature,
    splinenet_knn,
    to_channel_major,
    to_point_major,
)
from .modules import (KPAM, OFFSET_PRED_MODULE, DGCNNEncoderGn, NormalEdgeHead, SoftProjection,  # noqa: F401
                      SppnetDGCNNEncoderGn)

__version__ = "0.2.0"

from .image_to_graph_optimized import (  # noqa: F401
    create_grid_edges_optimized, get_cached_edge_index, image_to_graph_pixel_optimized)
from .image_to_graph_patch import image_to_graph_patch  # noqa: F401
from .image_to_graph_superpixel import image_to_graph_superpixel  # noqa: F401
from .batched import GraphBatch, build_pixel_graphs, build_patch_graphs, build_superpixel_graphs  # noqa: F401

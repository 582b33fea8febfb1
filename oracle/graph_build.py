"""Oracle restatement of the reference's image -> graph builders (numpy, CPU).

TEST INFRASTRUCTURE - see oracle/__init__.py.  Every function cites the reference
lines it restates; outputs are pinned bit-exactly against the reference's own
functions by oracle/make_golden.py (run in the build container, where
/root/reference is mounted) and by tests/test_oracle_golden.py on the committed
vectors.

The builders here start from an already-resized ``uint8 [H, W, 3]`` array: the PIL
decode + ``Image.resize`` step (image_to_graph_optimized.py:65-70) is outside the
hot path.
"""
from __future__ import annotations

import numpy as np


def grid_edge_counts(H: int, W: int, diagonals: bool = False):
    """(E_h, E_v, E_d) for an H x W grid (image_to_graph_optimized.py:22-33)."""
    e_h = H * (W - 1)
    e_v = (H - 1) * W
    e_d = (H - 1) * (W - 1) if diagonals else 0
    return e_h, e_v, e_d


def grid_edges(H: int, W: int, diagonals: bool = False) -> np.ndarray:
    """Directed grid edge list ``int64 [2, E]``.

    Restates create_grid_edges_optimized (image_to_graph_optimized.py:7-39) in
    closed form: edges are emitted family by family - horizontal, vertical, then
    (optionally) the two diagonal families - each family in row-major order of its
    source pixel.  Node id of pixel (i, j) is ``i * W + j``.
    """
    e_h, e_v, e_d = grid_edge_counts(H, W, diagonals)
    E = e_h + e_v + 2 * e_d
    out = np.empty((2, E), dtype=np.int64)

    # horizontal: (i, j) -> (i, j + 1), j < W - 1                      (:22)
    k = np.arange(e_h, dtype=np.int64)
    if W > 1:
        i, j = np.divmod(k, W - 1)
        out[0, :e_h] = i * W + j
        out[1, :e_h] = i * W + j + 1
    # vertical: (i, j) -> (i + 1, j), i < H - 1; source id == running index (:25)
    k = np.arange(e_v, dtype=np.int64)
    out[0, e_h:e_h + e_v] = k
    out[1, e_h:e_h + e_v] = k + W
    if diagonals and e_d:
        k = np.arange(e_d, dtype=np.int64)
        i, j = np.divmod(k, W - 1)
        o = e_h + e_v
        # d1: (i, j) -> (i + 1, j + 1)                                  (:31)
        out[0, o:o + e_d] = i * W + j
        out[1, o:o + e_d] = (i + 1) * W + j + 1
        # d2: (i, j + 1) -> (i + 1, j)                                  (:33)
        o += e_d
        out[0, o:o + e_d] = i * W + j + 1
        out[1, o:o + e_d] = (i + 1) * W + j
    return out


def pixel_graph(img_u8: np.ndarray, diagonals: bool = False):
    """Pixel graph of one resized image.

    Restates image_to_graph_pixel_optimized lines 71-87: ``x`` is the row-major
    pixel list (uint8, un-normalised), ``pos`` the integer (row, col) of each
    pixel, ``edge_index`` the grid edges.  Returns the numpy triple in the
    reference's dtypes (uint8, int64, int64).
    """
    H, W, C = img_u8.shape
    x = np.ascontiguousarray(img_u8).reshape(H * W, C)
    v = np.arange(H * W, dtype=np.int64)
    pos = np.stack([v // W, v % W], axis=1)
    return x, pos, grid_edges(H, W, diagonals)


def patch_graph(img_u8: np.ndarray, patch_size: int = 8):
    """Patch graph (image_to_graph_patch.py:25-54).

    One node per ``patch_size x patch_size`` tile (tiles that do not fit are
    dropped), feature = mean RGB over the tile computed the way ``np.mean`` does
    on a uint8 block (float64 accumulation, 0..255 range), position = tile centre
    ``(i*p + p//2, j*p + p//2)``, edges = non-diagonal grid over the tiles.
    Returns float64 ``x [n, 3]``, int64 ``pos [n, 2]``, int64 ``edge_index``.
    """
    H, W, C = img_u8.shape
    p = int(patch_size)
    nh, nw = H // p, W // p
    tiles = img_u8[: nh * p, : nw * p].reshape(nh, p, nw, p, C)
    # np.mean over a (p, p, C) uint8 block with axis=(0, 1): float64 sum / p*p.
    # The reference reduces each tile separately; summing integers <= 255 in
    # float64 is exact, so any summation order gives the same bits.
    x = tiles.astype(np.float64).sum(axis=(1, 3)) / float(p * p)
    x = x.reshape(nh * nw, C)
    ii, jj = np.divmod(np.arange(nh * nw, dtype=np.int64), max(nw, 1))
    pos = np.stack([ii * p + p // 2, jj * p + p // 2], axis=1)
    return x, pos, grid_edges(nh, nw, False)


def superpixel_graph_from_labels(img_u8: np.ndarray, segments: np.ndarray):
    """Label map -> superpixel graph (image_to_graph_superpixel.py:28, 34-71).

    ``segments`` is the integer label map SLIC would return (``[H, W]``).  Node id
    = rank of the label among the sorted unique labels (:34).  Features are the
    mean of ``img/255`` (float64, ``img_as_float`` of a uint8 image) over the
    segment, positions the centroid (mean row, mean col) (:41-49).  Two segments
    are adjacent when the default (4-connected, cross-shaped) binary dilation of
    one touches the other (:54-66), i.e. when some pixel of one has a pixel of
    the other directly left/right/above/below.  Edges are emitted for i < j in
    lexicographic order as (i, j) then (j, i).
    Returns float64 ``x [S, 3]``, float64 ``pos [S, 2]``, int64 ``edge_index``
    (``float64 [2, 0]`` when there are no edges, as the reference does, :70-71).
    """
    img = img_u8.astype(np.float64) / 255.0          # skimage img_as_float(uint8)
    labels, inv = np.unique(segments, return_inverse=True)
    inv = inv.reshape(segments.shape)
    S = len(labels)
    H, W = segments.shape
    x = np.empty((S, img.shape[2]), dtype=np.float64)
    pos = np.empty((S, 2), dtype=np.float64)
    rr, cc = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    for s in range(S):
        m = inv == s
        # np.mean(img[mask], axis=0): pairwise float64 sums over the masked
        # pixels in row-major order - reproduced by the same call.
        x[s] = np.mean(img[m], axis=0)
        pos[s, 0] = np.mean(rr[m])
        pos[s, 1] = np.mean(cc[m])
    adj = np.zeros((S, S), dtype=bool)
    a, b = inv[:, :-1], inv[:, 1:]
    adj[a, b] = True
    adj[b, a] = True
    a, b = inv[:-1, :], inv[1:, :]
    adj[a, b] = True
    adj[b, a] = True
    np.fill_diagonal(adj, False)
    iu, ju = np.nonzero(np.triu(adj, 1))              # lexicographic (i, j), i < j
    if len(iu) == 0:
        return x, pos, np.empty((2, 0))
    ei = np.empty((2, 2 * len(iu)), dtype=np.int64)
    ei[0, 0::2], ei[1, 0::2] = iu, ju
    ei[0, 1::2], ei[1, 1::2] = ju, iu
    return x, pos, ei


def to_model_inputs(x, pos, edge_index):
    """The cast the loader applies (utils/dataloader.py:49-51): f32, f32, int64."""
    import torch

    return (
        torch.tensor(np.asarray(x), dtype=torch.float32),
        torch.tensor(np.asarray(pos), dtype=torch.float32),
        torch.tensor(np.asarray(edge_index), dtype=torch.long),
    )


def batch_graphs(xs, poss, edge_indexes):
    """Block-diagonal batching (our extension, probe-verified
    to be equivalent to per-sample evaluation): concatenate node arrays, offset
    each graph's edge_index by the running node count, concatenate along dim 1."""
    off = 0
    eis = []
    for x, ei in zip(xs, edge_indexes):
        eis.append(np.asarray(ei, dtype=np.int64) + off)
        off += len(x)
    return np.concatenate(xs, 0), np.concatenate(poss, 0), np.concatenate(eis, 1)


def csr_by_key(key: np.ndarray, n_rows: int):
    """Stable CSR of edge ids grouped by ``key`` (destination or source node).

    Within a row the edge ids ascend - the order in which the CPU reference's
    ``index_add_`` / ``scatter_add_`` visits them (models/GNN.py:20, 99), which is
    the summation order the CUDA aggregation must reproduce.
    Returns int32 ``rowptr [n_rows + 1]`` and int32 ``eid [E]``.
    """
    key = np.asarray(key, dtype=np.int64)
    eid = np.argsort(key, kind="stable").astype(np.int32)
    counts = np.bincount(key, minlength=n_rows)
    rowptr = np.zeros(n_rows + 1, dtype=np.int32)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr, eid

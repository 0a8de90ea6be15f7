"""Host-side mirror of the reference's tree initialisation by recursive clustering on top of the C ABI.

tdm/src/main/scala/com/mass/tdm/cluster/RecursiveCluster.scala:16-62 (``RecursiveCluster(...).run(outputTreePath)``): item
embeddings -> recursive balanced 2-means bisection (dmg_kmeans_tree: the clustering and the distances on the GPU, the median
quickselect on the host) -> TreeBuilder.build.  Only clusterType = "kmeans" is built (the spectral variant is the reference's own
Java class over smile's eigen solver; `parallel` / `numThreads` choose a thread pool and do not change the result).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from ._capi import Engine
from .formats import tree_file


def read_file(embed_path: str, delimiter: str = ",") -> Tuple[np.ndarray, np.ndarray]:
    """RecursiveCluster.readFile (:128-140): one line per item, id first, then the embedding."""
    ids, embeds = [], []
    with open(embed_path) as f:
        for line in f:
            parts = line.rstrip("\n").split(delimiter)
            if not parts or not parts[0].strip():
                continue
            ids.append(int(parts[0].strip()))
            embeds.append([float(x.strip()) for x in parts[1:]])
    return np.asarray(ids, np.int32), np.asarray(embeds, np.float64)


class RecursiveCluster:
    def __init__(self, ids, embeddings, cluster_iter_num: int, cluster_type: str = "kmeans", engine: Optional[Engine] = None,
                 device: int = 0, seed: int = 0):
        if cluster_type != "kmeans":
            raise ValueError("clusterType must be 'kmeans' here ('spectral' is not built)")          # RecursiveCluster.scala:26-32
        self.ids = np.asarray(ids, np.int32)
        self.embeddings = np.ascontiguousarray(embeddings, np.float64)
        self.cluster_iter_num, self.seed = int(cluster_iter_num), int(seed)
        self.engine = engine or Engine(device)

    @classmethod
    def from_file(cls, embed_path: str, cluster_iter_num: int, **kw) -> "RecursiveCluster":
        ids, emb = read_file(embed_path)
        return cls(ids, emb, cluster_iter_num, **kw)

    def run(self, output_tree_path: Optional[str] = None) -> Tuple[np.ndarray, np.ndarray]:
        """-> (ids, codes) as RecursiveCluster.run; writes the tree file when a path is given (TreeBuilder.build)"""
        codes = self.engine.kmeans_tree(self.embeddings, self.cluster_iter_num, self.seed)
        if output_tree_path:
            tree_file.build_tree(output_tree_path, self.ids, codes)
        return self.ids, codes

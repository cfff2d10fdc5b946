"""freud's neighbour-query hot path on B200 GPUs, behind freud's own Python names.

The sub-modules mirror the reference's (``freud/__init__.py``): ``box``, ``data``, ``locality``, ``density``, ``order``,
``pmft``, ``environment`` -- the classes on the path and the ones SURVEY.md section 8(f) pulls in -- plus ``parallel``
(shard arithmetic and the NCCL plumbing of the multi-GPU runs).  Importing the package loads no native code: the
extension and ``libfreud_b200.so`` are loaded on first use, and every compute fails loudly without a CUDA device.
"""

from . import box, data, density, environment, locality, order, parallel, pmft
from .box import Box
from .locality import AABBQuery, CellQuery, LinkCell, NeighborList

__all__ = ["AABBQuery", "Box", "CellQuery", "LinkCell", "NeighborList", "box", "data", "density", "environment",
           "locality", "order", "parallel", "pmft"]

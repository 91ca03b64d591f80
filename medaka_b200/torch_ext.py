"""Batch container for inference: the part of medaka/torch_ext.py the hot path uses (:102-173).

``Batch.collate`` stacks the per-window feature matrices into one float32 [B,T,F] tensor
(reference: ``torch.stack([...]).float()``, torch_ext.py:155).  Here the stack is written
straight into page-locked host memory when a pinned staging pool is supplied, so the
engine's H2D copy is a single asynchronous DMA instead of the reference's pageable
``.to(device)`` (medaka/models.py:309).
"""
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class Batch:
    """Batch of samples (inference fields only)."""

    read_level_features: torch.Tensor = None
    counts_matrix: torch.Tensor = None
    labels: torch.Tensor = None
    majority_vote_probs: torch.Tensor = None

    @classmethod
    def collate(cls, samples, counts_matrix=False, out=None):
        """Construct a batch from a list of `Sample` objects (2-D counts-matrix or 3-D read-level features).

        :param out: optional float32 array [len(samples), T, F] to fill (e.g. pinned memory).
        """
        first = samples[0].features
        if first.ndim == 3:
            # read-level features [positions, reads, features]: padded with empty reads to the deepest sample of the
            # batch, uint8 like the reference (torch_ext.py:127-141)
            npos, _, nfeats = first.shape
            depths = [s.features.shape[1] for s in samples]
            padded = np.zeros((len(samples), npos, max(depths), nfeats), dtype=np.uint8)
            for i, s in enumerate(samples):
                if s.features.shape[0] != npos or s.features.shape[2] != nfeats:
                    raise RuntimeError("read-level samples of one batch must share positions and feature length")
                padded[i, :, :depths[i], :] = s.features
            fields = {"read_level_features": torch.from_numpy(padded)}
            if samples[0].labels is not None:
                fields["labels"] = torch.stack([torch.from_numpy(np.asarray(s.labels)) for s in samples])
            return cls(**fields)
        if first.ndim != 2:
            raise ValueError(
                f"Unknown feature dimension {first.ndim}. Expect 3 for"
                "read level features or 2 for counts matrices.")
        shape = (len(samples),) + tuple(first.shape)
        for s in samples:
            if tuple(s.features.shape) != tuple(first.shape):
                raise RuntimeError("stack expects each tensor to be equal size, but got {} and {}".format(
                    list(first.shape), list(s.features.shape)))
        if out is None:
            out = np.empty(shape, dtype=np.float32)
        elif tuple(out.shape) != shape or out.dtype != np.float32:
            raise ValueError("out must be float32 of shape {}".format(shape))
        for i, s in enumerate(samples):
            out[i] = s.features          # converts to float32 like .float()
        fields = {"counts_matrix": torch.from_numpy(out)}
        if samples[0].labels is not None:
            fields["labels"] = torch.stack([torch.from_numpy(np.asarray(s.labels)) for s in samples])
        return cls(**fields)

    @property
    def features(self):
        """Return the features tensor."""
        if self.read_level_features is None:
            return self.counts_matrix
        return self.read_level_features

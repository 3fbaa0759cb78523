"""Cache of values derived from conditioning tensors (the cross-attention K|V of every DiT block).

Those values depend only on the prompt / image embeddings and the weights (reference
wan/modules/model.py:202-229 recomputes them in every forward; longcat attention.py:211-276 too),
so one sampling run needs them once per distinct embedding.  An entry is found again by

  1. storage identity: same data pointer, shape, strides, dtype and version counter as the tensor
     the entry was built from.  Every entry HOLDS its source tensors, so the caching allocator
     cannot hand their storage to another tensor while the entry lives - a pointer match is a
     content match, not an address coincidence;
  2. content: a fresh tensor of the same shape whose bytes equal a cached source (a pipeline that
     re-uploads the same embeddings every step).  This costs one device comparison of <= 4 MB.

A different prompt of the same shape therefore always misses, even when it lands on a recycled
address.
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence, Tuple

import torch


def _sig(t: torch.Tensor) -> Tuple:
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype, t.device, t._version)


class ContextCache:
    def __init__(self, capacity: int = 4):
        self.capacity = capacity
        self._entries: List[Tuple[Tuple[torch.Tensor, ...], Tuple, Any]] = []
        self.hits_identity = 0
        self.hits_content = 0
        self.misses = 0

    def __len__(self) -> int:
        return len(self._entries)

    def clear(self) -> None:
        self._entries.clear()

    def get(self, srcs: Sequence[torch.Tensor]) -> Optional[Any]:
        sig = tuple(_sig(t) for t in srcs)
        for held, hsig, value in self._entries:
            if hsig == sig and tuple(_sig(t) for t in held) == sig:     # the held tensors were not written since
                self.hits_identity += 1
                return value
        for held, hsig, value in self._entries:
            if tuple(_sig(t) for t in held) != hsig:                    # source mutated in place after caching: stale
                continue
            if all(h.shape == t.shape and h.dtype == t.dtype and h.device == t.device for h, t in zip(held, srcs)) \
                    and all(torch.equal(h, t) for h, t in zip(held, srcs)):
                self.hits_content += 1
                return value
        self.misses += 1
        return None

    def put(self, srcs: Sequence[torch.Tensor], value: Any) -> Any:
        self._entries = [e for e in self._entries if tuple(_sig(t) for t in e[0]) == e[1]]   # drop stale entries
        while len(self._entries) >= self.capacity:
            self._entries.pop(0)
        held = tuple(srcs)
        self._entries.append((held, tuple(_sig(t) for t in held), value))
        return value

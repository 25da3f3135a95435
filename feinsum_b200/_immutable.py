"""
A tiny hashable, immutable mapping.

The reference front-end stores its derived tables (``index_to_dim_length``,
``arg_to_shape`` ...) in ``immutables.Map`` (reference
``src/feinsum/einsum.py:19,212-272``).  That wheel is not part of this image,
and the only features the hot path relies on are: Mapping protocol, hashing,
value equality and a functional ``update``/``set``/``delete``.  This class
provides exactly those.
"""

from __future__ import annotations

from collections.abc import Iterable, Iterator, Mapping
from typing import Any, Generic, TypeVar

K = TypeVar("K")
V = TypeVar("V")


class Map(Mapping[K, V], Generic[K, V]):
    __slots__ = ("_d", "_h")

    def __init__(self, items: Mapping[K, V] | Iterable[tuple[K, V]] = (), **kw: V):
        d: dict[Any, Any] = dict(items)
        d.update(kw)
        object.__setattr__(self, "_d", d)
        object.__setattr__(self, "_h", None)

    def __setattr__(self, name: str, value: Any) -> None:
        raise AttributeError("Map is immutable")

    def __getitem__(self, key: K) -> V:
        return self._d[key]  # type: ignore[no-any-return]

    def __iter__(self) -> Iterator[K]:
        return iter(self._d)

    def __len__(self) -> int:
        return len(self._d)

    def __contains__(self, key: object) -> bool:
        return key in self._d

    def __hash__(self) -> int:
        h = self._h
        if h is None:
            h = hash(frozenset(self._d.items()))
            object.__setattr__(self, "_h", h)
        return h  # type: ignore[no-any-return]

    def __eq__(self, other: object) -> bool:
        if isinstance(other, Map):
            return self._d == other._d
        if isinstance(other, Mapping):
            return self._d == dict(other)
        return NotImplemented

    def __repr__(self) -> str:
        return f"Map({self._d!r})"

    # functional updates -------------------------------------------------
    def update(self, other: Mapping[K, V] | Iterable[tuple[K, V]] = (), **kw: V) -> "Map[K, V]":
        d = dict(self._d)
        d.update(other)
        d.update(kw)
        return Map(d)

    def set(self, key: K, value: V) -> "Map[K, V]":
        d = dict(self._d)
        d[key] = value
        return Map(d)

    def delete(self, key: K) -> "Map[K, V]":
        d = dict(self._d)
        del d[key]
        return Map(d)

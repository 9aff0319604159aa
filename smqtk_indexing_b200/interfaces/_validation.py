"""Argument validation shared by the interface template methods."""
import itertools
from typing import Callable, Iterable, TypeVar

R = TypeVar("R")


def call_if_not_empty(items: Iterable, fn: Callable[[Iterable], R], err: BaseException) -> R:
    """Peek one element of ``items``; raise ``err`` when there is none, else
    hand ``fn`` an iterable equivalent to the original one.

    Mirrors ``check_empty_iterable`` (reference:
    smqtk_indexing/utils/iter_validation.py:8-28): works for one-shot iterators
    without consuming them.
    """
    it = iter(items)
    sentinel = object()
    head = next(it, sentinel)
    if head is sentinel:
        raise err
    return fn(itertools.chain((head,), it))

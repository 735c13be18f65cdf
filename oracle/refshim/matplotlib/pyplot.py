"""Stub: attribute access returns a no-op callable."""
import sys as _sys


class _Nop(object):
    def __call__(self, *a, **k):
        return _Nop()

    def __getattr__(self, n):
        return _Nop()

    def __iter__(self):
        return iter((_Nop(), _Nop()))


def __getattr__(name):
    return _Nop()

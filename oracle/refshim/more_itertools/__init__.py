"""Stub (test infrastructure): more_itertools.locate as used at extract_track_candidates.py:94."""


def locate(iterable, pred=bool):
    for i, v in enumerate(iterable):
        if pred(v):
            yield i

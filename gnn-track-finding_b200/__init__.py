"""gtf_b200 -- B200-native Gaussian-mixture message passing for track finding.

Drop-in for the message-passing hot path of nishalad95/GNN-track-finding."""
from . import fields, synth, nxio, lib  # noqa: F401
from .batch import EventBatch  # noqa: F401
from . import stages  # noqa: F401

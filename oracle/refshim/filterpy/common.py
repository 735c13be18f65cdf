"""filterpy.common shim: only what the reference touches (Saver, reshape_z)."""
import numpy as np


class Saver(object):
    """No-op stand-in for filterpy.common.Saver (extract_track_candidates.py:219,230,288,308)."""

    def __init__(self, kf, *args, **kwargs):
        self._kf = kf

    def save(self):
        pass


def reshape_z(z, dim_z, ndim):
    """filterpy 1.4.5 common.reshape_z: make z (dim_z,1), then match x.ndim."""
    z = np.atleast_2d(z)
    if z.shape[1] == dim_z:
        z = z.T
    if z.shape != (dim_z, 1):
        raise ValueError('z must be convertible to shape ({}, 1)'.format(dim_z))
    if ndim == 1:
        z = z[:, 0]
    if ndim == 0:
        z = z[0, 0]
    return z

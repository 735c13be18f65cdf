"""Empty stub (test infrastructure): the reference imports matplotlib.pyplot at module scope only."""

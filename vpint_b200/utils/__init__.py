"""Callers either side of the propagation path (reference: VPint/utils/), kept where they only need NumPy and
the device solve.  The loaders, plots and metrics of the reference's utils package are out of scope."""

"""metsim_b200 -- B200-native implementation of MetSim's per-cell MTCLIM hot path.

Public surface (mirrors the reference's interface for this path):

* ``metsim_b200.methods.mtclim_b200.run(df, params)`` -- drop-in for
  ``metsim.methods.mtclim.run`` (metsim/methods/mtclim.py:28), registered under the key
  ``'mtclim_b200'`` by :func:`metsim_b200.register`.
* :class:`metsim_b200.engine.Engine` -- the batched form of ``MetSim.run_slice``
  (metsim/metsim.py:456-493): all cells of a shard at once, device-resident.
* :func:`metsim_b200.engine.run_host` -- same through ``msb_run_host`` with host buffers.

All arithmetic happens in ``csrc/libmetsim_b200.so`` (hand-written sm_100a CUDA behind the C ABI
of ``include/metsim_b200.h``); importing the compute entry points without it raises.
"""
from .registry import register  # noqa: F401

__all__ = ["register"]
__version__ = "0.1.0"

"""idefix_b200 -- B200-native explicit Godunov (M)HD update behind Idefix's own interface.

The grid work lives in ``libidefix_b200.so`` (hand-written sm_100a CUDA, csrc/) behind the C ABI
of ``include/idefix_b200.h``; this package is the host-side mirror of the reference's
Input / Grid / DataBlock / Fluid / TimeIntegrator / Dump objects for that path.
"""
from . import _abi
from ._abi import IdefixError
from .datablock import DataBlock, DataBlockHost, Fluid
from .dump import readDump, writeDump
from .grid import Definitions, Grid
from .input import Input
from .timeintegrator import TimeIntegrator, run

__all__ = ["_abi", "IdefixError", "DataBlock", "DataBlockHost", "Fluid", "readDump", "writeDump",
           "Definitions", "Grid", "Input", "TimeIntegrator", "run"]

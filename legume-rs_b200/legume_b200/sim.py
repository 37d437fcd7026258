"""Synthetic single-cell counts in the shape of `data-beans-sim topic`
(data-beans-sim/src/core.rs:253-345): log-normal dictionary, log-space batch effects, one-hot topic
proportions softened by pve_topic, Poisson counts at rate (depth/D) * delta * sum_k beta*theta.

The per-(topic, batch, gene) rate tables are built here with numpy; the Poisson draws themselves are
made on the GPU by lg_sim_poisson_csc (and identically by the CPU twin in oracle/) from a
counter-based hash of (seed, cell, gene), so any column range can be generated on any rank.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import CscBlock, Context, _ptr
from ._lib import lib

from .sim_tables import SimTables, _mix64, make_tables  # noqa: F401  (pure-numpy half, importable on its own)


def sim_block(ctx: Context, tables: SimTables, col_lo: int, col_hi: int):
    """generate columns [col_lo, col_hi) on the device; returns (CscBlock, topic, batch)"""
    import torch
    topic, batch = tables.cell_labels(col_lo, col_hi)
    dlam, dp0, dnp = tables.device_tables(ctx.device)
    dt = torch.from_numpy(topic).to(dlam.device)
    db = torch.from_numpy(batch).to(dlam.device)
    h = C.c_void_p()
    ctx.check(lib.lg_sim_poisson_csc(ctx.h, tables.seed, tables.D, col_lo, col_hi, _ptr(dt), _ptr(db), tables.ntopic,
                                     tables.nbatch, _ptr(dlam), _ptr(dp0), _ptr(dnp), C.byref(h)))
    return CscBlock(ctx, h), topic, batch

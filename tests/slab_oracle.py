"""Test-only mirror of smokephysai_b200.slab.SlabNavierStokes with the CPU oracle as the compute backend:
the same SlabGeometry, step order and halo exchanges, dense numpy arrays instead of the device arena.
Lets the decomposition logic (ghost depth vs erosion, exchange plan, P2P over gloo) be checked without a GPU."""
import numpy as np
import torch

import oracle
from smokephysai_b200.slab import SlabGeometry, build_step_plan, needs_single_exchange


class OracleSlab:
    def __init__(self, grid_size, dt, viscosity, rank, world, K, T, halo=None):
        H, W = grid_size
        self.T = T
        self.halo = halo if halo is not None else T + 4
        self.geom = SlabGeometry(H, W, world, rank, self.halo if world > 1 else 0)
        g = self.geom
        self.dt, self.nu, self.K, self.world = dt, viscosity, K, world
        self.f = {"u": np.zeros((g.hl + 1, W), np.float32), "v": np.zeros((g.hl, W + 1), np.float32),
                  "d": np.zeros((g.hl, W), np.float32), "p": np.zeros((g.hl, W), np.float32)}
        self.div = None

    def scatter(self, name, glob):
        rows = self.f[name].shape[0]
        self.f[name][...] = glob[self.geom.A: self.geom.A + rows]

    def owned(self, name):
        lo, hi = self.geom.owned_rows("u" if name == "u" else "c")
        return self.f[name][lo:hi]

    def exchange_list(self, names):
        return [(torch.from_numpy(self.f[n]), "u" if n == "u" else "c") for n in names]

    def fdd(self):
        f, dt, nu = self.f, self.dt, self.nu
        f["v"] = oracle.buoyancy(f["v"], f["d"], dt)
        f["u"] = oracle.diffusion_step(f["u"], dt, nu)
        f["v"] = oracle.diffusion_step(f["v"], dt, nu)
        f["d"] = oracle.diffusion_step(f["d"], dt, nu * 0.1)
        self.div = oracle.divergence(f["u"], f["v"], dt)

    def jacobi(self, t):
        self.f["p"] = oracle.jacobi(self.f["p"], self.div, t)

    def project_advect(self):
        f, dt, g = self.f, self.dt, self.geom
        f["u"], f["v"] = oracle.grad_subtract(f["u"], f["v"], f["p"], dt)
        f["u"] = oracle.advection_step_slab(f["u"], f["u"], f["v"], dt, g.A, g.H)
        f["v"] = oracle.advection_step_slab(f["v"], f["u"], f["v"], dt, g.A, g.H)
        f["d"] = oracle.advection_step_slab(f["d"], f["u"], f["v"], dt, g.A, g.H) * np.float32(0.995)

    def step_plan(self):
        single = needs_single_exchange(self.world, self.halo, self.K)
        return build_step_plan(self.world, self.K, self.T, single, self.fdd, self.jacobi, self.project_advect)

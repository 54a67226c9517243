"""Slab domain decomposition of one large supercell over the GPUs of one box
(BASELINE config 3).

One process per GPU.  The N0 x N1 x N2 box is cut into `world` slabs along k;
each rank owns N2/world layers plus `halo` ghost layers on each side.  A sweep
runs either
  * over NVLink peer memory (the default where the ring neighbours can map each other,
    `cmx_state_ipc_attach`): `cmx_sgc_sweep_slab`, the streaming kernel stores the
    boundary rows it changes into the neighbours' ghost layers and counts them on the
    neighbours' layer counters -- no collective in the data path; or
  * k-colour group by k-colour group (`cmx_sgc_sweep_kgroup`, asynchronous on the
    state's stream): after group g the boundary layers whose global k is congruent to g
    are the only ones that changed, and exactly those are sent to the ring neighbours
    (NCCL send/recv, enqueued on the same stream, no host synchronisation in a sweep).
The initial fill of the ghost layers after an upload is always the NCCL exchange.

The RNG counters use GLOBAL coordinates (`cmx_state_set_k_offset`), therefore
the trajectory is bit-identical for every number of slabs -- which the tests
use as the parity check of the decomposition (1 slab == 2 slabs == ...).

torch is used for what it is here for: NCCL plumbing and stream handles.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _capi


def halo_ops(rank: int, world: int, n2: int, halo: int, Sk: int, N2: int, kgroup: Optional[int]):
    """Halo messages of one rank after k-colour group `kgroup` (all layers if None).

    Returns a list of (kind, local_layer, peer) with kind "send"/"recv" and
    local_layer in [-halo, n2+halo): owned layers are 0..n2-1, ghosts below 0 and
    from n2 up.  Only boundary layers whose GLOBAL k is congruent to kgroup
    (mod Sk) changed in that group, so only those travel.  Sends and receives of
    neighbouring ranks pair up one to one, in order.
    """
    k0 = rank * n2
    up, dn = (rank + 1) % world, (rank - 1) % world
    ops = []
    for q in range(halo):
        # my low boundary layer q -> lower neighbour's high ghost n2+q
        if kgroup is None or (k0 + q) % Sk == kgroup:
            ops.append(("send", q, dn))
        # upper neighbour's low boundary layer q -> my high ghost n2+q
        if kgroup is None or ((k0 + n2) % N2 + q) % Sk == kgroup:
            ops.append(("recv", n2 + q, up))
        # my high boundary layer n2-halo+q -> upper neighbour's low ghost -halo+q
        if kgroup is None or (k0 + n2 - halo + q) % Sk == kgroup:
            ops.append(("send", n2 - halo + q, up))
        if kgroup is None or ((k0 - halo + q) % N2) % Sk == kgroup:
            ops.append(("recv", -halo + q, dn))
    return ops


class _DevMem:
    """Zero-copy torch view of library-owned device memory."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|i1", "data": (ptr, False),
                                         "version": 2, "strides": None}


class SlabRunner:
    def __init__(self, tables: _capi.Tables, N, eci, temperature: float, exch, rank: int, world: int,
                 local_rank: int, seed_init: int = 1, backend_device: Optional[str] = "cuda",
                 init_occ: Optional[np.ndarray] = None, p2p: bool = True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        if np.isscalar(N):
            N = (N, N, N)
        self.N = tuple(int(x) for x in N)
        self.rank, self.world = rank, world
        if self.N[2] % world:
            raise _capi.CmxError(_capi.CMX_ERR_INVALID, "N2 must be divisible by the number of slabs")
        self.n2 = self.N[2] // world
        self.k0 = rank * self.n2
        # a probe state (no halo) tells us the colouring the bound ECI need
        probe = _capi.State(tables, (self.N[0], self.N[1], max(self.n2, 2)))
        probe.set_eci(eci["index"], eci["value"])
        info = probe.sweep_info()
        probe.close()
        self.Sk = info["colour_strides"][2]
        self.halo = max(1, info["range_k"])
        if self.n2 % self.Sk or self.n2 < self.halo:
            raise _capi.CmxError(_capi.CMX_ERR_INVALID,
                                 f"slab thickness {self.n2} incompatible with k-colour stride {self.Sk}")
        self.state = _capi.State(tables, (self.N[0], self.N[1], self.n2), 1, halo=self.halo)
        self.state.set_eci(eci["index"], eci["value"])
        self.state.set_conditions(temperature, exch)
        self.state.set_k_offset(self.k0)
        self._info = self.state.sweep_info()
        self.n_sublat = tables.host.n_sublat
        self.layer = self.N[0] * self.N[1]
        ptr, nbytes = self.state.device_ptr()
        self.mem = torch.as_tensor(_DevMem(ptr, nbytes), device=f"cuda:{local_rank}")
        self.stream = torch.cuda.ExternalStream(self.state.stream())
        self.up = (rank + 1) % world
        self.dn = (rank - 1) % world
        # initial occupation: global i.i.d. random from the host (identical for any
        # decomposition), or a given global array
        n_cells_g = self.N[0] * self.N[1] * self.N[2]
        if init_occ is None:
            init_occ = np.random.default_rng(seed_init).integers(0, 3, n_cells_g * self.n_sublat).astype(np.int8)
        self.upload_global(init_occ)
        # halo exchange fused into the sweep kernel: every rank maps its ring
        # neighbours' slabs (CUDA IPC over NVLink) and the kernel stores the
        # boundary rows it changes straight into their ghost layers
        self.p2p = False
        if p2p:
            self._attach_peers()

    def _attach_peers(self) -> None:
        torch, dist = self.torch, self.dist
        if self.world == 1:
            self.state.ipc_attach(None, None)
        else:
            mine = torch.frombuffer(bytearray(self.state.ipc_export()), dtype=torch.uint8).to(self.mem.device)
            allh = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(allh, mine)
            hb = [bytes(h.cpu().numpy().tobytes()) for h in allh]
            self.stream.synchronize()
            dist.barrier()
            self.state.ipc_attach(hb[self.dn], hb[self.up])
            dist.barrier()  # nobody pushes before everybody has mapped and zeroed its flags
        self.p2p = self.state.p2p_active()

    # -- layout helpers -------------------------------------------------------
    def _layers(self, b: int, k_lo: int, k_hi: int):
        """torch view of local layers [k_lo, k_hi) (k = -halo .. n2+halo) of sublattice b."""
        sub = self.layer * (self.n2 + 2 * self.halo)
        a = b * sub + (k_lo + self.halo) * self.layer
        return self.mem[a:a + (k_hi - k_lo) * self.layer]

    def upload_global(self, occ_global: np.ndarray) -> None:
        n_cells_g = self.N[0] * self.N[1] * self.N[2]
        n_loc = self.layer * self.n2
        local = np.empty(n_loc * self.n_sublat, dtype=np.int8)
        for b in range(self.n_sublat):
            g = occ_global[b * n_cells_g:(b + 1) * n_cells_g]
            local[b * n_loc:(b + 1) * n_loc] = g[self.k0 * self.layer:(self.k0 + self.n2) * self.layer]
        self.state.upload_occ(local)
        self.exchange(None)

    def download_local(self) -> np.ndarray:
        return self.state.download_occ(dtype=np.int8)

    def gather_global(self) -> Optional[np.ndarray]:
        """Global occupation on rank 0 (tests)."""
        torch, dist = self.torch, self.dist
        loc = torch.from_numpy(self.download_local()).cuda()
        parts = [torch.empty_like(loc) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(loc, parts, dst=0)
        if self.rank != 0:
            return None
        n_loc = self.layer * self.n2
        n_cells_g = self.layer * self.N[2]
        out = np.empty(n_cells_g * self.n_sublat, dtype=np.int8)
        for r, p in enumerate(parts):
            p = p.cpu().numpy()
            for b in range(self.n_sublat):
                out[b * n_cells_g + r * n_loc:b * n_cells_g + (r + 1) * n_loc] = p[b * n_loc:(b + 1) * n_loc]
        return out

    # -- halo exchange --------------------------------------------------------
    def exchange(self, kgroup: Optional[int]) -> None:
        """Send the boundary layers that changed in `kgroup` (all if None) to the
        ring neighbours; enqueued on the state's stream."""
        torch, dist = self.torch, self.dist
        sched = halo_ops(self.rank, self.world, self.n2, self.halo, self.Sk, self.N[2], kgroup)
        if not sched:
            return
        with torch.cuda.stream(self.stream):
            if self.world == 1:
                self._self_exchange(kgroup)
                return
            ops = []
            for b in range(self.n_sublat):
                for kind, k, peer in sched:
                    fn = dist.isend if kind == "send" else dist.irecv
                    ops.append(dist.P2POp(fn, self._layers(b, k, k + 1), peer))
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def _self_exchange(self, kgroup):
        # single slab with ghost layers: periodic images are local copies
        h, n2 = self.halo, self.n2
        for b in range(self.n_sublat):
            self._layers(b, n2, n2 + h).copy_(self._layers(b, 0, h))
            self._layers(b, -h, 0).copy_(self._layers(b, n2 - h, n2))

    # -- driver -----------------------------------------------------------------
    def sweep(self, n_sweeps: int, seed: int, first_sweep: int = 0) -> None:
        """Asynchronous: everything is enqueued on the state's stream."""
        if self.p2p:
            # streaming kernel over peer memory: boundary rows and their completion counts go
            # straight into the ring neighbours' ghost layers / layer counters
            if n_sweeps > 0:
                self.state.sgc_sweep_slab(n_sweeps, seed, first_sweep)
            return
        for w in range(n_sweeps):
            for g in range(self.Sk):
                self.state.sgc_sweep_kgroup(seed, first_sweep + w, g)
                self.exchange(g)

    def synchronize(self):
        self.stream.synchronize()

    def info(self) -> dict:
        return self._info

    def counters(self):
        """Global (attempt, accept, dE_sum) summed over slabs."""
        torch, dist = self.torch, self.dist
        c = self.state.counters_read()[0]
        t = torch.tensor([float(c.n_attempt), float(c.n_accept), c.dE_sum], dtype=torch.float64,
                         device=self.mem.device)
        dist.all_reduce(t)
        return t.cpu().numpy()

"""Resident throughput path: the electrodynamic step of PyPIC3D/evolve.py:16-103 on compact, cell-sorted SoA particles.

`Simulation` takes the reference's pytrees (`TiledParticles`, `SpeciesConfig`, the `fields` 8-tuple, Static/Dynamic
parameters) for ONE tile per GPU, keeps the state resident in HBM in the kernels' native layout and advances it with

    K1  pic_fused_push_deposit   gather -> push -> deposit -> move -> particle BC     (one pass over particle memory)
        halo fold + refresh of J  (particle BCs)                                      ghost_cells.py:703-736
    K6  pic_update_B (half)  K5 pic_update_E  K6 pic_update_B (half) + refreshes      first_order_yee.py:12-162
    K2  counting sort by cell every `sort_interval` steps

`export_state()` hands the reference pytrees back (same shapes / slot identity on a single GPU).
Host orchestration is Python; every array operation on the step is a hand-written kernel behind the C ABI.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib, ops
from ._lib import PicSoA, PicLeave, PicError, check
from .particles.particle_class import TiledParticles
from .boundary_conditions.grid_and_stencil import BC_CONDUCTING


class _Species:
    __slots__ = ("buf", "ids", "cur", "n", "cap", "n_dev", "leave", "recv", "leave_buf", "recv_buf", "blk_off")


class LocalHalo:
    """Guard-cell exchange when every axis is local to this GPU (periodic self-exchange / walls / reduced axes)."""

    def __init__(self, params):
        self.p = params

    def fold_(self, fields, bcs):
        ops.halo_fold_(self.p, fields, bcs)

    def refresh_(self, fields, bcs, skip_axes=()):
        for axis in range(3):
            if axis not in skip_axes:
                ops.halo_refresh_axis_(self.p, axis, int(bcs[axis]), fields)

    def migrate(self, sim):
        return None

    def allreduce_max(self, values):
        return list(values)


class Simulation:
    def __init__(self, particles, species_config, fields, static_parameters, dynamic_parameters, *, sort_interval=10,
                 capacity_factor=1.25, track_ids=True, halo=None, gmesh=None, moff=(0, 0, 0), leave_fraction=0.25):
        sp, dp = static_parameters, dynamic_parameters
        if getattr(sp, "pml_active", False):
            raise NotImplementedError("PML is outside the hot path of pypic3d_b200")
        x = ops._chk(particles.x, "particles.x")
        if tuple(x.shape[:3]) != (1, 1, 1):
            raise ValueError("Simulation holds one tile per GPU; use pypic3d_b200.distributed for a multi-tile mesh "
                             f"(got particle tile topology {tuple(x.shape[:3])})")
        self.sp, self.dp, self.species_config = sp, dp, species_config
        self.dtype = x.dtype
        self.device = x.device
        gm = (1, 1, 1) if gmesh is None else tuple(int(v) for v in gmesh)
        self.p = ops.params_for(sp, dp, species_config, x, mesh=(1, 1, 1), gmesh=gm, moff=moff)
        # float64 twin of the parameter block: the conservation diagnostics of a float32 run are evaluated in float64
        self.p64 = _lib.make_params(sp, dp, species_config, np.float64, mesh=(1, 1, 1), gmesh=gm, moff=moff)
        self._halo64 = None
        self.S = int(self.p.n_species)
        self.deposition = 0 if sp.current_deposition == "esirkepov" else 1
        self.current_filter = sp.current_filter if self.deposition == 1 else "none"
        self.alpha = float(self.p.alpha)
        self.g = int(self.p.g)
        if self.g < 1:
            raise ValueError("guard_cells must be >= 1")
        self.sort_interval = int(sort_interval)
        self.step_count = 0
        self._J_ghosts_stale = False
        if halo is None:
            self.halo = LocalHalo(self.p)
        else:
            self.halo = halo(self.p) if callable(halo) and not hasattr(halo, "refresh_") else halo
        # the field half of the step in one kernel (pic_yee_fused) unless the digital filter is on; PIC_YEE=sweeps keeps the three
        # sweeps + refreshes (A/B control, same result bit for bit)
        self._yee_fused = (os.environ.get("PIC_YEE", "fused") == "fused") and float(self.p.alpha) == 1.0 and int(self.p.g) >= 2
        self._E2 = self._B2 = self._J2 = None
        self._swapped = []        # buffer swaps performed by the current step ("J", "EB"): replayed after a CUDA-graph replay
        self._graphs = {}         # (sort buffers, K1 options) -> (CUDAGraph, swaps)
        self._graph_warm = 0
        self.k1_events = None   # set to [] to record (start, end) CUDA events around every K1 launch (bench.py roofline)
        self.phase_events = None   # set to {} to record CUDA events around the phases of every step (bench.py `phases_ms`)
        self.distributed = any(self.p.gmesh[a] != self.p.mesh[a] for a in range(3))
        E, B, J, rho, phi, ext, pml_state, overflow = fields
        if pml_state is not None:
            raise NotImplementedError("PML is outside the hot path of pypic3d_b200")
        self._passthrough = (rho, phi, ext)
        clone = lambda t: ops._chk(t, "field", self.dtype).clone()
        self.E = [clone(c) for c in E]
        self.B = [clone(c) for c in B]
        self.J = [clone(c) for c in J]
        ext_E, ext_B = ext
        has_ext = any(bool(torch.any(c != 0)) for c in tuple(ext_E) + tuple(ext_B))
        self.ext_E = [clone(c) for c in ext_E] if has_ext else None
        self.ext_B = [clone(c) for c in ext_B] if has_ext else None
        self.overflow_previous = bool(overflow) if not isinstance(overflow, torch.Tensor) else bool(overflow.item())
        self.flags = torch.zeros(4, dtype=torch.int32, device=self.device)
        self.ncells = int(self.p.tile[0]) * int(self.p.tile[1]) * int(self.p.tile[2])
        self._cell_count = torch.zeros(self.ncells + 1, dtype=torch.int32, device=self.device)
        self._cell_offset = torch.zeros(self.ncells + 1, dtype=torch.int32, device=self.device)
        self._scan_scratch = torch.zeros((self.ncells + 1 + 2047) // 2048 + 1, dtype=torch.int32, device=self.device)
        self._blk_work = torch.zeros(self.ncells // 64 + 2, dtype=torch.int32, device=self.device)
        self._counter = torch.zeros(max(self.S, 1) + 27, dtype=torch.int32, device=self.device)
        self.cap_ref = int(x.shape[4])
        self.track_ids = bool(track_ids)
        self.k1_variant = self._pick_k1_variant(sp)
        self._import(particles, capacity_factor)
        self.sort_every = self._sort_intervals(particles)
        # K1 v9 options (include/pic_b200.h): bit 0 = shared-memory J tiles + TMA reduce (PIC_K9_JTILE = "1"; measured slower, off);
        # bit 1 = match-any group reduction instead of the segmented scan.  PIC_K9_GROUPRED = "auto" (default): per launch, for a
        # species that has drifted more than 4 % of a cell (rms) since its last sort -- its same-cell runs are fragmented by then
        # and the scan would pay one RED set per fragment; "1" always, "0" never.
        self._jtile_mode = os.environ.get("PIC_K9_JTILE", "0")
        self._groupred_mode = os.environ.get("PIC_K9_GROUPRED", "auto")
        self._red_mode = os.environ.get("PIC_K10_RED", "scan")   # (measured: scan 3.10 ms, smem 3.23 ms per proton launch)
        # K1 v10 work list of deferred slots (include/pic_b200.h pic_fused_pair3d): shared by the species, launches are stream-ordered
        self._pair_work = None
        if self.k1_variant == "pair":
            nbytes = int(_lib.lib().pic_pair_work_bytes(ctypes.byref(self.p), max(sp_.cap for sp_ in self.species)))
            self._pair_work = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._sorted_at = [0] * self.S
        self.leave_fraction = float(leave_fraction)
        if self.distributed:
            self._alloc_packets()
        self._refresh_imported_fields()
        self.sort()

    # ------------------------------------------------------------------------------------------ layout
    def _pick_k1_variant(self, sp):
        """"pair": K1 v10 (pic_fused_pair3d: supercell E/B tiles in shared memory, two particles per thread in packed f32x2
        arithmetic, fed by the blocked + padded sort) when the configuration is the one the tile kernels were built for;
        "tile": K1 v9 (pic_fused_tile3d), same configurations, kept as the A/B control (PIC_K1_VARIANT=tile);
        "global": pic_fused_push_deposit (every other configuration, or PIC_K1_VARIANT=global).  All give the same result."""
        p = self.p
        ok = (self.deposition == 0 and int(p.shape_factor) == 1 and int(p.g) == 2
              and int(p.pusher) in (0, 1)   # PIC_PUSHER_BORIS, PIC_PUSHER_BORIS_REL
              and all(int(p.tile[a]) % 4 == 0 and int(p.gmesh[a]) * int(p.tile[a]) > 1 for a in range(3))
              and self.ext_E is None)
        want = os.environ.get("PIC_K1_VARIANT", "pair")
        if not ok or want not in ("pair", "tile"):
            return "global"
        return want

    def _k1_options(self, s):
        opt = 1 if self._jtile_mode == "1" else 0
        # K1 v10, float: bit 2 = same-cell reduction through shared memory (pair_smem_red) instead of the segmented warp scan;
        # PIC_K10_RED = "scan" (default) | "smem"
        if self.k1_variant == "pair" and self._red_mode == "smem":
            opt |= 4
        if self._groupred_mode == "1":
            opt |= 2
        elif self._groupred_mode == "auto":
            dmin = min(float(self.p.dx), float(self.p.dy), float(self.p.dz))
            drift = self._vrms[s] * float(self.p.dt) * (self.step_count - self._sorted_at[s]) / dmin
            if drift > 0.04:
                opt |= 2
        return opt

    def _soa(self, sp_, which=None):
        k = sp_.cur if which is None else which
        s = PicSoA()
        for c in range(6):
            s.comp[c] = sp_.buf[k][c].data_ptr()
        s.id = sp_.ids[k].data_ptr() if sp_.ids is not None else None
        s.cap = sp_.cap
        s.n = sp_.n
        s.n_dev = sp_.n_dev.data_ptr()
        return s

    def _alloc_packets(self):
        """Fixed-size per-direction migration packets (PicLeave).  Capacity of a face packet = the particles that could
        cross that face in one step at light speed from a uniform plasma x `leave_fraction` (thermal plasmas use ~1 % of it);
        edges and corners scale with the product of the per-axis fractions.  Identical on every rank by construction.
        Layout: ONE buffer for all species, packets grouped by the rank they travel to (leave) / come from (recv) -- then by
        species, then by direction code -- so that everything exchanged with one peer is one contiguous message (on a (2,2,2) mesh:
        7 messages per step instead of 52)."""
        p = self.p
        d = (p.dx, p.dy, p.dz)
        frac = [min(1.0, p.C * p.dt / (p.tile[a] * d[a])) for a in range(3)]
        split = [p.gmesh[a] != p.mesh[a] for a in range(3)]
        # packet sizes must agree on every rank: base them on the largest per-species capacity of the whole job
        n0s = self.halo.allreduce_max([max(1, int(sp_.cap)) for sp_ in self.species])
        caps = []
        for n0 in n0s:
            c = [0] * 27
            for dcode in range(27):
                off = (1 - dcode // 9, 1 - (dcode // 3) % 3, 1 - dcode % 3)
                if off != (0, 0, 0) and all(off[a] == 0 or split[a] for a in range(3)):
                    f = 1.0
                    for a in range(3):
                        if off[a] != 0:
                            f *= frac[a]
                    c[dcode] = int(n0 * f * self.leave_fraction) + 1024
            caps.append(c)
        dirs = self.halo.active_dirs(tuple(p.particle_bc)) if hasattr(self.halo, "active_dirs") else [(dc, None, None) for dc in range(27)]

        from .distributed import grouped_packet_layout

        def layout(key):        # key: 1 = destination rank (leave), 2 = source rank (recv)
            row_off, rows, slices = grouped_packet_layout(dirs, caps, key)
            tables = [PicLeave() for _ in self.species]
            used = {dc for dc, _, _ in dirs}
            for s_, t in enumerate(tables):
                for dc in range(27):
                    t.row_off[dc] = row_off[s_][dc]
                    t.cap[dc] = caps[s_][dc] if dc in used else 0
            return tables, rows, slices
        Ls, rows_l, self._send_slices = layout(1)
        Rs, rows_r, self._recv_slices = layout(2)
        self._leave_all = torch.zeros(max(rows_l, 1) * 7, dtype=self.dtype, device=self.device)
        self._recv_all = torch.zeros(max(rows_r, 1) * 7, dtype=self.dtype, device=self.device)
        for sp_, L, R in zip(self.species, Ls, Rs):
            sp_.leave_buf, sp_.recv_buf = self._leave_all, self._recv_all
            L.buf, R.buf = self._leave_all.data_ptr(), self._recv_all.data_ptr()
            sp_.leave, sp_.recv = L, R

    def _import(self, particles, capacity_factor):
        L = _lib.lib()
        x, u, a = ops._chk(particles.x, "particles.x", self.dtype), ops._chk(particles.u, "particles.u", self.dtype), ops._chk(particles.active, "particles.active")
        if tuple(x.shape[:4]) != (1, 1, 1, self.S) or int(x.shape[4]) != self.cap_ref:
            raise ValueError(f"particle layout {tuple(x.shape)} does not match (1,1,1,{self.S},{self.cap_ref},3)")
        counts = a.reshape(self.S, -1).sum(dim=1).tolist() if a.numel() else [0] * self.S
        first = not hasattr(self, "species")
        if first:
            self.species = []
        st = ops._stream()
        self._counter.zero_()
        for s in range(self.S):
            if first:
                sp_ = _Species()
                # a multiple of 64 slots keeps every row of the (6, cap) SoA 16-byte aligned (TMA bulk copies in K1 v9)
                # (+ 4 slots per supercell: the blocked sort of K1 v10 pads every supercell's slice to a multiple of 4 slots)
                pad = 4 * (self.ncells // 64) if self.k1_variant == "pair" else 0
                sp_.cap = (max(16, int(np.ceil(counts[s] * float(capacity_factor))) + 16 + pad) + 63) // 64 * 64
                sp_.buf = [torch.empty((6, sp_.cap), dtype=self.dtype, device=self.device) for _ in range(2)]
                sp_.ids = [torch.empty(sp_.cap, dtype=torch.int32, device=self.device) for _ in range(2)] if self.track_ids else None
                sp_.n_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
                sp_.leave = sp_.recv = sp_.leave_buf = sp_.recv_buf = None
                sp_.blk_off = torch.zeros(self.ncells // 64 + 1, dtype=torch.int32, device=self.device)
                self.species.append(sp_)
            sp_ = self.species[s]
            if counts[s] > sp_.cap:
                raise PicError(f"species {s}: {counts[s]} particles exceed the resident capacity {sp_.cap}")
            sp_.cur = 0
            sp_.n = sp_.cap                       # host-side upper bound; the live count is sp_.n_dev on the device
            sp_.n_dev.zero_()
            soa = self._soa(sp_)
            soa.n_dev = None                      # the import kernel counts into n_dev itself
            check(L.pic_soa_import(ctypes.byref(self.p), s, ops._p(x), ops._p(u), ops._p(a), self.cap_ref, ctypes.byref(soa),
                                   ops._p(sp_.n_dev), st), "pic_soa_import")

    def _sort_intervals(self, particles):
        """Per-species sort cadence: `sort_interval` for the fastest species, proportionally longer (up to 20x) for species
        whose rms displacement per step is smaller -- a heavy, cold species stays cell-sorted far longer than electrons.
        Sorting never changes results (only memory order), so this is purely a cost knob.  Multi-GPU runs use the same cadence:
        the sort is also what compacts migrated slots, but a slow species receives proportionally few arrivals per step (they
        wait in the SoA tail, which the fix-up pass of K1 advances through the scalar body), so its tail stays short."""
        base = max(1, self.sort_interval)
        self._vrms = [0.0] * self.S
        if particles.x.shape[4] > 0:
            a = particles.active.reshape(self.S, -1)
            v2 = (particles.u.to(torch.float64) ** 2).sum(-1).reshape(self.S, -1)
            self._vrms = torch.sqrt((v2 * a).sum(1) / a.sum(1).clamp(min=1)).tolist()
        if self.sort_interval <= 0 or particles.x.shape[4] == 0:
            return [base] * self.S
        vrms = self._vrms
        vmax = max(vrms) if max(vrms) > 0 else 1.0
        return [int(min(20 * base, max(base, round(base * vmax / v)))) if v > 0 else 20 * base for v in vrms]

    def load_state(self, particles, fields3=None):
        """Replace the resident state from reference-layout pytrees (same shapes as at construction); re-sorts.
        The sort cadence and the drift estimate behind the K1 reduction choice keep the rms velocities seen at construction
        (cost knobs only: neither changes results)."""
        self._import(particles, 1.0)
        if fields3 is not None:
            for dst, src in zip((self.E, self.B, self.J), fields3):
                for d, s_ in zip(dst, src):
                    d.copy_(ops._chk(s_, "field", self.dtype))
            self._J_ghosts_stale = False          # the caller's J comes with its ghosts; a kept J may still owe its refresh
            self._refresh_imported_fields()
        self.sort()

    def _refresh_imported_fields(self):
        """The reference refreshes E's guard cells at the top of update_B and B's at the top of update_E (first_order_yee.py:115, 41);
        the resident step relies on guard cells that are valid from the previous step, so imported fields are refreshed once here."""
        fbc = tuple(self.p.field_bc)
        self.halo.refresh_(self.E, fbc)
        self.halo.refresh_(self.B, fbc)

    def sort(self, which=None):
        """K2: counting sort of the given species (default: all) by local cell; also compacts dead (absorbed / migrated)
        slots away."""
        L = _lib.lib()
        st = ops._stream()
        n = self.ncells + 1
        for s, sp_ in enumerate(self.species):
            if which is not None and s not in which:
                continue
            src, dst = self._soa(sp_), self._soa(sp_, 1 - sp_.cur)
            self._cell_count.zero_()
            check(L.pic_sort_histogram(ctypes.byref(self.p), ctypes.byref(src), ops._p(self._cell_count), st), "pic_sort_histogram")
            if self.k1_variant == "pair":   # blocked + padded: every supercell's slice starts on a multiple of 4 slots (K1 v10)
                check(L.pic_sort_blocked_offsets(ctypes.byref(self.p), ops._p(self._cell_count), ops._p(self._cell_offset),
                                                 ops._p(self._blk_work), ops._p(sp_.blk_off), ops._p(self._scan_scratch), sp_.cap,
                                                 ops._p(self.flags), st), "pic_sort_blocked_offsets")
            else:
                check(L.pic_sort_scan(n, ops._p(self._cell_count), ops._p(self._cell_offset), ops._p(self._scan_scratch), st), "pic_sort_scan")
            self._cell_count.zero_()
            check(L.pic_sort_scatter(ctypes.byref(self.p), ctypes.byref(src), ctypes.byref(dst), ops._p(self._cell_offset),
                                     ops._p(self._cell_count), st), "pic_sort_scatter")
            if self.k1_variant == "pair":   # (the cursor array now holds the per-cell counts again) padding slots -> dead
                check(L.pic_sort_blocked_finish(ctypes.byref(self.p), ops._p(self._cell_offset), ops._p(self._cell_count),
                                                ops._p(sp_.blk_off), ctypes.byref(dst), st), "pic_sort_blocked_finish")
            sp_.cur = 1 - sp_.cur           # (the scatter also set n_dev = number of slots in use, on the device)
            if hasattr(self, "_sorted_at"):
                self._sorted_at[s] = self.step_count
            if self.k1_variant == "tile":   # first slot of every 4x4x4-cell supercell of the blocked sort order (K1 v9)
                sp_.blk_off.copy_(self._cell_offset[::64])

    # ------------------------------------------------------------------------------------------ the step
    def step(self, n_steps=1):
        for _ in range(int(n_steps)):
            if self._use_graph():
                self._step_graphed()
            else:
                self._step_once()

    def _mark(self, name):
        """Phase boundary for the per-phase timeline (only when `phase_events` is a dict): the time since the previous mark is
        booked under `name`."""
        if self.phase_events is None:
            return
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        prev = self.phase_events.get("_last")
        if prev is not None and name is not None:
            self.phase_events.setdefault(name, []).append((prev, e))
        self.phase_events["_last"] = e

    def phase_summary(self, n_steps):
        """{phase: ms per step} from the recorded events (host sync)."""
        torch.cuda.synchronize()
        return {k: sum(a.elapsed_time(b) for a, b in v) / max(1, n_steps) for k, v in (self.phase_events or {}).items() if k != "_last"}

    def _use_graph(self):
        """Replay the step as ONE CUDA graph launch where launch latency is what a step costs (the packaged demos: 6 k - 260 k
        particles, ~25 launches of a few microseconds each).  Single GPU, fused Yee, no per-launch timing requested;
        PIC_GRAPH=0 disables, PIC_GRAPH=1 forces (any size)."""
        mode = os.environ.get("PIC_GRAPH", "auto")
        if (mode == "0" or getattr(self, "_graph_off", False) or self.distributed or self.k1_events is not None
                or self.phase_events is not None or not self._yee_fused):
            return False
        return mode == "1" or sum(sp_.cap for sp_ in self.species) <= (1 << 22)

    def _step_graphed(self):
        """One step through the graph cache.  A graph is valid for one set of buffer identities: the key is which sort buffer every
        species currently lives in, the parity of the E/B (and filtered-J) ping-pong and the K1 options; the sort itself runs
        outside the graph (it flips the buffers, and a new key is captured -- at most a handful of graphs exist)."""
        if self._graph_warm < 3:                       # lazily allocated scratch, function attributes, ... : first steps eagerly
            self._graph_warm += 1
            return self._step_once()
        key = (tuple(sp_.cur for sp_ in self.species), self.E[0].data_ptr(), self.B[0].data_ptr(), self.J[0].data_ptr(),
               tuple(self._k1_options(s) for s in range(self.S)), self.k1_variant)
        ent = self._graphs.get(key)
        if ent is None:
            g = torch.cuda.CUDAGraph()
            state = (self.E, self._E2, self.B, self._B2, self.J, self._J2, self._J_ghosts_stale)
            self._swapped = []
            try:
                with torch.cuda.graph(g):
                    self._step_core()
            except Exception:
                # something on this configuration's path cannot be captured: run eagerly from now on (nothing was executed)
                self.E, self._E2, self.B, self._B2, self.J, self._J2, self._J_ghosts_stale = state
                self._graph_off = True
                torch.cuda.synchronize()
                return self._step_once()
            swaps = tuple(self._swapped)
            # the capture performed the buffer swaps on the host but executed nothing: restore, replay applies them again
            self.E, self._E2, self.B, self._B2, self.J, self._J2, self._J_ghosts_stale = state
            ent = self._graphs[key] = (g, swaps)
        g, swaps = ent
        g.replay()
        for what in swaps:
            if what == "EB":
                self.E, self._E2 = self._E2, self.E
                self.B, self._B2 = self._B2, self.B
            elif what == "J":
                self.J, self._J2 = self._J2, self.J
        self._J_ghosts_stale = True
        self._after_fields()

    def _step_once(self):
        self._swapped = []
        self._step_core()
        self._mark("yee+EB_refresh")
        self._after_fields()
        self._mark("sort")

    def _step_core(self):
        L = _lib.lib()
        p, st = self.p, ops._stream()
        fbc = tuple(p.field_bc)
        pbc = tuple(p.particle_bc)
        self._mark(None)
        for c in self.J:
            c.zero_()
        if self.distributed:
            for sp_ in self.species:
                check(L.pic_packets_reset(ctypes.byref(p), ctypes.byref(sp_.leave), st), "pic_packets_reset")
        extE = ops._v(self.ext_E) if self.ext_E is not None else None
        extB = ops._v(self.ext_B) if self.ext_B is not None else None
        for s, sp_ in enumerate(self.species):
            soa = self._soa(sp_)
            if self.k1_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            leave = ctypes.byref(sp_.leave) if sp_.leave is not None else None
            rc = _lib.PIC_EUNSUPPORTED
            if self.k1_variant == "pair":
                rc = L.pic_fused_pair3d(ctypes.byref(p), s, ctypes.byref(soa), ops._p(sp_.blk_off), self.ncells // 64,
                                        self._k1_options(s), ops._v(self.E), ops._v(self.B), ops._v(self.J), leave, ops._p(self.flags),
                                        ops._p(self._pair_work), self._pair_work.numel(), st)
            elif self.k1_variant == "tile":
                rc = L.pic_fused_tile3d(ctypes.byref(p), s, ctypes.byref(soa), ops._p(sp_.blk_off), self.ncells // 64,
                                        self._k1_options(s), ops._v(self.E), ops._v(self.B), ops._v(self.J), leave, ops._p(self.flags), st)
            if self.k1_variant in ("pair", "tile"):
                if rc == _lib.PIC_EUNSUPPORTED:     # e.g. no TMA driver entry point: the global-gather K1 computes the same step
                    self.k1_variant = "global"      # (it reads the padded stream as it is: padding slots are dead slots)
                else:
                    check(rc, "pic_fused_" + self.k1_variant + "3d")
            if self.k1_variant not in ("pair", "tile"):
                check(L.pic_fused_push_deposit(ctypes.byref(p), s, self.deposition, ctypes.byref(soa), ops._v(self.E), ops._v(self.B),
                                               extE, extB, ops._v(self.J), leave, ops._p(self.flags), st), "pic_fused_push_deposit")
            if self.k1_events is not None:
                e1.record()
                self.k1_events.append((e0, e1))
        self._mark("zero_J+K1")
        self.halo.migrate(self)
        self._mark("migrate")
        # J: fold ghost deposits to their owners, then refresh (Esirkepov.py:357-359 / J_from_rhov.py:226-228)
        self._J_merged = (self.distributed and self._yee_fused and self.current_filter == "none" and hasattr(self.halo, "fold_refresh_")
                          and all((int(pbc[a]) == 0 and int(fbc[a]) == 0 and int(p.tile[a]) >= 2 * int(p.g)) if int(p.gmesh[a]) != int(p.mesh[a])
                                  else (int(fbc[a]) != 0 or int(p.tile[a]) >= int(p.g)) for a in range(3))
                          and os.environ.get("PIC_J_MERGED", "1") == "1")
        self._oneshot = self._J_merged and hasattr(self.halo, "fold_refresh_oneshot_") and os.environ.get("PIC_HALO_ONESHOT", "1") == "1"
        if self._oneshot:       # ... and all split axes in ONE round (faces, edges and corners sent explicitly)
            self.halo.fold_refresh_oneshot_(self.J, pbc)
        elif self._J_merged:    # multi-GPU, periodic split axes: fold and refresh of J in one exchange per axis
            self.halo.fold_refresh_(self.J, pbc)
        else:
            self.halo.fold_(self.J, pbc)
        if self.current_filter in ("bilinear", "digital"):                       # J_from_rhov.py:234-255
            self.halo.refresh_(self.J, pbc)
            if self._J2 is None:
                self._J2 = [torch.empty_like(c) for c in self.J]
            for c, o in zip(self.J, self._J2):
                ops.filter27(p, self.current_filter, self.alpha, c, out=o)
            self.J, self._J2 = self._J2, self.J
            self._swapped.append("J")
        # update_E reads J on the tile interior only, so the ghost refresh the reference performs here
        # (Esirkepov.py:359 / J_from_rhov.py:228,246) is deferred until somebody looks at J (export_state): one guard-cell
        # exchange less per step; the exported J is identical.
        self._J_ghosts_stale = True
        self._mark("J_fold_refresh")
        if self._yee_fused and self._step_fields_fused(fbc, pbc):
            return
        # B half step from E_old (evolve.py:88); E halos are valid from the previous step
        ops.update_B_(p, self.B, self.E)
        self.halo.refresh_(self.B, fbc)
        # E full step (evolve.py:92)
        ops.update_E_(p, self.E, self.B, self.J)
        self.halo.refresh_(self.E, fbc)
        if self.alpha != 1.0:
            self.E = [ops.filter27(p, "digital", self.alpha, c) for c in self.E]
        walls = False
        for axis in range(3):
            if fbc[axis] == BC_CONDUCTING:                                        # first_order_yee.py:80-89
                for c in range(3):
                    if c != axis:
                        ops.zero_wall_(p, self.E[c], axis)
                walls = True
        if walls or self.alpha != 1.0:
            self.halo.refresh_(self.E, fbc)
        # B half step from E_new (+ filter) (evolve.py:96)
        ops.update_B_(p, self.B, self.E)
        if self.alpha != 1.0:
            self.halo.refresh_(self.B, fbc)
            self.B = [ops.filter27(p, "digital", self.alpha, c) for c in self.B]
        self.halo.refresh_(self.B, fbc)

    def _step_fields_fused(self, fbc, pbc):
        """B(half) -> E -> B(half) in one pass over the fields (pic_yee_fused) into the spare E/B arrays, then swap.  Guard cells of
        single-rank periodic axes are written by the kernel; axes split across ranks exchange E and B together afterwards, and
        need J's guard planes refreshed beforehand (the refresh the reference performs after the fold, Esirkepov.py:359)."""
        p = self.p
        if self._E2 is None:
            self._E2 = [torch.zeros_like(c) for c in self.E]
            self._B2 = [torch.zeros_like(c) for c in self.B]
        # axes whose guard cells the kernel serves itself: single-rank axes that are non-periodic (walls) or periodic and at least
        # as wide as the guard depth (wrapped indices).  The others -- split across ranks, or reduced (one cell wide) -- are read
        # from the guard cells, so J's must be refreshed there first and E's, B's afterwards.
        done = tuple(a for a in range(3) if int(p.gmesh[a]) == int(p.mesh[a]) and (int(fbc[a]) != 0 or int(p.tile[a]) >= int(p.g)))
        if len(done) < 3 and not getattr(self, "_J_merged", False):
            # (with the FIELD boundary conditions: what the kernel needs in J's guard plane is the J of the cell the field
            # stencil continues into; the guard cells the reference leaves in J -- particle BCs -- are restored on export)
            self.halo.refresh_(self.J, fbc, skip_axes=done)
            self._J_ghosts_stale = True
        rc = _lib.lib().pic_yee_fused(ctypes.byref(p), ops._v(self.E), ops._v(self.B), ops._v(self.J), ops._v(self._E2), ops._v(self._B2),
                                      ops._stream())
        if rc == _lib.PIC_EUNSUPPORTED:
            self._yee_fused = False
            return False
        check(rc, "pic_yee_fused")
        self.E, self._E2 = self._E2, self.E
        self.B, self._B2 = self._B2, self.B
        self._swapped.append("EB")
        if len(done) < 3:
            if getattr(self, "_oneshot", False):
                self.halo.refresh_oneshot_(self.E + self.B, fbc, skip_axes=done)
            else:
                self.halo.refresh_(self.E + self.B, fbc, skip_axes=done)
        return True

    def _after_fields(self):
        self.step_count += 1
        if self.sort_interval > 0:
            due = [s for s in range(self.S) if self.step_count % self.sort_every[s] == 0]
            if due:
                self.sort(due)

    # ------------------------------------------------------------------------------------------ results
    def overflow(self):
        """Host sync: the reference's overflow flag (invalid > 1 tile jump or a capacity overflow), OR-ed over steps."""
        f = int(self.flags[0].item())
        if f & 8:
            raise PicError("pic_fused_pair3d was handed supercell slices that do not come from the blocked, padded sort")
        return self.overflow_previous or f != 0

    def n_particles(self):
        """Live particles on this GPU (host sync): the slots in use minus the dead ones (x = NaN: holes left by migrated or
        absorbed particles, and the padding slots of the blocked sort)."""
        total = 0
        for sp_ in self.species:
            n = min(int(sp_.n_dev.item()), sp_.cap)
            total += n - int(torch.isnan(sp_.buf[sp_.cur][0][:n]).sum().item())
        return total

    def export_state(self, cap_ref=None, out=None, fields=True):
        """Reference pytrees: (TiledParticles, fields 8-tuple).  Slots are restored by id on a single GPU.
        `out=(x, u, active)` reuses caller-owned tensors of the reference shape; `fields=False` skips the field copies
        (returns (particles, None))."""
        L = _lib.lib()
        st = ops._stream()
        cap = self.cap_ref if cap_ref is None else int(cap_ref)
        if out is not None:
            x, u, a = out
            if int(x.shape[4]) < cap:
                raise PicError("export_state(out=...): capacity too small")
            cap = int(x.shape[4])
            x.zero_(); u.zero_(); a.zero_()
        else:
            x = torch.zeros((1, 1, 1, self.S, cap, 3), dtype=self.dtype, device=self.device)
            u = torch.zeros_like(x)
            a = torch.zeros((1, 1, 1, self.S, cap), dtype=torch.bool, device=self.device)
        self._counter.zero_()
        for s, sp_ in enumerate(self.species):
            soa = self._soa(sp_)
            if self.distributed:
                soa.id = None
            check(L.pic_soa_export(ctypes.byref(self.p), s, ctypes.byref(soa), ops._p(x), ops._p(u), ops._p(a), cap,
                                   ops._p(self._counter[s:s + 1]), st), "pic_soa_export")
        if self.distributed or not self.track_ids:
            live = self._counter[:self.S].cpu().tolist()
            if any(int(n) > cap for n in live):          # fixed-capacity contract of the reference layout (:169-230)
                self.flags[0:1] |= 2
        if not fields:
            return TiledParticles(x=x, u=u, active=a), None
        if self._J_ghosts_stale:
            self.halo.refresh_(self.J, tuple(self.p.particle_bc))
            self._J_ghosts_stale = False
        rho, phi, ext = self._passthrough
        overflow = torch.tensor(self.overflow(), device=self.device)
        fields = (tuple(c.clone() for c in self.E), tuple(c.clone() for c in self.B), tuple(c.clone() for c in self.J), rho, phi, ext,
                  None, overflow)
        return TiledParticles(x=x, u=u, active=a), fields

    # ------------------------------------------------------------------------------------------ conservation diagnostics
    def _diag(self):
        """(params, halo, cast) of the conservation diagnostics: the run's own for float64; for a float32 run the float64 twins,
        so that what is measured is the state of the float32 simulation, not the round-off of a float32 rho deposit from
        float32 positions (which alone is 7e-4 of max |d rho / dt| at 256^3)."""
        if self.dtype == torch.float64:
            return self.p, self.halo, (lambda t: t)
        if self._halo64 is None:
            self._halo64 = type(self.halo)(self.p64, self.halo.group, self.halo.device) if self.distributed else LocalHalo(self.p64)
        return self.p64, self._halo64, (lambda t: t.to(torch.float64))

    def charge_density(self):
        """rho of the resident particles on this rank's ghosted tile: node deposit of every species (deposition/rho.py:66-150),
        ghost deposits folded to their owners across ranks (particle BCs), ghosts left zero.  Diagnostics only: goes through
        the reference-layout export; float64 arithmetic whatever the dtype of the run."""
        p, halo, cast = self._diag()
        parts, _ = self.export_state(fields=False)
        x = cast(parts.x)
        rho = ops.deposit(p, "rho", x, x, parts.active, cast(self.E[0]))
        del x, parts
        halo.fold_([rho], tuple(self.p.particle_bc))
        return rho

    def gauss_residual(self, rho=None):
        """div E - rho/eps on the tile interior; E's guard cells are valid after every step."""
        p, halo, cast = self._diag()
        rho = self.charge_density() if rho is None else rho
        return ops.div_residual(p, [cast(c) for c in self.E], rho, -1.0 / float(self.p.eps))

    def conservation_step(self):
        """Advance ONE step and measure what Esirkepov's deposition promises (esirkepov_test.py:700-744 at any size, any number of
        GPUs): max |(rho_new - rho_old)/dt + div J| against max |(rho_new - rho_old)/dt|, and the drift of the Gauss residual
        div E - rho/eps over the step against max |rho|/eps.  Returns rank-local maxima (reduce with MAX over ranks).  The
        residuals are evaluated in float64 from the state of the run (see _diag)."""
        p, halo, cast = self._diag()
        rho0 = self.charge_density()
        g0 = self.gauss_residual(rho0)
        self.step(1)
        rho1 = self.charge_density()
        if self._J_ghosts_stale:                    # div J at the first interior plane reads the lower ghost plane
            self.halo.refresh_(self.J, tuple(self.p.particle_bc))
            self._J_ghosts_stale = False
        inv_dt = 1.0 / float(self.p.dt)
        cont = ops.div_residual(p, [cast(c) for c in self.J], rho1, inv_dt, rho0, -inv_dt)
        g1 = self.gauss_residual(rho1)
        g = int(self.p.g)
        I = (0, 0, 0, slice(g, -g), slice(g, -g), slice(g, -g))
        return {"continuity_residual_max": float(cont[I].abs().max()),
                "continuity_scale": float(((rho1[I] - rho0[I]) * inv_dt).abs().max()),
                "gauss_drift_max": float((g1[I] - g0[I]).abs().max()),
                "gauss_scale": float(rho1[I].abs().max()) / float(self.p.eps),
                "evaluated_in": "float64"}


def time_loop_electrodynamic_resident(particles, species_config, fields, static_parameters, dynamic_parameters, n_steps=1, **kw):
    """Convenience: import -> n_steps fused steps -> export, with the reference's (particles, fields) signature."""
    sim = Simulation(particles, species_config, fields, static_parameters, dynamic_parameters, **kw)
    sim.step(n_steps)
    return sim.export_state()

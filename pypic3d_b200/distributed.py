"""Multi-GPU domain decomposition: one process per GPU, one reference tile per rank (the reference's one-tile-per-device
mesh, ghost_cells.py:98-116 / particle_tile_communication.py:292-426), NCCL over NVLink via `torch.distributed` P2P.

Two exchange steps per time step, nearest neighbours only:
  * guard cells: `refresh_` / `fold_` walk the axes x -> y -> z exactly like the reference (`_local_refresh_scalar_tile`,
    `_local_fold_scalar_tile`), sending FULL-transverse-extent faces so edges and corners propagate through the sequence;
    on an axis that is not split across ranks the single-GPU kernel runs instead (periodic self-exchange / walls);
  * particles: the fused kernel writes leavers into 27 per-direction packet buffers; `migrate` exchanges the counts,
    then the packets (the reference's 26 dense streams `_send_particle_stream` become <= 26 compact messages), and appends.
Rank layout: rank = (cx * my + cy) * mz + cz.

`kernels` abstracts pack/unpack/local passes so the protocol can be exercised on CPU tensors under the gloo backend in
the test-suite; the product always uses `CudaHaloKernels` (hand-written CUDA behind the C ABI) -- there is no CPU product path.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib, ops

HALO_SET, HALO_ADD, HALO_SUB = 0, 1, 2


class CudaHaloKernels:
    def pack(self, p, axis, start, nplanes, fields, buf):
        ops.pack_planes(p, axis, start, nplanes, fields, buf)

    def unpack(self, p, axis, start, nplanes, fields, buf, mode):
        ops.unpack_planes_(p, axis, start, nplanes, fields, buf, mode)

    def refresh_axis(self, p, axis, bc, fields):
        ops.halo_refresh_axis_(p, axis, bc, fields)

    def fold_axis(self, p, axis, bc, fields):
        ops.halo_fold_axis_(p, axis, bc, fields)

    def pack_boxes(self, p, boxes, fields, buf):
        ops.pack_boxes(p, boxes, fields, buf)

    def unpack_boxes(self, p, boxes, fields, buf, mode):
        ops.unpack_boxes_(p, boxes, fields, buf, mode)


def rank_of(coords, mesh):
    return (coords[0] * mesh[1] + coords[1]) * mesh[2] + coords[2]


def coords_of(rank, mesh):
    return (rank // (mesh[1] * mesh[2]), (rank // mesh[2]) % mesh[1], rank % mesh[2])


def neighbor(coords, mesh, axis, step, periodic):
    """Rank of the neighbour `step` (+1/-1) along `axis`, or None at an open chain end (ghost_cells.py:72-83)."""
    c = list(coords)
    c[axis] += step
    if c[axis] < 0 or c[axis] >= mesh[axis]:
        if not periodic:
            return None
        c[axis] %= mesh[axis]
    return rank_of(c, mesh)


def gather_tiles(tile, gmesh, group=None):
    """Diagnostics boundary of a multi-GPU run: every rank contributes its (1,1,1,Lx,Ly,Lz) field tile and receives the tile-major
    array (ntx,nty,ntz,Lx,Ly,Lz) of the whole job -- what `diagnostics.assemble_tiled_scalar_field` and the reference's
    `jax.device_get` of a sharded field (output_adapters.py:48) start from.  Ranks are laid out like `rank_of`."""
    gmesh = tuple(int(v) for v in gmesh)
    n = gmesh[0] * gmesh[1] * gmesh[2]
    t = tile.reshape((1,) + tuple(tile.shape[3:])).contiguous()
    if n == 1:
        return tile.reshape((1, 1, 1) + tuple(tile.shape[3:])).clone()
    out = [torch.empty_like(t) for _ in range(n)]
    dist.all_gather(out, t, group=group)
    return torch.cat(out, 0).reshape(gmesh + tuple(tile.shape[3:]))


DIRS = [(1 - sx, 1 - sy, 1 - sz) for sx in range(3) for sy in range(3) for sz in range(3)]   # dir code -> offset (ox,oy,oz)


def grouped_packet_layout(dirs, caps, key):
    """Row layout of the shared packet buffer that makes everything exchanged with one peer ONE contiguous message.
    dirs: [(direction code, destination rank or None, source rank or None)] (DistributedHalo.active_dirs); caps[s][dcode]: packet
    capacity in rows per species (0: unused direction); key: 1 = group by destination rank (the leave buffer K1 writes), 2 = by
    source rank (the receive buffer the append reads).  Within a peer: species, then direction code ascending -- the sender's
    group for (me -> peer) and the receiver's group for (peer <- me) hold the same directions in the same order, so a message can
    be copied verbatim.  Returns (row_off[s][27], rows, [(peer, first element, one-past-last element)]); a packet is
    cap + 1 rows (header row with the count) of 7 reals."""
    S = len(caps)
    row_off = [[0] * 27 for _ in range(S)]
    peers = sorted({t[key] for t in dirs if t[key] is not None})
    rows, slices = 0, []
    for peer in peers + [None]:
        lo = rows
        for s_ in range(S):
            for dc, dst, src in dirs:
                if (dst, src)[key - 1] != peer or caps[s_][dc] == 0:
                    continue
                row_off[s_][dc] = rows
                rows += caps[s_][dc] + 1
        if peer is not None and rows > lo:
            slices.append((peer, lo * 7, rows * 7))
    return row_off, rows, slices


class DistributedHalo:
    def __init__(self, params, group=None, device=None, kernels=None):
        self.p = params
        self.group = group
        self.device = device
        self.k = kernels if kernels is not None else CudaHaloKernels()
        self.mesh = tuple(int(v) for v in params.gmesh)
        self.coords = tuple(int(v) for v in params.moff)
        self.rank = rank_of(self.coords, self.mesh)
        self.L = tuple(int(params.tile[a]) + 2 * int(params.g) for a in range(3))
        self.g = int(params.g)
        self._bufs = {}

    # ------------------------------------------------------------------ helpers
    def _is_split(self, axis):
        return self.mesh[axis] > 1

    def _buf(self, key, n, like):
        b = self._bufs.get((key, like.dtype))
        if b is None or b.numel() < n:
            b = torch.empty(n, dtype=like.dtype, device=like.device)
            self._bufs[(key, like.dtype)] = b
        return b[:n]

    def _plane_elems(self, axis, ncomp):
        u, v = [a for a in range(3) if a != axis]
        return ncomp * self.g * self.L[u] * self.L[v]

    def _exchange(self, sends, recvs):
        """sends/recvs: lists of (peer_rank, tensor) in matching order on both sides."""
        reqs = []
        opl = []
        for peer, t in sends:
            opl.append(dist.P2POp(dist.isend, t, peer, group=self.group))
        for peer, t in recvs:
            opl.append(dist.P2POp(dist.irecv, t, peer, group=self.group))
        if opl:
            reqs = dist.batch_isend_irecv(opl)
            for r in reqs:
                r.wait()

    def allreduce_max(self, values):
        """Element-wise max over ranks of a short list of ints (set-up time only)."""
        t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return [int(v) for v in t.tolist()]

    # ------------------------------------------------------------------ guard cells
    def refresh_(self, fields, bcs, skip_axes=()):
        """In-place refresh x -> y -> z (ghost_cells.py:181-215).  `skip_axes`: axes whose guard cells the caller has already
        filled (pic_yee_fused writes the guard copies of single-rank periodic axes itself); the order and the full transverse
        extent of the remaining exchanges are unchanged, so edges and corners still propagate."""
        g, p = self.g, self.p
        for axis in range(3):
            bc = int(bcs[axis])
            if axis in skip_axes:
                continue
            if not self._is_split(axis):
                self.k.refresh_axis(p, axis, bc, fields)
                continue
            n = self._plane_elems(axis, len(fields))
            L = self.L[axis]
            up = neighbor(self.coords, self.mesh, axis, +1, bc == 0)
            dn = neighbor(self.coords, self.mesh, axis, -1, bc == 0)
            s_up, s_dn = self._buf(("s_up", axis), n, fields[0]), self._buf(("s_dn", axis), n, fields[0])
            r_lo, r_hi = self._buf(("r_lo", axis), n, fields[0]), self._buf(("r_hi", axis), n, fields[0])
            self.k.pack(p, axis, L - 2 * g, g, fields, s_up)     # upper interior -> +1 neighbour's lower ghost (:187)
            self.k.pack(p, axis, g, g, fields, s_dn)             # lower interior -> -1 neighbour's upper ghost (:191)
            sends, recvs = [], []
            if up is not None:
                sends.append((up, s_up))
            if dn is not None:
                sends.append((dn, s_dn))
            # receive order must mirror the peers' send order: first what their "+1 send" delivers (from my -1 side)
            if dn is not None:
                recvs.append((dn, r_lo))
            else:
                r_lo.zero_()                                     # open chain end: zeros (:189-191)
            if up is not None:
                recvs.append((up, r_hi))
            else:
                r_hi.zero_()
            self._exchange(sends, recvs)
            self.k.unpack(p, axis, 0, g, fields, r_lo, HALO_SET)
            self.k.unpack(p, axis, L - g, g, fields, r_hi, HALO_SET)

    def fold_(self, fields, bcs):
        """In-place fold-add of ghost deposits to their owners, then zero ghosts (ghost_cells.py:263-316)."""
        g, p = self.g, self.p
        for axis in range(3):
            bc = int(bcs[axis])
            if not self._is_split(axis):
                self.k.fold_axis(p, axis, bc, fields)
                continue
            n = self._plane_elems(axis, len(fields))
            L = self.L[axis]
            up = neighbor(self.coords, self.mesh, axis, +1, bc == 0)
            dn = neighbor(self.coords, self.mesh, axis, -1, bc == 0)
            s_lo, s_hi = self._buf(("s_up", axis), n, fields[0]), self._buf(("s_dn", axis), n, fields[0])
            r_from_up, r_from_dn = self._buf(("r_lo", axis), n, fields[0]), self._buf(("r_hi", axis), n, fields[0])
            self.k.pack(p, axis, 0, g, fields, s_lo)             # lower ghost belongs to the -1 neighbour's upper interior (:271)
            self.k.pack(p, axis, L - g, g, fields, s_hi)         # upper ghost belongs to the +1 neighbour's lower interior (:272)
            sends, recvs = [], []
            if up is not None:
                sends.append((up, s_hi))
            if dn is not None:
                sends.append((dn, s_lo))
            if dn is not None:
                recvs.append((dn, r_from_dn))                    # their upper ghost -> my lower interior
            if up is not None:
                recvs.append((up, r_from_up))                    # their lower ghost -> my upper interior
            self._exchange(sends, recvs)
            if up is not None:
                self.k.unpack(p, axis, L - 2 * g, g, fields, r_from_up, HALO_ADD)
            if dn is not None:
                self.k.unpack(p, axis, g, g, fields, r_from_dn, HALO_ADD)
            if bc == 1:                                          # conducting wall ranks (:238-260)
                if self.coords[axis] == 0:
                    self.k.unpack(p, axis, g, g, fields, s_lo, HALO_SUB)
                if self.coords[axis] == self.mesh[axis] - 1:
                    self.k.unpack(p, axis, L - 2 * g, g, fields, s_hi, HALO_SUB)
            z = self._buf(("zero", axis), n, fields[0])
            z.zero_()
            self.k.unpack(p, axis, 0, g, fields, z, HALO_SET)
            self.k.unpack(p, axis, L - g, g, fields, z, HALO_SET)

    def fold_refresh_(self, fields, bcs):
        """fold_ followed by refresh_ along the split axes in ONE exchange per axis (periodic axes only; the caller checks):
        both sides send their 2g boundary planes (g ghost + g interior) and add what they receive plane by plane, so the interior
        planes receive the neighbour's ghost deposits (the fold) and the ghost planes end up with the neighbour's totals (the
        refresh) -- every node, owned or ghost, holds the sum of all ranks' deposits, the same values as fold_ + refresh_ up to
        the order of two additions.  Valid for split axes at least 2g cells wide (the two 2g slabs of a tile must not overlap:
        a ghost plane then has exactly one owner-side partner); Simulation checks and falls back to fold_ + refresh_.  Axes that are not split are folded locally (their ghosts are left zero: the fused Yee kernel
        wraps those indices itself).  x -> y -> z with the full transverse extent, so edges and corners propagate as in
        ghost_cells.py:199-215 / :263-316."""
        g, p = self.g, self.p
        for axis in range(3):
            bc = int(bcs[axis])
            if not self._is_split(axis):
                self.k.fold_axis(p, axis, bc, fields)
                continue
            assert bc == 0, "fold_refresh_ handles periodic split axes"
            n = 2 * self._plane_elems(axis, len(fields))
            L = self.L[axis]
            up = neighbor(self.coords, self.mesh, axis, +1, True)
            dn = neighbor(self.coords, self.mesh, axis, -1, True)
            s_hi, s_lo = self._buf(("fr_s_hi", axis), n, fields[0]), self._buf(("fr_s_lo", axis), n, fields[0])
            r_dn, r_up = self._buf(("fr_r_dn", axis), n, fields[0]), self._buf(("fr_r_up", axis), n, fields[0])
            self.k.pack(p, axis, L - 2 * g, 2 * g, fields, s_hi)   # my upper interior + upper ghost -> +1 neighbour's lower ghost + lower interior
            self.k.pack(p, axis, 0, 2 * g, fields, s_lo)           # my lower ghost + lower interior -> -1 neighbour's upper interior + upper ghost
            self._exchange([(up, s_hi), (dn, s_lo)], [(dn, r_dn), (up, r_up)])
            self.k.unpack(p, axis, 0, 2 * g, fields, r_dn, HALO_ADD)
            self.k.unpack(p, axis, L - 2 * g, 2 * g, fields, r_up, HALO_ADD)

    # ------------------------------------------------------------------ one-shot exchange (all split axes in ONE round)
    def _box_plan(self, ncomp, kind):
        """Boxes this rank sends / receives when every split (periodic) axis is served in one round, grouped by peer.
        kind "refresh": direction o -> my interior slab next to the (+o) face goes to the (+o) neighbour's ghost slab on its (-o)
        side; a split axis with o = 0 spans its INTERIOR (its guard cells are other boxes), an axis that is not split its FULL
        extent (its guard cells are valid already, or refreshed locally afterwards).  kind "sum": my 2g-thick slab (interior + ghost) next to the (+o) face goes to the (+o) neighbour's
        2g-thick slab on its (-o) side and is ADDED there; axes with o = 0 span the FULL extent -- after the round every node, owned
        or ghost, holds the sum of all ranks' deposits (fold + refresh of J at once).  Returns (send boxes, recv boxes,
        send slices, recv slices): boxes in message order, slices = [(peer, lo, hi)] element ranges of the packed buffers."""
        key = (ncomp, kind)
        plan = getattr(self, "_plans", {}).get(key)
        if plan is not None:
            return plan
        g, L = self.g, self.L
        split = [self._is_split(a) for a in range(3)]
        dirs = [o for o in DIRS if o != (0, 0, 0) and all(o[a] == 0 or split[a] for a in range(3))]

        def box(o, side):
            # side "send": the slab next to my (+o) face; side "recv": the slab next to my (-o) face (what the (-o) neighbour sent)
            lo, sz = [], []
            for a in range(3):
                s_ = o[a] if side == "send" else -o[a]
                if o[a] == 0:
                    # an axis that is not split travels with its guard cells (valid already when the Yee kernel wrote them, refreshed
                    # locally afterwards otherwise); a split axis contributes its interior: its guard cells are other boxes
                    r = (0, L[a]) if (kind == "sum" or not split[a]) else (g, L[a] - g)
                elif kind == "refresh":
                    r = ((L[a] - 2 * g, L[a] - g) if s_ > 0 else (g, 2 * g)) if side == "send" else ((L[a] - g, L[a]) if s_ > 0 else (0, g))
                else:
                    r = (L[a] - 2 * g, L[a]) if s_ > 0 else (0, 2 * g)
                lo.append(r[0]); sz.append(r[1] - r[0])
            return (tuple(lo), tuple(sz))

        def peer_of(o, sign):
            c = [(self.coords[a] + sign * o[a]) % self.mesh[a] for a in range(3)]
            return rank_of(c, self.mesh)

        def grouped(side):
            # sender: group by destination N(+o); receiver: group by source N(-o); inside a group the direction order is the same
            items = sorted(((peer_of(o, +1 if side == "send" else -1), DIRS.index(o), o) for o in dirs))
            boxes, slices, off = [], [], 0
            cur, lo_ = None, 0
            for peer, _, o in items:
                if peer != cur:
                    if cur is not None:
                        slices.append((cur, lo_, off))
                    cur, lo_ = peer, off
                b = box(o, side)
                boxes.append(b)
                off += ncomp * b[1][0] * b[1][1] * b[1][2]
            if cur is not None:
                slices.append((cur, lo_, off))
            return boxes, slices, off
        sb, ss, sn = grouped("send")
        rb, rs, rn = grouped("recv")
        plan = (sb, rb, ss, rs, sn, rn)
        if not hasattr(self, "_plans"):
            self._plans = {}
        self._plans[key] = plan
        return plan

    def exchange_boxes_(self, fields, kind):
        """One round for all split axes: `kind` "refresh" (E, B guard cells) or "sum" (J: fold and refresh at once).  Periodic
        split axes only (the caller checks); axes that are not split are left to the caller's local pass."""
        sb, rb, ss, rs, sn, rn = self._box_plan(len(fields), kind)
        sbuf = self._buf(("box_s", kind), sn, fields[0])
        rbuf = self._buf(("box_r", kind), rn, fields[0])
        self.k.pack_boxes(self.p, sb, fields, sbuf)
        self._exchange([(peer, sbuf[lo:hi]) for peer, lo, hi in ss], [(peer, rbuf[lo:hi]) for peer, lo, hi in rs])
        self.k.unpack_boxes(self.p, rb, fields, rbuf, HALO_SET if kind == "refresh" else HALO_ADD)

    def refresh_oneshot_(self, fields, bcs, skip_axes=()):
        """refresh_ with ONE exchange round: faces, edges and corners of all split axes travel explicitly, then the axes that are
        not split (and not in `skip_axes`) are refreshed locally over the full extent."""
        self.exchange_boxes_(fields, "refresh")
        for axis in range(3):
            if not self._is_split(axis) and axis not in skip_axes:
                self.k.refresh_axis(self.p, axis, int(bcs[axis]), fields)

    def fold_refresh_oneshot_(self, fields, bcs):
        """fold_refresh_ with ONE exchange round: every rank adds its neighbours' overlapping 2g-thick slabs, then the axes that are
        not split are folded locally."""
        self.exchange_boxes_(fields, "sum")
        for axis in range(3):
            if not self._is_split(axis):
                self.k.fold_axis(self.p, axis, int(bcs[axis]), fields)

    # ------------------------------------------------------------------ particles
    def active_dirs(self, particle_bcs):
        """Direction codes that can carry leavers: offsets only along split axes, existing neighbour on every offset axis."""
        out = []
        for d, off in enumerate(DIRS):
            if off == (0, 0, 0) or any(off[a] != 0 and not self._is_split(a) for a in range(3)):
                continue
            dst, src = list(self.coords), list(self.coords)
            ok = True
            for a in range(3):
                if off[a] == 0:
                    continue
                per = int(particle_bcs[a]) == 0
                dst[a] += off[a]; src[a] -= off[a]
                if per:
                    dst[a] %= self.mesh[a]; src[a] %= self.mesh[a]
            dst_ok = all(0 <= dst[a] < self.mesh[a] for a in range(3))
            src_ok = all(0 <= src[a] < self.mesh[a] for a in range(3))
            out.append((d, rank_of(dst, self.mesh) if dst_ok else None, rank_of(src, self.mesh) if src_ok else None))
        return out

    def exchange_packets(self, leave_bufs, recv_bufs, leave, particle_bcs):
        """One batched round: for every species and every active direction send the fixed-size packet (header + cap rows,
        layout `PicLeave`) to the destination rank and receive the matching packet from the source rank.
        leave_bufs / recv_bufs: per-species flat tensors; leave: per-species objects with row_off[27] / cap[27]."""
        dirs = self.active_dirs(particle_bcs)
        sends, recvs = [], []
        for s in range(len(leave_bufs)):
            for d, dst, src in dirs:
                cap = int(leave[s].cap[d])
                if cap == 0:
                    continue
                lo, hi = int(leave[s].row_off[d]) * 7, (int(leave[s].row_off[d]) + cap + 1) * 7
                if dst is not None:
                    sends.append((dst, leave_bufs[s][lo:hi]))
                if src is not None:
                    recvs.append((src, recv_bufs[s][lo:hi]))
        self._exchange(sends, recvs)

    def exchange_grouped(self, send_buf, recv_buf, send_slices, recv_slices):
        """One batched round with ONE message per peer: send_slices / recv_slices are [(peer rank, lo, hi)] element ranges of the
        shared packet buffers whose layout groups the packets by peer (Simulation._alloc_packets)."""
        self._exchange([(peer, send_buf[lo:hi]) for peer, lo, hi in send_slices], [(peer, recv_buf[lo:hi]) for peer, lo, hi in recv_slices])

    def migrate(self, sim):
        """Exchange the leaver packets written by K1 and append the arrivals to the resident SoA (K3/K4 of SURVEY.md
        section 7).  No host synchronisation: packet sizes are fixed and the row counts travel in the packet headers."""
        L = _lib.lib()
        st = ops._stream()
        self.exchange_grouped(sim._leave_all, sim._recv_all, sim._send_slices, sim._recv_slices)
        for sp_ in sim.species:
            soa = sim._soa(sp_)
            _lib.check(L.pic_soa_append_packets(ctypes.byref(self.p), ctypes.byref(soa), ctypes.byref(sp_.recv), ops._p(sim.flags), st),
                       "pic_soa_append_packets")

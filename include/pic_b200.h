/*
 * pic_b200.h -- C ABI of libpic_b200.so: hand-written sm_100a CUDA kernels for the PyPIC3D
 * electrodynamic PIC step (push + deposit + Yee + guard cells).
 *
 * The reference (uwplasma/PyPIC3D v0.1.3) is pure Python/JAX and has NO FFI; its boundary is the Python
 * calling convention of PyPIC3D/evolve.py:16-22 and of the sub-entry points listed next to each function
 * below.  Every entry point here is what a binding for that call site would bind: plain device pointers,
 * sizes and one POD parameter block -- no torch / JAX types.  All functions
 *   - run asynchronously on the CUDA stream passed as `stream` (a cudaStream_t cast to void*),
 *   - return 0 on success, a negative PIC_E* code on a bad argument, or a positive cudaError_t,
 *   - never throw, never allocate device memory, never synchronise (except where stated).
 * Threading: one host thread per GPU; calls on one stream are ordered.
 *
 * Layouts (identical to the reference so import/export are plain copies):
 *   tiled scalar field : (ntx,nty,ntz, Lx,Ly,Lz), L = W + 2g, C order, z fastest   (ghost_cells.py:13)
 *   tiled vector field : three such arrays (x,y,z components)                       (evolve.py:27)
 *   TiledParticles     : x,u (ntx,nty,ntz,S,cap,3) reals, active (ntx,nty,ntz,S,cap) bytes
 *                        (particles/particle_class.py:17-31); u is the velocity v, not gamma*v.
 *   SpeciesConfig      : charge, mass, weight (S,) doubles; update_x, update_u (S,3) bytes
 *                        (particles/particle_class.py:5-14)
 * `dtype` in PicParams selects float (0) or double (1) for every `void*` real array.
 */
#ifndef PIC_B200_H
#define PIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIC_F32 0
#define PIC_F64 1

#define PIC_PUSHER_BORIS 0        /* pusher/boris.py:15  boris_single_particle              */
#define PIC_PUSHER_BORIS_REL 1    /* pusher/boris.py:61  relativistic_boris_single_particle */
#define PIC_PUSHER_HC 2           /* pusher/higuera_cary.py:58                              */

#define PIC_BC_PERIODIC 0         /* grid_and_stencil.py:10 ; particles: periodic wrap      */
#define PIC_BC_CONDUCTING 1       /* grid_and_stencil.py:11 ; particles: reflecting         */
#define PIC_BC_ABSORBING 2        /* particles only (particle_tile_communication.py:45)     */

#define PIC_FILTER_DIGITAL 0      /* utilities/filters.py:98  */
#define PIC_FILTER_BILINEAR 1     /* utilities/filters.py:73  */

#define PIC_HALO_SET 0
#define PIC_HALO_ADD 1
#define PIC_HALO_SUB 2

#define PIC_EINVAL (-1)
#define PIC_EUNSUPPORTED (-2)

#define PIC_MAX_SPECIES 16

/* POD mirror of StaticParameters + the scalar leaves of DynamicParameters (parameters.py:14-53).
 * `mesh` is the tile-array shape resident on THIS GPU, `gmesh`/`moff` place it in the global tile mesh
 * (one process per GPU: mesh = (1,1,1), gmesh = process grid, moff = this rank's coordinates). */
typedef struct PicParams {
    int32_t dtype;            /* PIC_F32 / PIC_F64                                         */
    int32_t shape_factor;     /* 1 = CIC, 2 = TSC (deposition/shapes.py)                    */
    int32_t pusher;           /* PIC_PUSHER_*                                               */
    int32_t g;                /* guard cells                                                */
    int32_t mesh[3];          /* local tile-array shape                                     */
    int32_t gmesh[3];         /* global tile mesh shape                                     */
    int32_t moff[3];          /* offset of the local tile array inside the global mesh      */
    int32_t tile[3];          /* tile widths W (cells)                                      */
    int32_t field_bc[3];      /* StaticParameters.boundary_conditions                       */
    int32_t particle_bc[3];   /* StaticParameters.particle_boundary_conditions              */
    int32_t n_species;
    int32_t pad0;
    double dt, dx, dy, dz;
    double wind[3];           /* x_wind, y_wind, z_wind                                     */
    double C, eps, mu, alpha;
    double center0[3];        /* grids.center[a][0]  (= -wind/2 - d)                        */
    double vertex0[3];        /* grids.vertex[a][0]  (= -wind/2 - d/2)                      */
    double charge[PIC_MAX_SPECIES], mass[PIC_MAX_SPECIES], weight[PIC_MAX_SPECIES];
    uint8_t update_x[PIC_MAX_SPECIES][3], update_u[PIC_MAX_SPECIES][3];
} PicParams;

const char* pic_version(void);
int pic_params_size(void);       /* sizeof(PicParams) -- lets a binding verify its struct layout */

/* ---------------- reference-layout operators (drop-in for the reference's sub-entry points) ------------ */

/* pusher/particle_push.py:13 particle_push: gather E,B (6x up-to-27-point) + Boris/HC; writes u_out.
 * E/B: 3 tiled component arrays each (already E+E_ext, B+B_ext: utils.py:205). */
int pic_push(const PicParams* p, const void* x, const void* u_in, void* u_out, const uint8_t* active,
             int64_t cap, const void* const E[3], const void* const B[3], void* stream);

/* deposition/Esirkepov.py:105-331 deposit_one_tile for every local tile: J must be zeroed by the caller;
 * ghost deposits are left in the ghosts (fold with pic_halo_fold_axis). */
int pic_deposit_esirkepov(const PicParams* p, const void* x, const void* u, const uint8_t* active, int64_t cap,
                          void* const J[3], void* stream);
/* deposition/J_from_rhov.py:83-200 deposit_one_tile (node/face rho*v weights). */
int pic_deposit_direct(const PicParams* p, const void* x, const void* u, const uint8_t* active, int64_t cap,
                       void* const J[3], void* stream);
/* deposition/rho.py:66-150 deposit_one_tile. */
int pic_deposit_rho(const PicParams* p, const void* x, const uint8_t* active, int64_t cap, void* rho, void* stream);

/* particles/particle_tile_communication.py:82 update_tiled_particle_positions: x_out = x + active*upd*u*dt. */
int pic_move(const PicParams* p, const void* x_in, void* x_out, const void* u, const uint8_t* active, int64_t cap,
             double dt, void* stream);

/* particles/particle_tile_communication.py:440 refresh_tiled_particle_tiles over the LOCAL tile mesh:
 * global particle BCs, re-own by tile, move to the <=26 neighbour tiles (k-th incoming -> k-th free slot),
 * overflow flag (int32 on device, OR-ed in).  Out arrays must not alias the inputs.
 * scratch: int32[ntiles*S*cap * 2].  Requires mesh == gmesh (single process). */
int pic_retile(const PicParams* p, const void* x_in, const void* u_in, const uint8_t* active_in, void* x_out,
               void* u_out, uint8_t* active_out, int64_t cap, int32_t* scratch, int32_t* overflow, void* stream);

/* solvers/first_order_yee.py:42-72 : E[A] += dt*(C^2 curl B - J/eps) on every tile interior (no refresh). */
int pic_update_E(const PicParams* p, void* const E[3], const void* const B[3], const void* const J[3], void* stream);
/* solvers/first_order_yee.py:116-142 : B[A] -= (dt/2) curl E on every tile interior (half step). */
int pic_update_B(const PicParams* p, void* const B[3], const void* const E[3], void* stream);

/* The whole field half of the step -- update_B (half, E_old) -> update_E -> update_B (half, E_new), evolve.py:88-96 with
 * solvers/first_order_yee.py:12-162 -- for ONE ghosted tile in one pass: 15 reals per cell instead of 30 plus three refreshes.
 * Reads E, B (guard cells valid two deep) and J (interior folded; on an axis split across ranks also its first upper guard plane
 * refreshed), writes E_out, B_out != E, B: interior, the guard cells of single-rank periodic axes (wrapped copies), zeros stay in
 * the exterior guards of non-periodic walls; the guard cells of axes split across ranks are left to the halo exchange.
 * Conducting walls zero the tangential E on the first / last interior plane.  Bit-identical to the three sweeps + refreshes.
 * PIC_EUNSUPPORTED for g < 2 or alpha != 1 (the digital filter needs the separate sweeps). */
int pic_yee_fused(const PicParams* p, const void* const E[3], const void* const B[3], const void* const J[3], void* const E_out[3],
                  void* const B_out[3], void* stream);

/* utilities/filters.py:73/98 : out[A] = sum_27 k*in[A+off]; ghosts copied.  in != out. */
int pic_filter(const PicParams* p, int kind, double alpha, const void* in, void* out, void* stream);

/* boundary_conditions/ghost_cells.py:142-196 (_local_refresh_reduced_axis/_refresh_axis) for one axis over the
 * local tile mesh, in place, for `ncomp` arrays.  bc = the selected tuple's entry for this axis. */
int pic_halo_refresh_axis(const PicParams* p, int axis, int bc, int ncomp, void* const* fields, void* stream);
/* boundary_conditions/ghost_cells.py:218-289 (_local_fold_reduced_axis/_fold_axis incl. conducting wall). */
int pic_halo_fold_axis(const PicParams* p, int axis, int bc, int ncomp, void* const* fields, void* stream);
/* boundary_conditions/ghost_cells.py:344-362 _apply_local_zero_boundary_axis (planes g and -g-1 on wall tiles). */
int pic_zero_wall(const PicParams* p, int axis, void* field, void* stream);

/* Multi-GPU halo plumbing for one distributed axis (mesh[axis]==1 < gmesh[axis]): copy `nplanes` planes
 * starting at local index `start` of `ncomp` arrays to/from a packed buffer [comp][plane][transverse].
 * The packed faces travel by ncclSend/ncclRecv (ghost_cells.py:187,191,271,272 ppermute). */
int pic_pack_planes(const PicParams* p, int axis, int start, int nplanes, int ncomp, const void* const* fields,
                    void* buf, void* stream);
int pic_unpack_planes(const PicParams* p, int axis, int start, int nplanes, int ncomp, void* const* fields,
                      const void* buf, int mode, void* stream);

/* The same for up to 26 axis-aligned boxes of one ghosted tile in ONE launch: box b = [lo[3b..], lo + size[3b..]) in array
 * indices, packed as [comp][x][y][z] one box after the other.  These are the faces, edges and corners a rank exchanges with its
 * neighbours when all split axes are served in one round (the x -> y -> z sequence of ghost_cells.py:199-215 carries edges and
 * corners through three dependent rounds; sending them explicitly needs one). */
int pic_pack_boxes(const PicParams* p, int nbox, const int32_t* lo, const int32_t* size, int ncomp, const void* const* fields,
                   void* buf, void* stream);
int pic_unpack_boxes(const PicParams* p, int nbox, const int32_t* lo, const int32_t* size, int ncomp, void* const* fields,
                     const void* buf, int mode, void* stream);

/* ---- electrostatic field solve, one ghosted tile (mesh == 1,1,1) ----
 * solvers/electrostatic_yee.py:71-156 solve_poisson_with_conjugate_gradient: matrix-free CG for -lapl(phi) = rho/eps with the
 * reference's update order and stopping rule (k < max_iter && sum r^2 > tol^2); phi is the initial guess on entry and the
 * solution (ghosts filled per field_bc: periodic refresh / constant-potential conducting walls) on return.
 * work_r, work_p, work_lp: scratch arrays of the tile's ghosted size; scal: double[8] device scratch (scal[4] = iterations);
 * the host looks at the device-side "done" flag every `check_every` iterations; *iters_out (host, may be NULL) = iterations. */
int pic_poisson_cg(const PicParams* p, const void* rho, void* phi, void* work_r, void* work_p, void* work_lp, double* scal,
                   double tol, int max_iter, int check_every, int* iters_out, void* stream);
/* solvers/electrostatic_yee.py:20-37 _apply_tiled_phi_constant_boundaries on one tile (in place). */
int pic_phi_boundaries(const PicParams* p, void* field, void* stream);
/* boundary_conditions/ghost_cells.py:365-386: exterior ghost planes of `axis` := adjacent interior plane (in place). */
int pic_constant_wall(const PicParams* p, int axis, void* field, void* stream);
/* solvers/electrostatic_yee.py:212-246: E_c = -(phi[+1] - phi[-1]) / (2 d_c) on the tile interior (ghosts untouched). */
int pic_gradient_neg(const PicParams* p, const void* phi, void* const E[3], void* stream);

/* Conservation diagnostics (SURVEY.md section 8b pic_gauss_residual; not in the reference, which only tests the continuity
 * invariant: tests/code_tests/esirkepov_test.py:700-744).  out[A] = div_backward(F)[A] + ca*a[A] + cb*b[A] on every tile interior
 * (ghosts of `out` untouched; a / b may be NULL).  Gauss residual: F = E, a = rho, ca = -1/eps.  Discrete continuity:
 * F = J, a = rho_new, ca = 1/dt, b = rho_old, cb = -1/dt. */
int pic_div_residual(const PicParams* p, const void* const F[3], const void* a, double ca, const void* b, double cb, void* out,
                     void* stream);

/* utils.py:160-187 compute_energy pieces: out[0] += sum over interiors of f^2 (double accumulate). */
int pic_sum_squares_interior(const PicParams* p, const void* field, double* out, void* stream);
/* out[0] += sum active*(sqrt(p^2C^2+m^2C^4)-mC^2), out[1] += sum active*|v|*m  (utils.py:170-202). */
int pic_particle_energy(const PicParams* p, const void* u, const uint8_t* active, int64_t cap, double* out, void* stream);

/* ---------------- resident fast path: cell-sorted SoA particles, one tile per GPU ---------------------- */

/* One species in the resident layout: SoA rows x,y,z,vx,vy,vz of capacity `cap` (dead slots have x = NaN). */
typedef struct PicSoA {
    void* comp[6];
    int32_t* id;        /* original reference slot (tile-major s*cap+slot) or -1; may be NULL */
    int64_t cap;
    int64_t n;          /* host-side upper bound of the slots in use (sizes the launch) */
    int32_t* n_dev;     /* device counter of the slots in use (live + dead holes); when non-NULL the kernels read the
                           count from here, so appends / sorts never need a host round trip */
} PicSoA;

/* Per-direction migration packets of one species (multi-GPU).  Packet d occupies rows row_off[d] .. row_off[d]+cap[d] of
 * `buf` ([rows][7] reals): row row_off[d] is the header (its first 4 bytes are the int32 row count), the following
 * cap[d] rows are x,y,z,vx,vy,vz,species.  Fixed sizes let ranks exchange packets without first exchanging counts.
 * direction d = ((1-ox)*3 + (1-oy))*3 + (1-oz), cap[d] = 0 for directions that cannot occur. */
typedef struct PicLeave {
    void* buf;
    int32_t row_off[27];
    int32_t cap[27];
} PicLeave;

/* TiledParticles (reference layout, single local tile) -> compact SoA of species s.  d_count: int32 device counter
 * (zeroed by the caller) receiving the number of particles written. */
int pic_soa_import(const PicParams* p, int species, const void* x, const void* u, const uint8_t* active, int64_t cap_ref,
                   const PicSoA* soa, int32_t* d_count, void* stream);
/* SoA -> reference layout (slots by id when id != NULL, else compacted); outputs must be pre-zeroed. */
int pic_soa_export(const PicParams* p, int species, const PicSoA* soa, void* x, void* u, uint8_t* active, int64_t cap_ref,
                   int32_t* d_count, void* stream);

/* Counting sort by local cell (K2).  Three calls: histogram -> exclusive scan -> scatter.
 * cell_count/cell_offset: int32[ncells+1] (ncells = tile[0]*tile[1]*tile[2], last bin = dead particles). */
int pic_sort_histogram(const PicParams* p, const PicSoA* src, int32_t* cell_count, void* stream);
int pic_sort_scan(int64_t n, const int32_t* in, int32_t* out, int32_t* block_scratch, void* stream);
int pic_sort_scatter(const PicParams* p, const PicSoA* src, const PicSoA* dst, const int32_t* cell_offset,
                     int32_t* cell_cursor, void* stream);

/* K1: the fused step for one species over the resident layout --
 *   gather E,B (+ext) -> Boris/HC -> deposit (Esirkepov: x -> x+v*dt ; direct: at x+v*dt/2) -> move -> particle BC
 * == evolve.py:33-79 for one local tile.  J is accumulated with atomics into the ghosted tile (fold afterwards).
 * deposition: 0 = esirkepov, 1 = direct.  ext_E/ext_B may be NULL (external fields absent).
 * Leavers (multi-GPU, distributed axes) are written into the per-direction packets of `leave` (see PicLeave); pass NULL
 * when mesh == gmesh.  flags[0] |= 1 on an invalid (>1 tile) jump, |= 2 when a packet / SoA capacity overflowed. */
int pic_fused_push_deposit(const PicParams* p, int species, int deposition, const PicSoA* soa,
                           const void* const E[3], const void* const B[3], const void* const extE[3],
                           const void* const extB[3], void* const J[3], const PicLeave* leave, int32_t* flags,
                           void* stream);

/* K1 v9, the supercell tile variant of pic_fused_push_deposit for the headline configuration (Esirkepov, CIC, all three axes
 * active, g == 2, every tile width a multiple of 4, Boris or relativistic Boris, no external fields): same result, different
 * data path.  The blocked sort order keeps the particles of one 4x4x4-cell supercell contiguous; per supercell one thread issues
 * TMA copies of the 8x8x8-node E/B neighbourhood (padded to 8x10x8 in shared memory) of all six components and of the supercell's slice of the six particle
 * arrays into a three-stage shared-memory ring (mbarrier complete_tx), and the CTA's warps gather, push and deposit from
 * shared memory.  blk_off: int32[nblk+1] device array, blk_off[b] = first slot of supercell b in the cell-sorted SoA =
 * cell_offset[64*b] of the last pic_sort_scan (nblk = tile[0]*tile[1]*tile[2]/64); slots >= blk_off[nblk] (appended since
 * the sort) and particles that drifted more than one cell out of their supercell gather from global memory, so the result
 * does not depend on how stale the sort is.  The SoA rows should be 16-byte aligned (cap a multiple of 4 reals); otherwise
 * particles are read from global memory.  flags: int32[>=3]; flags[0] as for pic_fused_push_deposit, flags[2] counts the
 * particles that took the global-memory gather.  Returns PIC_EUNSUPPORTED for any other configuration or when the driver has
 * no cuTensorMapEncodeTiled (call pic_fused_push_deposit instead).  Replaces, like pic_fused_push_deposit, evolve.py:33-79.
 * options bit 0 (float32 only): accumulate the same-cell currents in per-supercell shared-memory J tiles (shared-memory
 * atomics) and flush each tile with one TMA reduce per component instead of global REDs (measured slower; kept as evidence).
 * options bit 1: reduce the same-cell currents over ALL lanes of a warp that share a cell (__match_any_sync groups summed by
 * pointer doubling) instead of the segmented scan over contiguous runs -- REDs no longer grow when cell changers fragment the
 * runs of a stale-sorted species, at the price of ~30 more instructions per 32 particles. */
int pic_fused_tile3d(const PicParams* p, int species, const PicSoA* soa, const int32_t* blk_off, int nblk, int options,
                     const void* const E[3], const void* const B[3], void* const J[3], const PicLeave* leave, int32_t* flags,
                     void* stream);

/* K1 v10, the pair variant of pic_fused_tile3d (same configurations, same result, same arguments): the per-thread body advances
 * two particles at once in packed f32x2 arithmetic (float; one at a time in double), a dedicated producer warp feeds the TMA
 * ring.  Two launches: the tile kernel advances every particle that stays in its cell and is covered by its supercell's
 * tile, and writes the slot indices of all others (cell changers: union-stencil deposit, wrap / reflect / absorb / ownership;
 * particles that drifted out of their tile since the last sort) into `work`, with the pre-move position of every cell
 * changer; the fix-up kernel then finishes the list (cell changers: deposit + boundary conditions + store; uncovered slots
 * and slots appended since the sort: the whole step through the scalar global-memory body).  `work`: device scratch of
 * work_len >= pic_pair_work_bytes(p, soa->cap) bytes, 16-byte aligned, contents meaningless between calls.  Requires the supercell slices of `blk_off` to
 * start on 16-byte boundaries, i.e. the SoA to come from the blocked, padded sort below (flags[0] |= 8 otherwise).
 * options bit 1 as for pic_fused_tile3d; bit 2 (float32): reduce the same-cell currents through per-warp shared-memory rows
 * (one ballot finds the runs of equal cells, lane t sums one current component of run t / 3 and issues its four REDs) instead
 * of the segmented warp scan.
 * Replaces, like pic_fused_push_deposit, PyPIC3D/evolve.py:33-79 (particle_push -> Esirkepov_current ->
 * update_tiled_particle_positions -> refresh_tiled_particle_tiles) for one local tile. */
int pic_fused_pair3d(const PicParams* p, int species, const PicSoA* soa, const int32_t* blk_off, int nblk, int options,
                     const void* const E[3], const void* const B[3], void* const J[3], const PicLeave* leave, int32_t* flags,
                     void* work, int64_t work_len, void* stream);
int64_t pic_pair_work_bytes(const PicParams* p, int64_t cap);

/* Blocked, padded counting sort (tile widths multiples of 4): cells in 4x4x4-supercell-major order, every supercell's slice
 * padded with dead slots (x = NaN) to a multiple of 4 slots.  Sequence: pic_sort_histogram -> pic_sort_blocked_offsets ->
 * pic_sort_scatter (with the cell_offset computed here) -> pic_sort_blocked_finish.
 * blk_off: int32[nblk + 1] out, first slot of every supercell (blk_off[nblk] = padded length of the stream); blk_work:
 * int32[nblk + 1] scratch; scan_scratch as for pic_sort_scan; flags[0] |= 2 when the padded stream exceeds `cap`.
 * (No reference counterpart: the reference keeps fixed-capacity slots with an `active` mask, particles/particle_class.py:17-31.) */
int pic_sort_blocked_offsets(const PicParams* p, const int32_t* cell_count, int32_t* cell_offset, int32_t* blk_work,
                             int32_t* blk_off, int32_t* scan_scratch, int64_t cap, int32_t* flags, void* stream);
int pic_sort_blocked_finish(const PicParams* p, const int32_t* cell_offset, const int32_t* cell_count, const int32_t* blk_off,
                            const PicSoA* dst, void* stream);

/* Zero the 27 packet headers of `leave` (before K1 of a step). */
int pic_packets_reset(const PicParams* p, const PicLeave* leave, void* stream);
/* Append every received packet of `recv` (same layout as PicLeave) to the SoA tail, advancing soa->n_dev on the device. */
int pic_soa_append_packets(const PicParams* p, const PicSoA* soa, const PicLeave* recv, int32_t* flags, void* stream);

/* Microbenchmarks used for design evidence (profiles/): returns elapsed device ms for `iters` launches. */
int pic_microbench(int which, int iters, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* PIC_B200_H */

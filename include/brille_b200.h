/* brille_b200.h -- C ABI of the B200-native batched Q-point interpolation path of brille.
 *
 * The hot path this library replaces is, in the reference (paths relative to brille's repository):
 *
 *   wrap/_common_grid.hpp:272-341   pybind11 `ir_interpolate_at(Q, useparallel, threads, do_not_move_points)`
 *     src/bz_trellis.hpp:153-196    BrillouinZoneTrellis3::ir_interpolate_at      (bz_nest.hpp:90-128, bz_mesh.hpp:98-136)
 *       src/bz_move.cpp:165-296     BrillouinZone::ir_moveinto   (tau search 14-53,103-163; wedge rotation 257-285)
 *       src/trellis_poly.hpp:309-351  PolyTrellis::interpolate_at   (node lookup 382-434; nodes trellis_node.hpp:130-149,273-364)
 *       src/interpolatordual.hpp:149-155 / interpolator_at.tpp:91-127   permuted, phase-aligned linear interpolation
 *       src/interpolator.hpp:386-428 / interpolator_gamma.tpp:49-139    rotation back to Q (Gamma phase, real/recip/axial)
 *
 * The reference has no C API for this path (its only C-level artefact is the generated single header
 * brille.h); the functions below are what a cgo/ctypes/pybind stub binds instead of those C++ members.
 * INTEGRATION.md shows the few lines a brille maintainer adds to wrap/_common_grid.hpp.
 *
 * Conventions
 *   - plain pointers and sizes only; all arrays are C-contiguous, row-major
 *   - every function returns 0 on success or a negative B200_E_* code; b200_last_error() gives the message
 *     (the messages mirror the reference's exception texts so a binding can re-throw them unchanged)
 *   - like the reference the call is all-or-nothing: one Q that cannot be placed fails the call
 *   - host construction (lattice, symmetry, polyhedra, TetGen, fill(), sort()) stays brille's C++; its result is
 *     handed over ONCE as the flat tables declared here (b200_grid_create / b200_grid_set_data)
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with B200_E_CUDA
 */
#ifndef BRILLE_B200_H_
#define BRILLE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1

/* ---- error codes --------------------------------------------------------------------------------- */
#define B200_OK 0
#define B200_E_INVALID (-1)       /* bad argument / inconsistent tables                                   */
#define B200_E_CUDA (-2)          /* CUDA runtime failure (no device, out of memory, launch failure)       */
#define B200_E_OUTSIDE_BZ (-3)    /* "Not all points inside Brillouin zone"          bz_move.cpp:151-158   */
#define B200_E_OUTSIDE_WEDGE (-4) /* "... is outside of the irreducible BrillouinZone" bz_move.cpp:288-294 */
#define B200_E_NOT_FOUND (-5)     /* "interpolate_at failed to find N points"   trellis_poly.hpp:345-349,   */
                                  /* null-node access trellis_node.hpp:593-610, nest.hpp:412-415           */
#define B200_E_UNSUPPORTED (-6)   /* LengthUnit/RotatesLike combination not implemented interpolator.hpp:394-427 */
#define B200_E_NODATA (-7)        /* "The interpolation data must be filled before interpolating."         */

/* ---- enumerations (numeric values are brille's) --------------------------------------------------- */
enum b200_grid_kind { B200_GRID_TRELLIS = 0, B200_GRID_NEST = 1, B200_GRID_MESH = 2 };
/* src/enums.hpp:43  NodeType */
enum b200_node_type { B200_NODE_ASSUMED_NULL = 0, B200_NODE_FOUND_NULL = 1, B200_NODE_NULL = 2, B200_NODE_CUBE = 3, B200_NODE_POLY = 4 };
/* src/rotates.hpp:27 RotatesLike */
enum b200_rotates_like { B200_ROT_VECTOR = 0, B200_ROT_PSEUDOVECTOR = 1, B200_ROT_GAMMA = 2 };
/* src/enums.hpp:37 LengthUnit */
enum b200_length_unit { B200_LEN_NONE = 0, B200_LEN_ANGSTROM = 1, B200_LEN_INVERSE_ANGSTROM = 2, B200_LEN_REAL_LATTICE = 3, B200_LEN_RECIPROCAL_LATTICE = 4 };

/* flags of the interpolation entry points */
#define B200_FLAG_NO_MOVE 1u  /* do_not_move_points=True: Q already inside the irreducible zone (bz_trellis.hpp:160-164) */

/* per-Q status bits reported through b200_probe_t.status */
#define B200_ST_OUTSIDE_BZ 1u
#define B200_ST_OUTSIDE_WEDGE 2u
#define B200_ST_NOT_FOUND 4u
#define B200_ST_FALLBACK_TET 8u /* PolyNode "should contain" escape hatch was taken (trellis_node.hpp:295-306) */
#define B200_ST_NEIGHBOUR 16u   /* node_subscript stepped to a neighbouring node (trellis_poly.hpp:391-424)    */

/* ---- tables: Brillouin zone (everything BrillouinZone::ir_moveinto reads) --------------------------- */
typedef struct b200_bz_tables {
  int32_t transform_needed;   /* conventional->primitive transform used by moveinto (bz_move.cpp:111)            */
  int32_t P6t[9];             /* 6*P^T, int, Q_p = (P6t*Q)/6                      transform.hpp:160-197           */
  int32_t invPt[9];           /* q = invPt*q_p, tau = invPt*tau_p                 transform.hpp:200-238           */
  double w_recip_metric[9];   /* reciprocal metric of the working (primitive) lattice: LVec dot   array_functions.hpp:246-290 */
  double w_real_metric[9];    /* real-space metric of the working lattice: LVec::star()           array_lvec_methods.tpp:53-67 */
  double w_recip_volume;      /* lat.volume(inverse_angstrom) of the working lattice: LVec cross  array_functions.hpp:190-201 */
  double o_recip_metric[9];   /* same three for the conventional (outer) lattice: isinside re-check and wedge test */
  double o_real_metric[9];
  double o_recip_volume;
  double to_xyz[9];           /* B: x = B q (row-major 3x3)                       lattice_dual.hpp:672-685        */
  int32_t n_faces;            /* F = faces of the first Brillouin zone                                            */
  const double* pa;           /* (F,3) three points per face plane, working-lattice rlu     bz_move.cpp:123-126   */
  const double* pb;
  const double* pc;
  const double* normals;      /* (F,3) n/|n| in working-lattice rlu                          bz_move.cpp:137-138   */
  const int32_t* taus;        /* (F,3) round(2*face point)                                   bz_move.cpp:139       */
  const double* tau_lens;     /* (F)   lattice norm of taus                                  bz_move.cpp:140       */
  const double* ca;           /* (F,3) face plane points in the conventional lattice: isinside  bz_move.cpp:149    */
  const double* cb;
  const double* cc;
  int32_t n_wedge;            /* K = irreducible-wedge normals (may be 0)                    bz.hpp:757-763        */
  const double* wedge_normals;/* (K,3) conventional rlu                                                           */
  int32_t no_ir_mirroring;    /* 0 => the le_ge branch without tolerance is used             array2.tpp:666-667    */
  double float_tolerance;     /* BrillouinZone::float_tolerance  (approx_float::Config::reciprocal)  bz.hpp:114-115 */
  int32_t approx_tolerance;   /* BrillouinZone::approx_tolerance (approx_float::Config::digit)                    */
  int32_t n_ops;              /* G = order of the point group (with inversion if time reversal)  bz.hpp:743-746    */
  const int32_t* rotations;   /* (G,9) row-major R_j in PointSymmetry storage order                               */
  const int32_t* inverse_index;/* (G)  PointSymmetry::get_inverse_index                      pointsymmetry.cpp:131-142 */
  int32_t identity_index;     /* PointSymmetry::find_identity_index                          pointsymmetry.cpp:148-151 */
} b200_bz_tables_t;

/* ---- tables: PolyTrellis structure (trellis_poly.hpp, trellis_node.hpp) ------------------------------ */
typedef struct b200_trellis_tables {
  int32_t n_knots[3];
  const double* knots[3];       /* knots_[d], Cartesian 1/angstrom                trellis_poly.hpp:128            */
  uint32_t n_nodes;             /* (n_knots[0]-1)*(n_knots[1]-1)*(n_knots[2]-1)                                   */
  const uint8_t* node_type;     /* (n_nodes) NodeType                              trellis_node.hpp:436-449        */
  const uint32_t* node_index;   /* (n_nodes) index into the cube / poly arrays, 0xffffffff for null nodes         */
  uint32_t n_cubes;
  const uint32_t* cube_vertices;/* (n_cubes,8) order (000)(100)(110)(010)(101)(001)(011)(111) trellis_node.hpp:79-81 */
  uint32_t n_polys;
  const uint32_t* poly_offsets; /* (n_polys+1) CSR offsets into the tetrahedron arrays                            */
  uint32_t n_tets;
  const uint32_t* tet_vertices; /* (n_tets,4) vi_t                                  trellis_node.hpp:210            */
  const double* tet_circum;     /* (n_tets,4) ci_t: circumsphere centre xyz + radius                              */
  const double* tet_volume;     /* (n_tets)   vol_t                                                               */
  uint32_t n_vertices;
  const double* vertices;       /* (n_vertices,3) Cartesian                                                       */
} b200_trellis_tables_t;

/* ---- tables: Nest structure (nest.hpp) ------------------------------------------------------------------------
 * The tetrahedron tree flattened breadth-first: node 0 is the root (no tetrahedron of its own); the children of
 * node i are the nodes [child_begin[i], child_end[i]) in the order NestNode::branches() holds them.                */
typedef struct b200_nest_tables {
  uint32_t n_nodes;
  const uint32_t* node_vertices; /* (n_nodes,4) NestLeaf::vi                          nest.hpp:61                  */
  const double* node_circum;     /* (n_nodes,4) NestLeaf::centre_radius (xyz + radius) nest.hpp:62                  */
  const double* node_volume;     /* (n_nodes)   NestLeaf::volume_                      nest.hpp:63                  */
  const uint8_t* node_is_leaf;   /* (n_nodes)   NestNode::is_leaf()                    nest.hpp:143                 */
  const uint32_t* child_begin;   /* (n_nodes) */
  const uint32_t* child_end;     /* (n_nodes) */
  uint32_t n_vertices;
  const double* vertices;        /* (n_vertices,3) Nest::vertices_ (all vertices), Cartesian                        */
  double tolerance;              /* approx_.reciprocal<double>()                       nest.hpp:343-345,399-400     */
  int32_t digit;                 /* approx_.digit()                                                                 */
} b200_nest_tables_t;

/* ---- tables: Mesh structure (mesh.hpp, triangulation_layers.hpp) ----------------------------------------------
 * TetTri: n_layers tetrahedral meshes, coarse to fine; tetrahedron t of layer l is row tet_offset[l]+t of the
 * per-tetrahedron arrays and its vertex indices are local to layer l (row vert_offset[l]+v of `vertices`).
 * connections[l][t] (candidate tetrahedra of layer l+1 for tetrahedron t of layer l) is
 * conn_index[conn_offset[c] .. conn_offset[c+1]) with c = tet_offset[l]+t for l < n_layers-1.                     */
typedef struct b200_mesh_tables {
  uint32_t n_layers;
  const uint32_t* tet_offset;    /* (n_layers+1) */
  const uint32_t* vert_offset;   /* (n_layers+1) */
  const uint32_t* tets;          /* (tet_offset[n_layers],4) vertices_per_tetrahedron   triangulation_layers.hpp:51 */
  const double* centres;         /* (.,3) circum_centres                                                            */
  const double* radii;           /* (.)   circum_radii                                                              */
  const double* vol6;            /* (.)   6.0*volume(tet)                               triangulation_layers.hpp:267 */
  const double* vertices;        /* (vert_offset[n_layers],3) vertex_positions of every layer                       */
  const uint32_t* conn_offset;   /* (tet_offset[n_layers-1]+1) CSR offsets                                          */
  const uint32_t* conn_index;    /* layer-local tetrahedron indices of the next finer layer                         */
} b200_mesh_tables_t;

/* ---- tables: interpolation data (DualInterpolator, PermutationTable, GammaTable) ---------------------- */
typedef struct b200_interp_desc {
  const void* data;       /* (n_vertices, branches*span) row-major; double or complex<double> (re,im pairs)       */
  int32_t is_complex;     /* 0: double, 1: complex<double>                                                        */
  uint32_t branches;      /* modes per point                                interpolator.hpp:311-333              */
  uint32_t elements[3];   /* scalars, vector ELEMENTS (3N), matrix ELEMENTS (9N) per mode  interpolator.hpp:91     */
  int32_t rotates_like;   /* b200_rotates_like                                                                    */
  int32_t length_unit;    /* b200_length_unit                                                                     */
} b200_interp_desc_t;

typedef struct b200_data_tables {
  uint32_t n_vertices;
  b200_interp_desc_t values;   /* eigenvalues  (interpolated WITHOUT phase alignment)  interpolatordual.hpp:152   */
  b200_interp_desc_t vectors;  /* eigenvectors (phase aligned when complex)            interpolatordual.hpp:153   */
  /* permutations written by sort(); n_perm_rows <= 1 means "all identity" and the pair tables may be NULL        */
  uint32_t n_perm_rows;
  const uint32_t* perm_rows;   /* (n_perm_rows, branches); row 0 is the identity   permutation_table.hpp:285-291  */
  const uint32_t* cube_perm;   /* (n_cubes,8,8)  row index for the vertex pair (cube corner a -> corner b)         */
  const uint32_t* tet_perm;    /* (n_tets,4,4)   row index for the vertex pair (tet corner a -> corner b)          */
  /* GammaTable (phonon.hpp:91-194), required when either interpolator rotates like Gamma                         */
  uint32_t n_atoms;
  const uint32_t* gamma_F0;    /* (n_atoms, G) l = F0(k, r)                                                       */
  const uint32_t* gamma_vidx;  /* (n_atoms, G) index into gamma_vectors                                           */
  uint32_t n_gamma_vectors;
  const double* gamma_vectors; /* (n_gamma_vectors,3)  R^-1 r_l - r_k in real-lattice units                       */
  const double* rot_cart;      /* (G,9) A R A^-1, used when length_unit == angstrom     interpolator.hpp:409-423   */
} b200_data_tables_t;

/* ---- optional per-Q intermediate results (parity probes); every pointer may be NULL ------------------- */
typedef struct b200_probe {
  double* q_ir;      /* (nQ,3) q inside the irreducible zone, rlu                                                 */
  double* x_ir;      /* (nQ,3) the same point, Cartesian                                                          */
  int32_t* tau;      /* (nQ,3) reciprocal lattice vector, conventional lattice                                    */
  int32_t* ridx;     /* (nQ)   Ridx                                                                               */
  int32_t* invridx;  /* (nQ)   invRidx                                                                            */
  uint32_t* cell;    /* (nQ)   trellis: linear node index; nest/mesh: id of the containing tetrahedron            */
  int32_t* tet;      /* (nQ)   trellis: global tetrahedron index if a poly node was used, else -1                 */
  int32_t* n_vert;   /* (nQ)   number of (vertex, weight) pairs emitted                                           */
  uint32_t* vertex;  /* (nQ,8) vertex indices in emission order (unused slots 0xffffffff)                         */
  double* weight;    /* (nQ,8) weights in emission order (unused slots 0)                                         */
  uint32_t* status;  /* (nQ)   B200_ST_* bits                                                                     */
} b200_probe_t;

typedef struct b200_grid b200_grid_t; /* opaque: device-resident tables + workspace + streams of ONE GPU */

/* ---- life cycle ---------------------------------------------------------------------------------------- */
/* Replaces the per-call table building of the reference (bz_move.cpp:118-140, bz_trellis.hpp:182-186):
 * copies the tables to `device` and derives the device-side layouts.  `structure` points at the
 * b200_trellis_tables_t / b200_nest_tables_t / b200_mesh_tables_t selected by `kind`.                       */
int b200_grid_create(int kind, const b200_bz_tables_t* bz, const void* structure, int device, b200_grid_t** out);
/* Mirrors Grid::replace_data + PermutationTable::refresh after fill()/sort() (wrap/_common_grid.hpp:26-47). */
int b200_grid_set_data(b200_grid_t* grid, const b200_data_tables_t* data);
void b200_grid_destroy(b200_grid_t* grid);

/* ---- the hot path ----------------------------------------------------------------------------------------
 * ir_interpolate_at with HOST buffers: Q (nQ,3) f64 in rlu of the conventional lattice; vals_out
 * (nQ, values.branches*span) and vecs_out (nQ, vectors.branches*span) of the dtypes given in set_data.
 * Host<->device copies are streamed in chunks inside the call.  Replaces wrap/_common_grid.hpp:276-301.      */
int b200_ir_interpolate_at(b200_grid_t* grid, const double* Q, size_t nQ, uint32_t flags,
                           void* vals_out, void* vecs_out, b200_probe_t* probe);
/* Same with DEVICE buffers on the grid's GPU; enqueued on `stream` (a cudaStream_t, NULL = default stream)
 * and NOT synchronised unless `n_failed` is non-NULL (then the per-Q status is reduced and returned).        */
int b200_ir_interpolate_at_device(b200_grid_t* grid, const double* dQ, size_t nQ, uint32_t flags,
                                  void* d_vals_out, void* d_vecs_out, b200_probe_t* d_probe,
                                  void* stream, uint64_t* n_failed);
/* interpolate_at (first Brillouin zone only, no wedge rotation, no rotate_in_place): bz_trellis.hpp:105-121,
 * wrap/_common_grid.hpp:408-437.  Same buffer conventions as above (host buffers).                          */
int b200_interpolate_at(b200_grid_t* grid, const double* Q, size_t nQ, uint32_t flags,
                        void* vals_out, void* vecs_out, b200_probe_t* probe);
/* BrillouinZone.ir_moveinto (wrap/_bz.cpp:434-463) / moveinto (:378-405) on the device; host buffers;
 * fills probe->q_ir, tau, ridx, invridx, status (x_ir if requested).  ir = 1 ir_moveinto, 0 moveinto,
 * 2 ir_moveinto_wedge (wrap/_bz.cpp:498-520, bz_move.cpp:299-356: the wedge rotation of Q itself, no translation; ridx is
 * the operation with Q = R q_ir; like the reference it does not fail: a point no operation places comes back as q_ir = 0 with
 * operation 0 and B200_ST_OUTSIDE_WEDGE in probe->status), 3 isinside (wrap/_bz.cpp:378-384, bz.hpp:631-640: probe->status gets B200_ST_OUTSIDE_BZ
 * for the points outside the first Brillouin zone; never fails).                                             */
int b200_moveinto(b200_grid_t* grid, const double* Q, size_t nQ, int ir, b200_probe_t* probe);

/* ---- sort() on the device --------------------------------------------------------------
 * Replaces the parallel loop of DualInterpolator::sort() (interpolatordual.hpp:398-434): for every connected vertex
 * pair (i, j) of `pairs` (n_pairs x 2, i < j; the keys of brille's PermutationTable) the modes x modes cost matrix of
 * Interpolator::add_cost (interpolator_cost.tpp:18-58) is built from the data given to b200_grid_set_data and the
 * Jonker-Volgenant assignment (lapjv.hpp:281-538) is solved.  row_out (n_pairs x modes) receives the permutation the
 * reference stores for (i, j), col_out the one it stores for (j, i).  The cost configuration is what
 * set_flags_weights leaves in the Interpolators (interpolator.hpp:246-305): relative weights of the scalar / vector /
 * matrix parts and the vector cost function (0 sin^2 of the Hermitian angle, 1 distance, 2 1 - |<a|b>|^2, 3 vector
 * angle, 4 Hermitian angle).  Real-valued eigenvectors are refused (B200_E_UNSUPPORTED): the reference reads an
 * uninitialised buffer for them (utilities.tpp:530).  cost_out (optional, n_pairs x modes x modes) receives the cost
 * matrices.  Host buffers; synchronous.                                                                           */
typedef struct b200_sort_config {
  double values_costmult[3];
  double vectors_costmult[3];
  int32_t values_vector_cost;
  int32_t vectors_vector_cost;
} b200_sort_config_t;
int b200_grid_sort_pairs(b200_grid_t* grid, const uint32_t* pairs, size_t n_pairs, const b200_sort_config_t* config,
                         int32_t* row_out, int32_t* col_out, double* cost_out);

/* The assignment solver of sort() on its own: `n` cost matrices (n x modes x modes, host) in, the row and column solutions the
 * reference's lapjv (lapjv.hpp:281-538) returns for each of them out -- ties included.  Used by the tests to drive the solver
 * with matrices full of ties, which grids do not produce.  Runs on `device`; synchronous.                              */
int b200_solve_assignments(const double* cost, size_t n, uint32_t modes, int32_t* row_out, int32_t* col_out, int device);

/* ---- device-resident consumer: one-phonon structure factor ----------------------------------
 * What brille's callers do with the output of ir_interpolate_at straight away (Euphonic's BrilleInterpolator /
 * QpointPhononModes.calculate_structure_factor, brilleu's s_qw -- the loop validation/profiling.md:30-67 times; the
 * commented-out ir_interpolate_at_dw of wrap/_common_grid.hpp:343-405 fuses the same kind of per-atom Debye-Waller
 * reduction behind the interpolation):
 *
 *     F(Q,nu) = sum_k coef_k exp(-qv^T W_k qv) exp(2 pi i Q.r_k) (qv . eps_{nu,k}(Q)^[*]),     sf(Q,nu) = |F(Q,nu)|^2
 *
 * with eps the interpolated, rotated eigenvectors exactly as ir_interpolate_at returns them, Q the input point (rlu)
 * and qv = q_transform Q.  The eigenvectors stay in HBM: per Q only the eigenvalue row and `modes` doubles leave the
 * device instead of 16*modes*3*atoms bytes.  Requires complex eigenvector data made of 3-vectors only, one per atom
 * (vectors.elements = {0, 3*n_atoms, 0}); anything else is B200_E_UNSUPPORTED.  The arrays are copied by set.
 * When the pipelined cell kernel runs (Gamma-rotated data, at most 32 atoms, enough points) the reduction
 * is FUSED into its finish: the eigenvectors are formed in registers, reduced with warp shuffles and never written anywhere;
 * the few points that sit on cell faces take the general kernel into a compact scratch (nQ/16 rows) and the list mode of the
 * reduction kernel.  Otherwise the path writes the eigenvectors to a device scratch and k_structure_factor reduces them.
 * Both give the same numbers to rounding (the fused finish sums in a different order); each is bit-reproducible across
 * chunkings.  Option "sf_fused" (b200_grid_set_option) 1 (default) / 0 selects.                                           */
typedef struct b200_sf_config {
  uint32_t n_atoms;
  const double* coef;        /* (n_atoms,2) complex coefficient per atom (re,im), e.g. b_k/sqrt(m_k)                   */
  const double* positions;   /* (n_atoms,3) fractional atom positions r_k, or NULL: no exp(2 pi i Q.r_k) factor         */
  const double* debye_waller;/* (n_atoms,9) symmetric W_k in the basis of qv, or NULL: no Debye-Waller factor           */
  double q_transform[9];     /* row-major: identity when the eigenvectors are in lattice units, B for Cartesian ones   */
  int32_t conjugate;         /* 1: qv . conj(eps) (Euphonic), 0: qv . eps                                              */
} b200_sf_config_t;
int b200_grid_set_structure_factor(b200_grid_t* grid, const b200_sf_config_t* config);
/* HOST buffers: vals_out as in b200_ir_interpolate_at, sf_out (nQ, vectors.branches) doubles.                         */
int b200_ir_structure_factor(b200_grid_t* grid, const double* Q, size_t nQ, uint32_t flags, void* vals_out, double* sf_out);
/* DEVICE buffers, enqueued on `stream`.  d_vecs_scratch: room for the eigenvectors of all nQ points
 * (nQ * vectors row bytes), or NULL: the library keeps a scratch of its own (sized to at most a third of the free
 * memory) and walks the points in as few chunks as fit.  Synchronises only if n_failed is non-NULL.  The fused finish is
 * used when d_vecs_scratch is NULL and n_failed is not (the call must read a counter back to know that the compact scratch
 * sufficed); a caller who passes a scratch gets the eigenvectors of the call in it.                                  */
int b200_ir_structure_factor_device(b200_grid_t* grid, const double* dQ, size_t nQ, uint32_t flags, void* d_vals_out,
                                    double* d_sf_out, void* d_vecs_scratch, void* stream, uint64_t* n_failed);

/* ---- device-resident consumer: powder average ("powder binning of |F(Q)|^2 per mode") ---------------------
 * What a powder (orientationally averaged) calculation does with ir_interpolate_at (Euphonic / brilleu; the loop
 * validation/profiling.md:30-67 times): the sphere of directions is sampled at every |Q|, and the one-phonon intensity
 * |F(Q, nu)|^2 of every mode is binned on (|Q|, omega_nu(Q)).  Here the structure factor is reduced on the device
 * (b200_grid_set_structure_factor must have been called) and accumulated into the histogram on the device: nothing per Q
 * crosses PCIe.  hist[iq * n_wbins + iw] += weight(|F|^2) for |B Q| in |Q| bin iq and the FIRST scalar of mode nu's eigenvalue
 * in energy bin iw; counts[iq] = number of points in |Q| bin iq (the normalisation of the average).                         */
typedef struct b200_powder_config {
  uint32_t n_qbins, n_wbins;
  double q_lo, q_hi;   /* |Q| axis, 1/angstrom                                                                          */
  double w_lo, w_hi;   /* energy axis, units of the eigenvalues                                                         */
  int32_t weight;      /* 0: |F|^2; 1: |F|^2 / omega for omega > 0 (the 1/omega of the one-phonon cross section)         */
} b200_powder_config_t;
/* caller-provided points (host, rlu), accumulated INTO hist_out (n_qbins x n_wbins) / counts_out (n_qbins): host doubles    */
int b200_ir_powder_bin(b200_grid_t* grid, const double* Q, size_t nQ, uint32_t flags, const b200_powder_config_t* config,
                       double* hist_out, double* counts_out);
/* the sweep with the points generated on the device: for every |Q| bin (|Q| at its centre) the directions j in [dir_lo, dir_hi)
 * of a reproducible sequence of n_dir isotropic directions (counter-based: splitmix64 of seed, bin and j), Q = B^-1 (|Q| d).
 * Ranks / GPUs take disjoint direction ranges and add their histograms.  b200_powder_points returns the points of the same
 * sequence on the host (for checks against the reference).                                                              */
int b200_ir_powder_sweep(b200_grid_t* grid, const b200_powder_config_t* config, uint64_t n_dir, uint64_t seed, uint64_t dir_lo,
                         uint64_t dir_hi, double* hist_out, double* counts_out);
int b200_powder_points(b200_grid_t* grid, const b200_powder_config_t* config, uint64_t n_dir, uint64_t seed, uint64_t dir_lo,
                       uint64_t dir_hi, double* Q_out);

/* Page-locked host memory for Q / output buffers: with pinned buffers the chunked copies of the host-buffer
 * entry points run at the PCIe rate and overlap the kernels.  Pageable output buffers work too: the library lands
 * every chunk in a page-locked bounce buffer of its own and moves it on with several host threads (about half the
 * rate of pinned buffers; option "bounce" 0 restores plain device-to-pageable copies).                      */
void* b200_alloc_pinned(size_t bytes);
void b200_free_pinned(void* ptr);

/* ---- introspection ------------------------------------------------------------------------------------- */
const char* b200_last_error(void);
int b200_abi_version(void);
int b200_device_count(void);
/* number of kernels this library has launched on the grid since creation (bench.py's gpu_launches claim)  */
uint64_t b200_grid_launch_count(const b200_grid_t* grid);
/* which kernels the last interpolation call enqueued (a diagnostic for tests / the smoke run: a call with few points takes the
 * single-kernel location and the general interpolation kernel, the bench takes the two-kernel location and the pipelined kernel) */
#define B200_PATH_SPLIT_LOCATE 1u   /* two-kernel location, points regrouped by node / spatial bin in between           */
#define B200_PATH_CELL_ONTHEFLY 2u  /* cell-batched kernel that stages and aligns the vertex rows per work item          */
#define B200_PATH_CELL_PIPELINED 4u /* persistent pipelined cell kernel fed from the per-cell record table               */
#define B200_PATH_GENERAL 8u        /* the general per-(Q, mode) kernel handled ALL points (not only the leftovers)       */
#define B200_PATH_SF_FUSED 16u      /* structure factor reduced inside the pipelined cell kernel                          */
#define B200_PATH_COOP_LOCATE 32u   /* second location kernel in its warp-cooperative form (node records staged in shared memory) */
uint32_t b200_grid_last_path(const b200_grid_t* grid);
/* average device time in ms of the kernels launched by the last *_device call, measured with CUDA events
 * on the launching stream when timing was enabled with b200_grid_enable_timing(grid, 1).
 * names: "locate" (and its parts "locate_a", "locate_sort", "locate_b" when the two-kernel location ran), "sort", "interpolate",
 * "consumer" (k_structure_factor of the unfused device-buffer call); returns <0 if
 * unknown / not timed.                                                                                              */
int b200_grid_enable_timing(b200_grid_t* grid, int on);
double b200_grid_kernel_ms(const b200_grid_t* grid, const char* name);
/* tuning knobs: "interp_path" 0 auto | 1 general per-(Q,mode) kernel | 2 cell-batched kernel whenever the data layout
 * allows it; "chunk" points per CTA work item of the cell-batched kernel (default 256, reduced automatically
 * when many atoms/modes would not fit shared memory); "cell_kernel" 0 auto | 1 cell kernel that stages and phase-aligns
 * the vertex rows on the fly | 2 persistent pipelined cell kernel fed from the per-cell record table (built once per
 * fill; auto uses it whenever the table fits in a quarter of the free device memory); "tile" points per register tile
 * of the pipelined kernel (4: 124 registers, 2 CTAs/SM, default; 2: 80 registers, 3 CTAs/SM - measured slower); "host_chunk" upper bound on
 * the points per chunk of the host-buffer pipeline (0 = sized from free device memory); "sf_fused" 1 (default) / 0: structure
 * factor reduced inside the pipelined cell kernel whenever possible / always through the eigenvector scratch; "bounce" 1
 * (default) / 0: pageable host destinations through page-locked bounce buffers and host threads / plain device-to-pageable
 * copies; "split_locate" 1 (default) / 0: two-kernel location with the points regrouped in between / single kernel;
 * "coop_locate" 1 (default) / 0: trellis, second location kernel warp-cooperative (node records staged in shared memory) / per lane;
 * "replay_stores" 1: diagnostic, with timing enabled the output stores of the pipelined kernel are replayed on their own
 * and timed as "replay" (b200_grid_kernel_ms)                                                                       */
int b200_grid_set_option(b200_grid_t* grid, const char* name, double value);
/* output row sizes in bytes for one Q (values, vectors) and algorithmic HBM bytes per Q of the path          */
int b200_grid_row_bytes(const b200_grid_t* grid, size_t* vals_bytes, size_t* vecs_bytes);

#ifdef __cplusplus
}
#endif
#endif /* BRILLE_B200_H_ */

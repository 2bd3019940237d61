// device_tables.cuh -- device-side layouts derived from the flat host tables (include/brille_b200.h)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

constexpr int MAX_FACES = 32;  // faces of a first Brillouin zone (<= 14 for 3-D lattices)
constexpr int MAX_OPS = 48;    // order of a crystallographic point group
constexpr int MAX_WEDGE = 16;  // irreducible wedge normals
constexpr int MAX_KNOTS = 1024;
constexpr int MAX_WEDGE_FAST = 6;
constexpr int MAX_WPLANES = 12;

// Everything ir_moveinto needs, small enough (< 12 KB) to be staged in shared memory by every CTA.
struct BZDev {
  // approx_float tolerances (approx_float.hpp:78-100): configured (tol,digit) and default (0,1)
  double cfg_rel, cfg_abs, def_rel, def_abs;
  int transform_needed, n_faces, n_wedge, n_ops, no_ir_mirroring, identity_index;
  double P6t[9];    // int matrix stored as double (products int*double are exact conversions)
  double invPt[9];
  int invPt_i[9];
  double w_recip_metric[9], w_real_metric[9], w_recip_volume;
  double o_recip_metric[9], o_real_metric[9], o_recip_volume;
  double to_xyz[9];
  // working-lattice faces: plane points (exact fallback) and covector fast path det_f(q) = sum_i (a_i-q_i) m_i
  double pa[MAX_FACES][3], pb[MAX_FACES][3], pc[MAX_FACES][3], pm[MAX_FACES][3];
  // conventional-lattice faces for the isinside re-check
  double ca[MAX_FACES][3], cb[MAX_FACES][3], cc[MAX_FACES][3], cm[MAX_FACES][3];
  double normals[MAX_FACES][3];  // n/|n|
  int taus[MAX_FACES][3];
  double tau_lens[MAX_FACES];
  double gw[MAX_WEDGE][3];  // (G* n_k) of same_lattice_dot, computed once in the reference's order
  double Rt[MAX_OPS][9];    // transposed rotations as doubles
  int inverse_index[MAX_OPS];
  // certified fast wedge test: (G* n_k).(R_j^T q) == (R_j G* n_k).q =: wc[j][k].q ; used when n_wedge <= MAX_WEDGE_FAST
  // and no_ir_mirroring; a value within eps_wedge of the tolerance threshold falls back to the reference arithmetic
  int wedge_fast;
  double eps_wedge;
  double wc[MAX_OPS][MAX_WEDGE_FAST][3];
  double inv_tau_lens[MAX_FACES];  // 1/|tau_j| for the certified fast rounding of d_j/|tau_j|
  // sign-pattern wedge lookup: the wedge-bounding planes of all G images of the wedge reduce to n_wplanes distinct planes
  // through the origin; the signs of q on them select the operation through wtable (0xff = none).  A point within wband
  // of any plane takes the in-order scan with the reference arithmetic instead.
  int n_wplanes;  // 0 => lookup not available
  double wband;
  double wplane[MAX_WPLANES][3];
  uint8_t wtable[1 << MAX_WPLANES];
};

struct TrellisDev {
  int n_knots[3];
  int knot_offset[3];
  const double* knots;  // concatenated knot vectors
  double knot0[3], knot_inv[3];  // first knot and bins per unit length of every axis: the guess of find_bin
  uint32_t n_nodes;
  const uint8_t* node_type;
  const uint32_t* node_index;
  const uint32_t* cube_vertices;  // (n_cubes, 8)
  const double* cube_pack;        // (n_cubes, 24): xyz of the 8 corners
  const uint32_t* poly_offsets;
  const uint32_t* tet_vertices;   // (n_tets, 4)
  const double* tet_pack;         // (n_tets, 18): cx cy cz r^2 | v0 v1 v2 v3 (xyz) | 6*vol | pad
  uint32_t n_cubes, n_tets, n_vertices;
};
constexpr int TET_PACK = 18;

// Nest: breadth-first flattened tetrahedron tree (nest.hpp); node 0 = root
struct NestDev {
  uint32_t n_nodes, n_vertices;
  const double* node_pack;        // (n_nodes, 18): cx cy cz r^2 | v0 v1 v2 v3 (xyz) | 6*vol | pad
  const uint32_t* node_vertices;  // (n_nodes, 4)
  const uint8_t* node_is_leaf;
  const uint32_t* child_begin;
  const uint32_t* child_end;
  double rel, abs_;               // approx tolerance pair of (approx_.reciprocal, approx_.digit)
};

constexpr int NEST_MAX_DEPTH = 48;  // levels of a Nest the device walks (brille's default max_branchings is 5)

// Mesh: layered tetrahedral meshes (triangulation_layers.hpp)
struct MeshDev {
  uint32_t n_layers, n_tets_last;
  const uint32_t* tet_offset;     // (n_layers+1)
  const double* tet_pack;         // (all tets, 18): cx cy cz radius | v0 v1 v2 v3 (xyz) | 6*vol | pad
  const uint32_t* tets;           // (all tets, 4) layer-local vertex indices
  const uint32_t* conn_offset;
  const uint32_t* conn_index;
};

// what the bucketing / cell kernel needs to know about the cells of any grid kind
struct CellsDev {
  uint32_t n_cubes, n_tets;        // trellis: cubes + tetrahedra; nest: 0 + nodes; mesh: 0 + finest-layer tetrahedra
  const uint32_t* cube_vertices;   // (n_cubes, 8)
  const uint32_t* tet_vertices;    // (n_tets, 4)
  const uint32_t* node_index;      // trellis only: payload index of a node
};

// Nest / Mesh: a uniform grid of bins over the bounding box of the vertices.  It plays the part of the trellis nodes in the
// two-kernel location: the points are regrouped by bin between the kernels, so that the lanes of a warp descend the tree / the
// layers next to each other (same branches, shared loads).  It only orders the work: the location itself is unchanged.
struct BinDev {
  double lo[3], inv[3];  // finest bin of x along d: floor((x[d] - lo[d]) * inv[d]), clamped to [0, n[d])
  int n[3];              // finest level (64 per axis with an extent); a call uses every 2^shift-th bin (LocateOut::bin_shift)
  uint32_t total;        // bins of the finest level; 0: not available
  // candidate tetrahedra per finest bin (CSR, ascending ids): every leaf (Nest) / finest-layer tetrahedron (Mesh) whose bounding
  // box, slightly widened, meets the bin.  The fast path of the Nest / Mesh location tests only these (nest_locate, mesh_locate).
  const uint32_t* cand_offset;  // (total + 1), or null
  const uint32_t* cand_index;
};
__host__ __device__ inline uint32_t bins_at_level(const BinDev& b, int shift) {
  uint32_t t = 1;
  for (int d = 0; d < 3; ++d) t *= (uint32_t)(((b.n[d] - 1) >> shift) + 1);
  return t;
}

struct GridDev {
  int kind;  // b200_grid_kind
  BinDev bins;
  TrellisDev tr;
  NestDev ne;
  MeshDev me;
  CellsDev cells;
};

// The weight row of a point is the head of a record that holds everything the cell-batched interpolation needs about the
// point:  weight[8] | q_ir[3] | (ridx | invridx << 16 , point index) | 32 bytes of padding.  The pipelined cell kernel fetches
// it with one bulk copy per point.  The 96 bytes of content are padded to one whole, aligned 128-byte line: the bulk-copy
// engine fetches every line a copy touches, and a 96-byte record at a 96-byte stride touches 1.5 lines on average (ncu:
// 2.14 GB of DRAM reads per 1e7 points for 0.96 GB of records; with the padding the kernel is 1.3 % faster).
constexpr uint32_t REC_DOUBLES = 16, REC_BYTES = 8 * REC_DOUBLES, REC_USED_BYTES = 96;  // stride in doubles / bytes, content
constexpr uint32_t REC_SMEM_DOUBLES = REC_USED_BYTES / 8;  // the records of an item are packed in shared memory

// A point parked between the two kernels of the split trellis location: what the second kernel needs, one 32-byte sector
// (the first kernel already wrote the q_ir / rotation / index part of the point's record).
struct ParkedPoint {
  double x[3];
  uint32_t rot_st;  // ridx | invridx << 8 | status << 16
  uint32_t cell;
};

// per-Q result of the locate stage, consumed by the interpolation stage (SoA, all device pointers)
struct LocateOut {
  double* q_ir;      // (n,3)
  double* x_ir;      // (n,3) optional
  int32_t* tau;      // (n,3) optional
  int32_t* ridx;     // (n)
  int32_t* invridx;  // (n)
  uint32_t* cell;    // (n) node linear index
  int32_t* tet;      // (n) global tetrahedron index or -1
  int32_t* n_vert;   // (n)
  uint32_t* vertex;  // (n,8)
  double* weight;    // (n,REC_DOUBLES): 8 weights + the rest of the point's record
  uint64_t* slots;   // (n) 8 packed corner slots (emission order), byte j = slot of emitted vertex j
  uint32_t* status;  // (n)
  // bucketing for the cell-batched interpolation kernel (all optional: NULL => not bucketed)
  uint32_t* key;         // (n) sub-bucket of the point: bucket * sub + invridx, where the bucket is c for cube c,
                         //     n_cubes + t for tetrahedron t and n_cubes + n_tets for everything that is not a full
                         //     generic cell (some weight ~ 0, failed points)
  uint32_t* rank;        // (n) arrival order of the point inside its sub-bucket
  uint32_t* cell_count;  // ((n_cubes + n_tets + 1) * sub) sub-bucket populations, zeroed before the launch
  // split trellis location (MODE_SPLIT_A): parked points and their node buckets (key/rank double as node bucket/rank)
  ParkedPoint* parked;   // (n)
  int lean;              // second kernel: points of a cell bucket write only their record (no probe was asked for)
  int bin_shift;         // nest / mesh: coarsening of the spatial bins for this call (about 100 points per bin)
  uint32_t* node_count;  // (n_nodes + 1), zeroed before the launch; bucket n_nodes = no node found
  uint32_t sub;          // sub-buckets per bucket = number of point group operations: the sort is by (cell, operation), so
                         //     that consecutive points of a cell share the rotation matrix
};

struct InterpDev {
  const double* data;  // (n_vert, branches*span) double or complex (as double pairs)
  int is_complex;
  uint32_t branches, span, no0, no1, no2;  // scalars, #3-vectors, #3x3 matrices per mode
  int rot_kind;  // 0 real vector, 1 reciprocal vector, 2 axial, 3 gamma (int R), 4 gamma (cartesian R), -1 nothing to rotate
};

struct DataDev {
  InterpDev values, vectors;
  uint32_t n_perm_rows;
  const uint32_t* perm_rows;  // (n_perm_rows, branches)
  const uint32_t* cube_perm;  // (n_cubes, 64)
  const uint32_t* tet_perm;   // (n_tets, 16)
  uint32_t n_atoms, n_ops;
  const uint32_t* gamma_F0;
  const uint32_t* gamma_vidx;
  const double* gamma_vectors;
  const double* rot_int;   // (G,9) rotations as doubles
  const double* rot_cart;  // (G,9)
  const double* rot_det;   // (G)
  const uint8_t* rot_is_identity;  // (G)
};

// per-Q record of the locate stage as the interpolation kernels read it
struct LocateIn {
  const double* q_ir;
  const int32_t* ridx;
  const int32_t* invridx;
  const uint32_t* cell;
  const int32_t* tet;
  const int32_t* n_vert;
  const uint32_t* vertex;
  const double* weight;
  const uint64_t* slots;
  const uint32_t* status;
  const uint8_t* node_type;    // trellis: type of node `cell`
  const uint32_t* node_index;  // trellis: payload index of node `cell`
};

// ---- cell-batched interpolation (cellinterp.cu) ------------------------------------------------------------------
struct CellItem {
  uint32_t key;    // bucket (cube c -> c, tetrahedron t -> n_cubes + t)
  uint32_t start;  // first position in the sorted arrays
  uint32_t len;    // number of points (<= chunk)
};

struct BucketDev {
  uint32_t n_buckets;          // n_cubes + n_tets + 1
  uint32_t sub;                // sub-buckets per bucket (point group operations)
  uint32_t chunk;              // points per CTA item
  const uint32_t* cell_count;  // (n_buckets * sub)
  uint32_t* cell_offset;       // (n_buckets * sub) exclusive scan of cell_count: first position of every sub-bucket
  uint32_t* cell_total;        // (n_buckets) bucket populations
  uint32_t* cell_start;        // (n_buckets) first position of every bucket
  uint32_t* item_start;        // (n_buckets) index of the first work item of every bucket
  uint32_t max_items;          // capacity of `items` (0: no work items wanted, e.g. the node buckets of the split location)
  CellItem* items;             // (max_items)
  uint32_t* n_items;           // [0] number of items, [1] start of the last (general) bucket, [2] its population
  uint32_t* order;             // (n) point indices in bucket order
};

// one-phonon structure factor configuration (device copy of b200_sf_config_t)
struct SFDev {
  uint32_t n_atoms;
  const double* coef;  // (n_atoms,2) complex coefficient per atom
  const double* pos;   // (n_atoms,3) fractional positions, or null: no exp(2 pi i Q.r) factor
  const double* dw;    // (n_atoms,9) Debye-Waller matrices in the basis of qv, or null
  double T[9];         // qv = T Q (row-major)
  int conjugate;       // 1: qv . conj(eps)
};

struct CellArgs {
  DataDev dd;
  const uint32_t* cube_vertices;
  const uint32_t* tet_vertices;
  uint32_t n_cubes;
  BucketDev bk;
  const double* weight;    // (n,8) per-point records of the locate stage, gathered through `order`
  const double* q_ir;      // (n,3)
  const int32_t* ridx;     // (n)
  const int32_t* invridx;  // (n)
  double* vals_out;
  double* vecs_out;
  int ir;
  uint32_t modes_per_pass;  // modes staged per pass (<= branches)
  // fused structure-factor finish of the pipelined kernel (sf_out != nullptr): vecs_out is not written
  const double* Q;          // (n,3) the input points
  double* sf_out;           // (n, branches)
  SFDev sf;
};

// pre-aligned per-cell records of the pipelined cell kernel (cellinterp_tma.cu)
struct CellTableDev {
  uint32_t n_cubes, n_tets;
  uint32_t mpp, n_pass;          // modes per pass, number of passes (record = n_pass tiles)
  uint64_t cube_bytes, tet_bytes;  // record sizes
  uint64_t total_bytes;
};

// mode bits of the locate kernel
constexpr uint32_t MODE_NO_MOVE = 1u;    // skip moveinto/ir_moveinto (do_not_move_points)
constexpr uint32_t MODE_IR = 2u;         // ir_moveinto (wedge rotation) rather than moveinto
constexpr uint32_t MODE_NO_LOCATE = 4u;  // moveinto only (b200_moveinto)
constexpr uint32_t MODE_SPLIT_A = 8u;    // stop after the node (trellis) / spatial bin (nest, mesh) is found and park the point (two-kernel location)
constexpr uint32_t MODE_NO_TAU = 16u;    // with MODE_IR: wedge rotation only, no translation (BrillouinZone::ir_moveinto_wedge)
constexpr uint32_t MODE_ISINSIDE = 32u;  // only test the point against the first-zone planes (BrillouinZone::isinside): status bit, no failure


// Function attributes (dynamic shared memory limit, occupancy) belong to a device: launchers that cache them index the cache by
// the current device, so that one process can drive several GPUs (brille_b200/sharding.py: ShardedGrid).
constexpr int MAX_DEVICES = 64;
inline int current_device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES) d = 0;
  return d;
}

// launchers (one per .cu file)
cudaError_t launch_locate(const BZDev* bzg, const GridDev& gd, const double* Q, size_t n, uint32_t mode, double eps_w,
                          double eps_o, const LocateOut& out, unsigned long long* fail_count, int sm_count,
                          cudaStream_t stream);
cudaError_t launch_locate_in_node(const BZDev* bzg, const GridDev& gd, size_t n, uint32_t mode, const LocateOut& out, const uint32_t* order,
                                  unsigned long long* fail_count, int sm_count, cudaStream_t stream, bool coop = true);
cudaError_t launch_interp(const DataDev& dd, const LocateIn& in, size_t n, int ir, double* vals, double* vecs, int sm_count,
                          cudaStream_t stream, const uint32_t* order, const uint32_t* segment, uint32_t compact_cap = 0,
                          unsigned long long* overflow = nullptr);
bool cell_path_eligible(const DataDev& dd);
uint32_t cell_modes_per_pass(const DataDev& dd, bool has_cubes, uint32_t chunk, size_t budget);
uint32_t cell_pick_chunk(const DataDev& dd, bool has_cubes, uint32_t preferred, size_t budget, uint32_t* mpp_out);
cudaError_t launch_bucket_sort(const BucketDev& bk, const uint32_t* key, const uint32_t* rank, size_t n, int sm_count,
                               cudaStream_t stream, const uint32_t* index = nullptr);
cudaError_t launch_interp_cell(const CellArgs& args, size_t n, cudaStream_t stream);
uint32_t cell_tma_pick(const DataDev& dd, bool has_cubes, uint32_t preferred, size_t budget, uint32_t* mpp_out, bool sf = false);
CellTableDev cell_table_layout(const DataDev& dd, uint32_t n_cubes, uint32_t n_tets, uint32_t mpp);
cudaError_t launch_build_cell_table(const DataDev& dd, const uint32_t* cube_vertices, const uint32_t* tet_vertices,
                                    const CellTableDev& ct, unsigned char* table, int sm_count, cudaStream_t stream);
cudaError_t launch_interp_cell_tma(const CellArgs& args, const CellTableDev& ct, const unsigned char* table, size_t n,
                                   int sm_count, cudaStream_t stream, int tile);
bool cell_sf_fusable(const DataDev& dd, const SFDev& sf);
cudaError_t launch_store_replay(const CellArgs& args, size_t n, int sm_count, cudaStream_t stream);

// device work space of sort() (sortpairs.cu), kept by the grid between calls
struct SortWorkspace {
  size_t batch = 0;
  uint32_t branches = 0;
  uint32_t* pairs = nullptr;
  double* cost = nullptr;
  int* row = nullptr;
  int* col = nullptr;
  cudaError_t ensure(size_t n_pairs, uint32_t branches);
  void release();
};
cudaError_t run_sort_pairs(const DataDev& dd, const double v_mult[3], int v_vfun, const double w_mult[3], int w_vfun, const uint32_t* h_pairs,
                           size_t n_pairs, int32_t* h_row, int32_t* h_col, double* h_cost, int sm_count, size_t max_ws_bytes,
                           SortWorkspace& ws, uint64_t* launches);
cudaError_t run_match_only(const double* h_cost, size_t n, uint32_t B, int32_t* h_row, int32_t* h_col, int sm_count);

}  // namespace b200

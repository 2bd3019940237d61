// capi.cu -- the C ABI of include/brille_b200.h: table upload, workspace, streams, chunked host pipeline.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "brille_b200.h"
#include "device_tables.cuh"
#include "consumer.cuh"


using namespace b200;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return fail(B200_E_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call);     \
  } while (0)

// The device-buffer entry points run underneath a caller (torch) that has a current device of its own: it is restored on return.
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// owning device buffer list
struct DevPool {
  std::vector<void*> ptrs;
  template <class T>
  cudaError_t upload(const T* host, size_t count, const T** out) {
    *out = nullptr;
    if (count == 0) return cudaSuccess;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e != cudaSuccess) return e;
    ptrs.push_back(p);
    e = cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice);
    *out = static_cast<const T*>(p);
    return e;
  }
  void release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
  }
};

struct Workspace {
  size_t capacity = 0;
  LocateOut lo{};
  double* x_ir = nullptr;
  int32_t* tau = nullptr;
  // counting sort by cell (cellinterp.cu)
  uint32_t n_buckets = 0, chunk = 0, sub = 0;
  uint32_t* key = nullptr;
  uint32_t* rank = nullptr;
  uint32_t* cell_count = nullptr;
  BucketDev bk{};
  // split trellis location: parked points, node buckets
  ParkedPoint* parked = nullptr;
  uint32_t* node_count = nullptr;
  uint32_t n_nodes = 0;
  BucketDev nbk{};
  std::vector<void*> ptrs;
  void release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
    capacity = 0;
    lo = LocateOut{};
    x_ir = nullptr;
    tau = nullptr;
    key = rank = cell_count = nullptr;
    bk = BucketDev{};
    parked = nullptr;
    node_count = nullptr;
    nbk = BucketDev{};
    n_nodes = 0;
    n_buckets = chunk = sub = 0;
  }
  template <class T>
  cudaError_t get(T** p, size_t count) {
    void* v = nullptr;
    cudaError_t e = cudaMalloc(&v, count * sizeof(T));
    if (e == cudaSuccess) ptrs.push_back(v);
    *p = static_cast<T*>(v);
    return e;
  }
  uint32_t n_atoms_cap = 0;
  cudaError_t ensure(size_t n, uint32_t nb, uint32_t ch, uint32_t n_atoms, uint32_t nsub, uint32_t nnodes) {
    if (n <= capacity && nb == n_buckets && ch == chunk && n_atoms == n_atoms_cap && nsub == sub && nnodes == n_nodes) return cudaSuccess;
    n_atoms_cap = n_atoms;
    if (n < capacity) n = capacity;
    release();
    cudaError_t e;
#define WS(field, type, per) if ((e = get<type>(&field, n * (per))) != cudaSuccess) return e;
    WS(lo.q_ir, double, 3)
    WS(x_ir, double, 3)
    WS(tau, int32_t, 3)
    WS(lo.ridx, int32_t, 1)
    WS(lo.invridx, int32_t, 1)
    WS(lo.cell, uint32_t, 1)
    WS(lo.tet, int32_t, 1)
    WS(lo.n_vert, int32_t, 1)
    WS(lo.vertex, uint32_t, 8)
    WS(lo.weight, double, REC_DOUBLES)
    WS(lo.slots, uint64_t, 1)
    WS(lo.status, uint32_t, 1)
    WS(key, uint32_t, 1)
    WS(rank, uint32_t, 1)
#undef WS
    if ((e = get<uint32_t>(&bk.order, n)) != cudaSuccess) return e;
    if ((e = get<uint32_t>(&cell_count, (size_t)nb * nsub)) != cudaSuccess) return e;
    if ((e = get<uint32_t>(&bk.cell_offset, (size_t)nb * nsub + 1)) != cudaSuccess) return e;
    if ((e = get<uint32_t>(&bk.cell_total, nb)) != cudaSuccess) return e;
    if ((e = get<uint32_t>(&bk.cell_start, nb)) != cudaSuccess) return e;
    if ((e = get<uint32_t>(&bk.item_start, nb)) != cudaSuccess) return e;
    if ((e = get<uint32_t>(&bk.n_items, 4)) != cudaSuccess) return e;
    if ((e = get<CellItem>(&bk.items, (n + ch - 1) / ch + nb)) != cudaSuccess) return e;
    bk.max_items = (uint32_t)((n + ch - 1) / ch + nb);
    if (nnodes) {  // node buckets of the split trellis location: bucket nnodes = "no node"
      const uint32_t nnb = nnodes + 1;
      if ((e = get<ParkedPoint>(&parked, n)) != cudaSuccess) return e;
      if ((e = get<uint32_t>(&node_count, nnb)) != cudaSuccess) return e;
      if ((e = get<uint32_t>(&nbk.cell_offset, (size_t)nnb + 1)) != cudaSuccess) return e;
      if ((e = get<uint32_t>(&nbk.cell_total, nnb)) != cudaSuccess) return e;
      if ((e = get<uint32_t>(&nbk.cell_start, nnb)) != cudaSuccess) return e;
      if ((e = get<uint32_t>(&nbk.item_start, nnb)) != cudaSuccess) return e;
      nbk.max_items = 0;
      if ((e = get<uint32_t>(&nbk.n_items, 4)) != cudaSuccess) return e;
      if ((e = get<uint32_t>(&nbk.order, n)) != cudaSuccess) return e;
      nbk.n_buckets = nnb;
      nbk.sub = 1;
      nbk.chunk = 1u << 30;  // at most one (unused) work item per node
      nbk.cell_count = node_count;
    }
    n_nodes = nnodes;
    bk.n_buckets = nb;
    bk.sub = nsub;
    sub = nsub;
    bk.chunk = ch;
    bk.cell_count = cell_count;
    n_buckets = nb;
    chunk = ch;
    capacity = n;
    return cudaSuccess;
  }
};

constexpr int N_FAIL = 4;  // per-call counters: outside BZ, outside wedge, not found, rows beyond the compact scratch of the fused consumer

// fused structure-factor finish (enqueue): where |F|^2 goes and the compact scratch for the eigenvectors of the few points that
// take the general interpolation kernel
struct SfFuse {
  double* dsf;
  double* gscratch;
  uint32_t gcap;  // rows of gscratch
};
constexpr int RC_NOT_FUSABLE = 1;  // enqueue: the pipelined cell kernel is not available for this call; nothing was launched

struct HostStage {  // device-side staging buffers of one pipeline slot of the host-pointer API
  size_t capacity = 0;
  double* dQ = nullptr;
  double* dvals = nullptr;
  double* dvecs = nullptr;
  double* dsf = nullptr;  // structure factor rows of the chunk (b200_ir_structure_factor)
  double* dgs = nullptr;  // fused consumer: compact eigenvector rows of the points that take the general kernel
  uint32_t gcap = 0;
  size_t vecs_capacity = 0;  // points dvecs has room for (the fused consumer runs larger chunks without it)
  // pageable destinations: the chunk's results are copied D2H into page-locked bounce buffers and from there, by several host
  // threads, into the caller's arrays
  char* hb = nullptr;        // [vals | vecs or sf] of one chunk
  size_t hb_bytes = 0;
  size_t pend_lo = 0, pend_n = 0;  // rows of the chunk waiting in hb
  Workspace ws;
  unsigned long long* d_fail = nullptr;  // N_FAIL counters
  unsigned long long* h_fail = nullptr;  // pinned mirror (keeps the D2H of the counters asynchronous)
  cudaStream_t stream = nullptr;
};

struct PowderWS {  // device buffers of the powder-average entry points (sized in points of the current data: dropped with it)
  size_t cap = 0;    // points per chunk
  double* dQ = nullptr;
  double* dvals = nullptr;
  double* dsf = nullptr;
  double* dvecs = nullptr;  // only when the reduction cannot be fused
  double* dgs = nullptr;    // fused: compact rows of the general-kernel points
  uint32_t gcap = 0;
  double* dhist = nullptr;  // n_qbins * n_wbins + n_qbins
  size_t hist_elems = 0;
  void release() {
    for (double** p : {&dQ, &dvals, &dsf, &dvecs, &dgs, &dhist}) {
      if (*p) cudaFree(*p);
      *p = nullptr;
    }
    cap = 0;
    gcap = 0;
    hist_elems = 0;
  }
};

struct b200_grid {
  int device = 0, sm_count = 148, kind = 0;
  BZDev h_bz;
  BZDev* d_bz = nullptr;
  double eps_w = 0, eps_o = 0;
  GridDev gd{};
  uint32_t n_vertices = 0;
  DevPool structure_pool, data_pool;
  DataDev dd{};
  bool has_data = false;
  int values_rot_error = 0, vectors_rot_error = 0;  // B200_E_UNSUPPORTED raised at ir_interpolate_at time
  size_t vals_row_bytes = 0, vecs_row_bytes = 0;
  Workspace ws;                 // device-pointer API
  unsigned long long* d_fail = nullptr;
  HostStage stage[2];
  size_t host_chunk = 0;   // max points per chunk of the host-buffer pipeline (0 = sized from free memory)
  int interp_path = 0;     // 0 auto, 1 general kernel only, 2 cell-batched kernel whenever eligible
  uint32_t chunk = 256;    // points per CTA item of the cell-batched kernel
  int split_locate = 1;    // trellis: two-kernel location with the points regrouped by node in between
  size_t bin_points = 300;  // nest / mesh: the regrouping bins are coarsened until they hold about this many points of the call
  int coop_locate = 1;     // trellis: second location kernel in its warp-cooperative form (node records staged in shared memory)
  int tile = 4;            // points per register tile of the pipelined cell kernel (4: 2 CTAs/SM, 2: 3 CTAs/SM)
  int cell_kernel = 0;     // 0 auto (pipelined kernel when its cell table fits), 1 on-the-fly staging kernel, 2 pipelined only
  unsigned char* cell_table = nullptr;  // pre-aligned per-cell records (cellinterp_tma.cu), built lazily per fill
  CellTableDev ct{};
  bool cell_table_refused = false;      // did not fit: do not try again until the data changes
  SortWorkspace sort_ws;                // device work space of b200_grid_sort_pairs
  // device-resident consumer (consumer.cu): configuration, and the eigenvector scratch of the device-buffer entry point
  DevPool sf_pool;
  SFDev sf{};
  bool has_sf = false;
  void* sf_scratch = nullptr;
  size_t sf_scratch_bytes = 0;
  int bounce = 1;               // host-buffer calls: pageable destinations through page-locked bounce buffers + host threads
  int replay_stores = 0;        // diagnostic: run k_store_replay after the pipelined cell kernel and time it ("replay")
  int sf_fused = 1;             // 1: reduce inside the pipelined cell kernel whenever possible (the eigenvectors never reach HBM)
  double* sf_gscratch = nullptr;  // device-buffer entry point: compact rows of the general-kernel points
  uint32_t sf_gcap = 0;
  PowderWS pw;
  uint64_t launches = 0;
  uint32_t last_path = 0;  // B200_PATH_* of the last enqueue
  bool timing = false;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::map<std::string, std::pair<double, int>> kernel_ms;  // name -> (sum ms, count) of the last device call
};

// Device / bounce buffers of the host pipeline and the consumer scratches are sized in POINTS times the row sizes of the data that
// was filled when they were allocated: they are dropped whenever the data (and with it the row sizes) changes.
static void drop_row_sized_buffers(b200_grid* g) {
  for (int s = 0; s < 2; ++s) {
    HostStage& h = g->stage[s];
    for (double** p : {&h.dQ, &h.dvals, &h.dvecs, &h.dsf, &h.dgs}) {
      if (*p) cudaFree(*p);
      *p = nullptr;
    }
    if (h.hb) cudaFreeHost(h.hb);
    h.hb = nullptr;
    h.hb_bytes = 0;
    h.capacity = h.vecs_capacity = 0;
    h.gcap = 0;
    h.pend_lo = h.pend_n = 0;
  }
  if (g->sf_gscratch) cudaFree(g->sf_gscratch);
  g->sf_gscratch = nullptr;
  g->sf_gcap = 0;
  if (g->sf_scratch) cudaFree(g->sf_scratch);
  g->sf_scratch = nullptr;
  g->sf_scratch_bytes = 0;
  g->pw.release();
}

static void drop_cell_table(b200_grid* g) {
  if (g->cell_table) cudaFree(g->cell_table);
  g->cell_table = nullptr;
  g->ct = CellTableDev{};
  g->cell_table_refused = false;
}

// ----------------------------------------------------------------------------------------------------
// table derivation
// ----------------------------------------------------------------------------------------------------
static void tol_pair(double tol, int digit, double* rel, double* abs_) {
  *rel = DBL_EPSILON * (double)digit * 10000.0;  // approx_float.hpp:78-100, TOL_MULT = 10000
  *abs_ = 5.0 / 1000000000000000.0;
  if (tol > *rel) *rel = tol;
  if (tol > *abs_) *abs_ = tol;
}
static void matvec_h(double* c, const double* A, const double* b) {
  for (int i = 0; i < 3; ++i) {
    c[i] = 0.0;
    for (int k = 0; k < 3; ++k) c[i] += A[i * 3 + k] * b[k];
  }
}
// covector m_f with det_f(q) = sum_i (a_i - q_i) m_i == V* (a-q).((b-a)x(c-a)); returns the rounding bound
static double face_covector(const double* a, const double* b, const double* c, double vol, double* m) {
  double u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
  double v[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
  m[0] = vol * (u[1] * v[2] - u[2] * v[1]);
  m[1] = vol * (u[2] * v[0] - u[0] * v[2]);
  m[2] = vol * (u[0] * v[1] - u[1] * v[0]);
  double bound = 0.0;
  for (int i = 0; i < 3; ++i) bound += (std::fabs(a[i]) + 4.0) * std::fabs(m[i]);
  return 256.0 * DBL_EPSILON * bound;
}

static int build_bz(const b200_bz_tables_t* t, BZDev* d, double* eps_w, double* eps_o) {
  if (t->n_faces < 4 || t->n_faces > MAX_FACES) return fail(B200_E_INVALID, "n_faces out of range");
  if (t->n_ops < 1 || t->n_ops > MAX_OPS) return fail(B200_E_INVALID, "n_ops out of range");
  if (t->n_wedge < 0 || t->n_wedge > MAX_WEDGE) return fail(B200_E_INVALID, "n_wedge out of range");
  std::memset(d, 0, sizeof(BZDev));
  tol_pair(t->float_tolerance, t->approx_tolerance, &d->cfg_rel, &d->cfg_abs);
  tol_pair(0.0, 1, &d->def_rel, &d->def_abs);
  d->transform_needed = t->transform_needed;
  d->n_faces = t->n_faces;
  d->n_wedge = t->n_wedge;
  d->n_ops = t->n_ops;
  d->no_ir_mirroring = t->no_ir_mirroring;
  d->identity_index = t->identity_index;
  for (int i = 0; i < 9; ++i) {
    d->P6t[i] = (double)t->P6t[i];
    d->invPt[i] = (double)t->invPt[i];
    d->invPt_i[i] = t->invPt[i];
    d->w_recip_metric[i] = t->w_recip_metric[i];
    d->w_real_metric[i] = t->w_real_metric[i];
    d->o_recip_metric[i] = t->o_recip_metric[i];
    d->o_real_metric[i] = t->o_real_metric[i];
    d->to_xyz[i] = t->to_xyz[i];
  }
  d->w_recip_volume = t->w_recip_volume;
  d->o_recip_volume = t->o_recip_volume;
  *eps_w = *eps_o = 0.0;
  for (int f = 0; f < t->n_faces; ++f) {
    for (int i = 0; i < 3; ++i) {
      d->pa[f][i] = t->pa[3 * f + i];
      d->pb[f][i] = t->pb[3 * f + i];
      d->pc[f][i] = t->pc[3 * f + i];
      d->ca[f][i] = t->ca[3 * f + i];
      d->cb[f][i] = t->cb[3 * f + i];
      d->cc[f][i] = t->cc[3 * f + i];
      d->normals[f][i] = t->normals[3 * f + i];
      d->taus[f][i] = t->taus[3 * f + i];
    }
    d->tau_lens[f] = t->tau_lens[f];
    *eps_w = std::max(*eps_w, face_covector(d->pa[f], d->pb[f], d->pc[f], t->w_recip_volume, d->pm[f]));
    *eps_o = std::max(*eps_o, face_covector(d->ca[f], d->cb[f], d->cc[f], t->o_recip_volume, d->cm[f]));
  }
  for (int k = 0; k < t->n_wedge; ++k) matvec_h(d->gw[k], t->o_recip_metric, t->wedge_normals + 3 * k);
  for (int j = 0; j < t->n_ops; ++j) {
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) d->Rt[j][a * 3 + b] = (double)t->rotations[9 * j + b * 3 + a];
    d->inverse_index[j] = t->inverse_index[j];
  }
  // certified fast wedge test (locate.cu): covectors R_j (G* n_k) and the rounding bound of their dot with q
  d->wedge_fast = (t->n_wedge > 0 && t->n_wedge <= MAX_WEDGE_FAST && t->no_ir_mirroring) ? 1 : 0;
  d->eps_wedge = 0.0;
  if (d->wedge_fast)
    for (int j = 0; j < t->n_ops; ++j)
      for (int k = 0; k < t->n_wedge; ++k) {
        double bound = 0.0;
        for (int a = 0; a < 3; ++a) {
          double s = 0.0;
          for (int b = 0; b < 3; ++b) s += (double)t->rotations[9 * j + a * 3 + b] * d->gw[k][b];
          d->wc[j][k][a] = s;
          bound += std::fabs(s) * 4.0;  // |q_i| <= 4 rlu inside any first Brillouin zone
        }
        d->eps_wedge = std::max(d->eps_wedge, 64.0 * DBL_EPSILON * bound);
      }
  for (int f = 0; f < t->n_faces; ++f) d->inv_tau_lens[f] = 1.0 / t->tau_lens[f];
  // sign-pattern lookup of the wedge operation
  d->n_wplanes = 0;
  if (d->wedge_fast) {
    int pid[MAX_OPS][MAX_WEDGE_FAST], sgn[MAX_OPS][MAX_WEDGE_FAST];
    int m = 0;
    bool ok = true;
    for (int j = 0; j < t->n_ops && ok; ++j)
      for (int k = 0; k < t->n_wedge && ok; ++k) {
        const double* c = d->wc[j][k];
        const double nrm = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        if (nrm == 0.0) { ok = false; break; }
        int found = -1, sg = 1;
        for (int p = 0; p < m && found < 0; ++p)
          for (int s = 1; s >= -1; s -= 2) {
            double diff = 0.0;
            for (int a = 0; a < 3; ++a) diff = std::max(diff, std::fabs(c[a] - s * d->wplane[p][a]));
            if (diff <= 1e-9 * nrm) { found = p; sg = s; break; }
          }
        if (found < 0) {
          if (m == MAX_WPLANES) { ok = false; break; }
          for (int a = 0; a < 3; ++a) d->wplane[m][a] = c[a];
          found = m++;
          sg = 1;
        }
        pid[j][k] = found;
        sgn[j][k] = sg;
      }
    if (ok) {
      auto matches = [&](int j, unsigned s) {
        for (int k = 0; k < t->n_wedge; ++k)
          if ((((s >> pid[j][k]) & 1u) != 0) != (sgn[j][k] > 0)) return false;
        return true;
      };
      for (unsigned s = 0; s < (1u << m); ++s) {
        int pick = 0xff;
        if (matches(t->identity_index, s)) pick = t->identity_index;  // bz_move.cpp:263-266 tests q itself first
        for (int j = 0; j < t->n_ops && pick == 0xff; ++j)
          if (matches(j, s)) pick = j;
        d->wtable[s] = (uint8_t)pick;
      }
      d->n_wplanes = m;
      // anything closer to a plane than this is decided by the reference arithmetic (1e-9 relative covers the plane merging)
      d->wband = d->cfg_abs * (1.0 + 4.0 * d->cfg_rel) + d->eps_wedge * 1e3 + 1e-9 * d->cfg_abs;
    }
  }
  return B200_OK;
}

static int build_trellis(b200_grid* g, const b200_trellis_tables_t* t) {
  TrellisDev& d = g->gd.tr;
  DevPool& pool = g->structure_pool;
  std::vector<double> knots;
  for (int i = 0; i < 3; ++i) {
    if (t->n_knots[i] < 2) return fail(B200_E_INVALID, "trellis needs at least two knots per axis");
    d.n_knots[i] = t->n_knots[i];
    d.knot_offset[i] = (int)knots.size();
    knots.insert(knots.end(), t->knots[i], t->knots[i] + t->n_knots[i]);
    const double span = t->knots[i][t->n_knots[i] - 1] - t->knots[i][0];
    d.knot0[i] = t->knots[i][0];
    d.knot_inv[i] = span > 0.0 ? (t->n_knots[i] - 1) / span : 0.0;
  }
  if (knots.size() > (size_t)MAX_KNOTS * 3) return fail(B200_E_INVALID, "too many trellis knots");
  if ((size_t)(t->n_knots[0] - 1) * (t->n_knots[1] - 1) * (t->n_knots[2] - 1) != t->n_nodes)
    return fail(B200_E_INVALID, "n_nodes does not match the knot vectors");
  d.n_nodes = t->n_nodes;
  d.n_cubes = t->n_cubes;
  d.n_tets = t->n_tets;
  d.n_vertices = t->n_vertices;
  // packed per-cell geometry: the locate kernel reads one contiguous record per cube / tetrahedron
  std::vector<double> cube_pack((size_t)t->n_cubes * 24), tet_pack((size_t)t->n_tets * TET_PACK, 0.0);
  for (size_t c = 0; c < t->n_cubes; ++c)
    for (int i = 0; i < 8; ++i) {
      uint32_t v = t->cube_vertices[8 * c + i];
      if (v >= t->n_vertices) return fail(B200_E_INVALID, "cube vertex index out of range");
      for (int k = 0; k < 3; ++k) cube_pack[24 * c + 3 * i + k] = t->vertices[3 * (size_t)v + k];
    }
  for (size_t k = 0; k < t->n_tets; ++k) {
    double* p = &tet_pack[k * TET_PACK];
    const double* ci = t->tet_circum + 4 * k;
    p[0] = ci[0]; p[1] = ci[1]; p[2] = ci[2];
    p[3] = ci[3] * ci[3];  // r^2 exactly as trellis_node.hpp:359
    for (int j = 0; j < 4; ++j) {
      uint32_t v = t->tet_vertices[4 * k + j];
      if (v >= t->n_vertices) return fail(B200_E_INVALID, "tetrahedron vertex index out of range");
      for (int c = 0; c < 3; ++c) p[4 + 3 * j + c] = t->vertices[3 * (size_t)v + c];
    }
    p[16] = t->tet_volume[k] * 6.0;  // trellis_node.hpp:330
  }
  CU(pool.upload(knots.data(), knots.size(), &d.knots));
  CU(pool.upload(t->node_type, t->n_nodes, &d.node_type));
  CU(pool.upload(t->node_index, t->n_nodes, &d.node_index));
  CU(pool.upload(t->cube_vertices, (size_t)t->n_cubes * 8, &d.cube_vertices));
  CU(pool.upload(cube_pack.data(), cube_pack.size(), &d.cube_pack));
  CU(pool.upload(t->poly_offsets, (size_t)t->n_polys + 1, &d.poly_offsets));
  CU(pool.upload(t->tet_vertices, (size_t)t->n_tets * 4, &d.tet_vertices));
  CU(pool.upload(tet_pack.data(), tet_pack.size(), &d.tet_pack));
  g->n_vertices = t->n_vertices;
  g->gd.cells.n_cubes = t->n_cubes;
  g->gd.cells.n_tets = t->n_tets;
  g->gd.cells.cube_vertices = d.cube_vertices;
  g->gd.cells.tet_vertices = d.tet_vertices;
  g->gd.cells.node_index = d.node_index;
  return B200_OK;
}

// pack one tetrahedron: cx cy cz c3 | v0 v1 v2 v3 | vol6 | pad     (c3 = r^2 for trellis/nest, r for mesh)
static void pack_tet(double* p, const double* c, double c3, const double* verts, const uint32_t* vi, double vol6) {
  p[0] = c[0]; p[1] = c[1]; p[2] = c[2]; p[3] = c3;
  for (int j = 0; j < 4; ++j)
    for (int k = 0; k < 3; ++k) p[4 + 3 * j + k] = verts[3 * (size_t)vi[j] + k];
  p[16] = vol6;
  p[17] = 0.0;
}

// uniform bins over the bounding box of the vertices (Nest / Mesh): the buckets of the two-kernel location
static void build_bins(b200_grid* g, const double* vertices, size_t n_vertices) {
  BinDev& b = g->gd.bins;
  b = BinDev{};
  if (!n_vertices) return;
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (size_t v = 0; v < n_vertices; ++v)
    for (int d = 0; d < 3; ++d) {
      lo[d] = std::min(lo[d], vertices[3 * v + d]);
      hi[d] = std::max(hi[d], vertices[3 * v + d]);
    }
  b.total = 1;
  for (int d = 0; d < 3; ++d) {
    const double ext = hi[d] - lo[d];
    b.n[d] = ext > 0.0 ? 64 : 1;
    b.lo[d] = lo[d];
    b.inv[d] = ext > 0.0 ? b.n[d] / ext : 0.0;
    b.total *= (uint32_t)b.n[d];
  }
}

// Candidate lists of the Nest / Mesh fast path: for every finest bin the tetrahedra `ids` (rows of `pack`, ascending) that meet the
// bin, widened by 1e-6 of the grid's extent (bounding boxes first, then the four face planes: a tetrahedron fills about a sixth
// of its box -- C3 on a Nest: 16 candidates per point by boxes alone, a quarter of that with the planes).
static int build_bin_candidates(b200_grid* g, const std::vector<double>& pack, const std::vector<uint32_t>& ids, uint32_t id_base) {
  BinDev& b = g->gd.bins;
  b.cand_offset = nullptr;
  b.cand_index = nullptr;
  if (!b.total || ids.empty()) return B200_OK;
  std::vector<uint32_t> count((size_t)b.total + 1, 0u);
  auto range = [&](uint32_t id, int lo[3], int hi[3]) {
    const double* p = &pack[(size_t)id * TET_PACK];
    for (int d = 0; d < 3; ++d) {
      double mn = p[4 + d], mx = p[4 + d];
      for (int j = 1; j < 4; ++j) {
        mn = std::min(mn, p[4 + 3 * j + d]);
        mx = std::max(mx, p[4 + 3 * j + d]);
      }
      const double pad = b.inv[d] > 0.0 ? 1e-6 * b.n[d] / b.inv[d] : 0.0;
      const double flo = (mn - pad - b.lo[d]) * b.inv[d], fhi = (mx + pad - b.lo[d]) * b.inv[d];
      lo[d] = std::max(0, std::min(b.n[d] - 1, (int)std::floor(flo)));
      hi[d] = std::max(0, std::min(b.n[d] - 1, (int)std::floor(fhi)));
    }
  };
  // a bin inside the bounding box of a tetrahedron but wholly beyond one of its face planes does not meet it (the box corner
  // nearest to the inside of the face decides); conservative: the edge-edge separating axes are not tested
  struct Planes { double n[4][3], d[4]; };
  auto planes_of = [&](uint32_t id) {
    const double* p = &pack[(size_t)id * TET_PACK];
    Planes P;
    for (int f = 0; f < 4; ++f) {  // face f is opposite vertex f
      const double* a = p + 4 + 3 * ((f + 1) & 3);
      const double* bb = p + 4 + 3 * ((f + 2) & 3);
      const double* c = p + 4 + 3 * ((f + 3) & 3);
      const double* o = p + 4 + 3 * f;
      const double u[3] = {bb[0] - a[0], bb[1] - a[1], bb[2] - a[2]}, v[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
      double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
      const double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      double side = n[0] * (o[0] - a[0]) + n[1] * (o[1] - a[1]) + n[2] * (o[2] - a[2]);
      const double sgn = side > 0 ? -1.0 : 1.0;  // outward: away from the opposite vertex
      for (int k = 0; k < 3; ++k) P.n[f][k] = len > 0 ? sgn * n[k] / len : 0.0;
      P.d[f] = P.n[f][0] * a[0] + P.n[f][1] * a[1] + P.n[f][2] * a[2];
    }
    return P;
  };
  double cell[3], padlen = 0.0;
  for (int d = 0; d < 3; ++d) {
    cell[d] = b.inv[d] > 0.0 ? 1.0 / b.inv[d] : 0.0;
    padlen = std::max(padlen, 1e-6 * b.n[d] * cell[d]);
  }
  auto meets = [&](const Planes& P, int i, int j, int k) {
    const double lo3[3] = {b.lo[0] + i * cell[0], b.lo[1] + j * cell[1], b.lo[2] + k * cell[2]};
    for (int f = 0; f < 4; ++f) {
      double m = -P.d[f];
      for (int d = 0; d < 3; ++d) m += P.n[f][d] * (P.n[f][d] > 0 ? lo3[d] : lo3[d] + cell[d]);  // corner with the smallest n.p
      if (m > padlen) return false;
    }
    return true;
  };
  for (int pass = 0; pass < 2; ++pass) {
    std::vector<uint32_t> fill;
    std::vector<uint32_t> index;
    if (pass == 1) {
      uint32_t run = 0;  // counts -> offsets
      for (size_t i = 0; i <= b.total; ++i) {
        const uint32_t c = count[i];
        count[i] = run;
        run += c;
      }
      fill.assign(count.begin(), count.end() - 1);
      index.resize(count[b.total]);
    }
    for (uint32_t id : ids) {
      int lo[3], hi[3];
      range(id, lo, hi);
      const Planes P = planes_of(id);
      for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
          for (int i = lo[0]; i <= hi[0]; ++i) {
            if (!meets(P, i, j, k)) continue;
            const size_t bin = (size_t)i + (size_t)b.n[0] * ((size_t)j + (size_t)b.n[1] * k);
            if (pass == 0) ++count[bin];
            else index[fill[bin]++] = id - id_base;
          }
    }
    if (pass == 1) {
      CU(g->structure_pool.upload(count.data(), count.size(), &b.cand_offset));
      CU(g->structure_pool.upload(index.data(), index.size(), &b.cand_index));
    }
  }
  return B200_OK;
}

static int build_nest(b200_grid* g, const b200_nest_tables_t* t) {
  NestDev& d = g->gd.ne;
  DevPool& pool = g->structure_pool;
  if (t->n_nodes < 1) return fail(B200_E_INVALID, "empty nest");
  std::vector<double> pack((size_t)t->n_nodes * TET_PACK, 0.0);
  for (size_t n = 1; n < t->n_nodes; ++n) {
    const uint32_t* vi = t->node_vertices + 4 * n;
    for (int j = 0; j < 4; ++j)
      if (vi[j] >= t->n_vertices) return fail(B200_E_INVALID, "nest vertex index out of range");
    const double* cr = t->node_circum + 4 * n;
    pack_tet(&pack[n * TET_PACK], cr, cr[3] * cr[3], t->vertices, vi, t->node_volume[n] * 6.0);  // nest.hpp:81,119
  }
  for (size_t n = 0; n < t->n_nodes; ++n)
    if (t->child_begin[n] > t->child_end[n] || t->child_end[n] > t->n_nodes) return fail(B200_E_INVALID, "nest child range out of bounds");
  {  // depth of the tree: the slow path of the location walks it with an explicit stack
    std::vector<uint32_t> depth(t->n_nodes, 0u);
    uint32_t deepest = 0;
    for (size_t n = 0; n < t->n_nodes; ++n)
      for (uint32_t c = t->child_begin[n]; c < t->child_end[n]; ++c) {
        if (c <= n) return fail(B200_E_INVALID, "nest nodes are not in breadth-first order");
        depth[c] = depth[n] + 1;
        deepest = std::max(deepest, depth[c]);
      }
    if (deepest >= (uint32_t)NEST_MAX_DEPTH) return fail(B200_E_UNSUPPORTED, "nest deeper than " + std::to_string(NEST_MAX_DEPTH) + " levels");
  }
  d.n_nodes = t->n_nodes;
  d.n_vertices = t->n_vertices;
  tol_pair(t->tolerance, t->digit, &d.rel, &d.abs_);
  CU(pool.upload(pack.data(), pack.size(), &d.node_pack));
  CU(pool.upload(t->node_vertices, (size_t)t->n_nodes * 4, &d.node_vertices));
  CU(pool.upload(t->node_is_leaf, (size_t)t->n_nodes, &d.node_is_leaf));
  CU(pool.upload(t->child_begin, (size_t)t->n_nodes, &d.child_begin));
  CU(pool.upload(t->child_end, (size_t)t->n_nodes, &d.child_end));
  g->n_vertices = t->n_vertices;
  g->gd.cells.n_cubes = 0;
  build_bins(g, t->vertices, t->n_vertices);
  {
    std::vector<uint32_t> leaves;
    for (uint32_t n = 1; n < t->n_nodes; ++n)
      if (t->node_is_leaf[n]) leaves.push_back(n);
    int rc = build_bin_candidates(g, pack, leaves, 0u);
    if (rc) return rc;
  }
  g->gd.cells.n_tets = t->n_nodes;  // bucket key = node index (only leaves ever receive points)
  g->gd.cells.tet_vertices = d.node_vertices;
  return B200_OK;
}

static int build_mesh(b200_grid* g, const b200_mesh_tables_t* t) {
  MeshDev& d = g->gd.me;
  DevPool& pool = g->structure_pool;
  if (t->n_layers < 1) return fail(B200_E_INVALID, "Can not locate without triangulation");
  const uint32_t L = t->n_layers, ntot = t->tet_offset[L];
  std::vector<double> pack((size_t)ntot * TET_PACK, 0.0);
  for (uint32_t l = 0; l < L; ++l) {
    const uint32_t nvl = t->vert_offset[l + 1] - t->vert_offset[l];
    const double* verts = t->vertices + 3 * (size_t)t->vert_offset[l];
    for (uint32_t k = t->tet_offset[l]; k < t->tet_offset[l + 1]; ++k) {
      const uint32_t* vi = t->tets + 4 * (size_t)k;
      for (int j = 0; j < 4; ++j)
        if (vi[j] >= nvl) return fail(B200_E_INVALID, "mesh vertex index out of range");
      pack_tet(&pack[(size_t)k * TET_PACK], t->centres + 3 * (size_t)k, t->radii[k], verts, vi, t->vol6[k]);
    }
  }
  const uint32_t nconn = L > 1 ? t->tet_offset[L - 1] : 0;
  for (uint32_t l = 0; l + 1 < L; ++l)
    for (uint32_t k = t->tet_offset[l]; k < t->tet_offset[l + 1]; ++k)
      for (uint32_t c = t->conn_offset[k]; c < t->conn_offset[k + 1]; ++c)
        if (t->conn_index[c] >= t->tet_offset[l + 2] - t->tet_offset[l + 1]) return fail(B200_E_INVALID, "mesh connection out of range");
  d.n_layers = L;
  d.n_tets_last = t->tet_offset[L] - t->tet_offset[L - 1];
  CU(pool.upload(t->tet_offset, (size_t)L + 1, &d.tet_offset));
  CU(pool.upload(pack.data(), pack.size(), &d.tet_pack));
  CU(pool.upload(t->tets, (size_t)ntot * 4, &d.tets));
  CU(pool.upload(t->conn_offset, (size_t)nconn + (L > 1 ? 1 : 0), &d.conn_offset));
  CU(pool.upload(t->conn_index, L > 1 ? (size_t)t->conn_offset[nconn] : 0, &d.conn_index));
  build_bins(g, t->vertices + 3 * (size_t)t->vert_offset[L - 1], t->vert_offset[L] - t->vert_offset[L - 1]);
  {  // candidates: the tetrahedra of the finest layer, as layer-local ids
    std::vector<uint32_t> fine;
    for (uint32_t k = t->tet_offset[L - 1]; k < t->tet_offset[L]; ++k) fine.push_back(k);
    int rc = build_bin_candidates(g, pack, fine, t->tet_offset[L - 1]);
    if (rc) return rc;
  }
  g->n_vertices = t->vert_offset[L] - t->vert_offset[L - 1];  // data live on the finest layer's vertices
  g->gd.cells.n_cubes = 0;
  g->gd.cells.n_tets = d.n_tets_last;
  g->gd.cells.tet_vertices = d.tets + 4 * (size_t)t->tet_offset[L - 1];
  return B200_OK;
}

// rotation dispatch of Interpolator::rotate_in_place (interpolator.hpp:386-428)
static int rot_kind_of(const b200_interp_desc_t& d, int* err) {
  *err = 0;
  const uint32_t no1 = d.elements[1] / 3u, no2 = d.elements[2] / 9u;
  int kind = -1;
  switch (d.length_unit) {
    case B200_LEN_REAL_LATTICE:
      if (d.rotates_like == B200_ROT_VECTOR) kind = 0;
      else if (d.rotates_like == B200_ROT_PSEUDOVECTOR) kind = 2;
      else if (d.rotates_like == B200_ROT_GAMMA) kind = 3;
      else *err = B200_E_UNSUPPORTED;
      break;
    case B200_LEN_RECIPROCAL_LATTICE:
      if (d.rotates_like == B200_ROT_VECTOR) kind = 1; else *err = B200_E_UNSUPPORTED;
      break;
    case B200_LEN_ANGSTROM:
      if (d.rotates_like == B200_ROT_GAMMA) kind = 4; else *err = B200_E_UNSUPPORTED;
      break;
    default:
      *err = B200_E_UNSUPPORTED;
  }
  if (*err) return -1;
  if (kind >= 3 && !d.is_complex) {  // "RotatesLike == Gamma requires complex valued data!" interpolator.hpp:616-620
    *err = B200_E_UNSUPPORTED;
    return -1;
  }
  if (no1 == 0 && no2 == 0) return -1;  // pure scalars: rip_* return false immediately
  if (kind >= 3) {
    uint32_t Nmat = (uint32_t)(std::sqrt((double)no2)) / 3u;  // interpolator_gamma.tpp:66-70
    if (no2 != 9 * Nmat * Nmat) return -1;
  }
  return kind;
}

static int fill_interp(b200_grid* g, const b200_interp_desc_t& s, uint32_t n_vertices, InterpDev* d, int* rot_err) {
  d->is_complex = s.is_complex;
  d->branches = s.branches;
  d->no0 = s.elements[0];
  if (s.elements[1] % 3) return fail(B200_E_INVALID, "Vectors must have 3N elements per branch");
  if (s.elements[2] % 9) return fail(B200_E_INVALID, "Matrices must have 9N elements per branch");
  d->no1 = s.elements[1] / 3u;
  d->no2 = s.elements[2] / 9u;
  d->span = s.elements[0] + s.elements[1] + s.elements[2];
  d->rot_kind = rot_kind_of(s, rot_err);
  size_t count = (size_t)n_vertices * s.branches * d->span * (s.is_complex ? 2 : 1);
  if (count && !s.data) return fail(B200_E_INVALID, "interpolation data pointer is NULL");
  CU(g->data_pool.upload(static_cast<const double*>(s.data), count, &d->data));
  return B200_OK;
}

// ----------------------------------------------------------------------------------------------------
// life cycle
// ----------------------------------------------------------------------------------------------------
extern "C" int b200_grid_create(int kind, const b200_bz_tables_t* bz, const void* structure, int device, b200_grid_t** out) {
  if (!bz || !structure || !out) return fail(B200_E_INVALID, "NULL argument");
  if (kind != B200_GRID_TRELLIS && kind != B200_GRID_NEST && kind != B200_GRID_MESH) return fail(B200_E_INVALID, "unknown grid kind");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(B200_E_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); brille_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(B200_E_INVALID, "device index out of range");
  CU(cudaSetDevice(device));
  b200_grid* g = new b200_grid();
  g->device = device;
  g->kind = kind;
  g->gd.kind = kind;
  cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, device);
  int rc = build_bz(bz, &g->h_bz, &g->eps_w, &g->eps_o);
  if (rc == B200_OK) {
    const BZDev* p = nullptr;
    cudaError_t ce = g->structure_pool.upload(&g->h_bz, 1, &p);
    g->d_bz = const_cast<BZDev*>(p);
    if (ce != cudaSuccess) rc = fail(B200_E_CUDA, std::string("CUDA error: ") + cudaGetErrorString(ce));
  }
  if (rc == B200_OK) {
    if (kind == B200_GRID_TRELLIS) rc = build_trellis(g, static_cast<const b200_trellis_tables_t*>(structure));
    else if (kind == B200_GRID_NEST) rc = build_nest(g, static_cast<const b200_nest_tables_t*>(structure));
    else rc = build_mesh(g, static_cast<const b200_mesh_tables_t*>(structure));
  }
  if (rc == B200_OK && cudaMalloc(&g->d_fail, N_FAIL * sizeof(unsigned long long)) != cudaSuccess)
    rc = fail(B200_E_CUDA, "cudaMalloc failed");
  for (int s = 0; s < 2 && rc == B200_OK; ++s) {
    if (cudaStreamCreateWithFlags(&g->stage[s].stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&g->stage[s].d_fail, N_FAIL * sizeof(unsigned long long)) != cudaSuccess ||
        cudaHostAlloc(&g->stage[s].h_fail, N_FAIL * sizeof(unsigned long long), cudaHostAllocDefault) != cudaSuccess)
      rc = fail(B200_E_CUDA, "stream/counter creation failed");
  }
  for (int i = 0; i < 8 && rc == B200_OK; ++i)
    if (cudaEventCreate(&g->ev[i]) != cudaSuccess) rc = fail(B200_E_CUDA, "event creation failed");
  if (rc != B200_OK) {
    b200_grid_destroy(g);
    return rc;
  }
  *out = g;
  return B200_OK;
}

extern "C" void b200_grid_destroy(b200_grid_t* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  g->structure_pool.release();
  g->data_pool.release();
  drop_cell_table(g);
  g->sort_ws.release();
  g->ws.release();
  g->sf_pool.release();
  if (g->sf_scratch) cudaFree(g->sf_scratch);
  if (g->sf_gscratch) cudaFree(g->sf_gscratch);
  g->pw.release();
  if (g->d_fail) cudaFree(g->d_fail);
  for (int s = 0; s < 2; ++s) {
    HostStage& h = g->stage[s];
    h.ws.release();
    if (h.dQ) cudaFree(h.dQ);
    if (h.dvals) cudaFree(h.dvals);
    if (h.dvecs) cudaFree(h.dvecs);
    if (h.dsf) cudaFree(h.dsf);
    if (h.dgs) cudaFree(h.dgs);
    if (h.hb) cudaFreeHost(h.hb);
    if (h.d_fail) cudaFree(h.d_fail);
    if (h.h_fail) cudaFreeHost(h.h_fail);
    if (h.stream) cudaStreamDestroy(h.stream);
  }
  for (int i = 0; i < 8; ++i)
    if (g->ev[i]) cudaEventDestroy(g->ev[i]);
  delete g;
}

extern "C" int b200_grid_set_data(b200_grid_t* g, const b200_data_tables_t* t) {
  if (!g || !t) return fail(B200_E_INVALID, "NULL argument");
  CU(cudaSetDevice(g->device));
  CU(cudaDeviceSynchronize());
  g->data_pool.release();
  drop_cell_table(g);
  drop_row_sized_buffers(g);
  g->has_data = false;
  if (t->n_vertices != g->n_vertices)
    return fail(B200_E_INVALID, "Provided " + std::to_string(t->n_vertices) + " arrays but " + std::to_string(g->n_vertices) + " were expected!");
  if (t->values.branches != t->vectors.branches)
    return fail(B200_E_INVALID, "Inconsistent values and vectors provided to DualInterpolator");  // interpolatordual.hpp:90-91
  DataDev d{};
  int rc = fill_interp(g, t->values, t->n_vertices, &d.values, &g->values_rot_error);
  if (rc) return rc;
  rc = fill_interp(g, t->vectors, t->n_vertices, &d.vectors, &g->vectors_rot_error);
  if (rc) return rc;
  const uint32_t B = t->values.branches;
  d.n_perm_rows = t->n_perm_rows;
  if (t->n_perm_rows > 1) {
    if (!t->perm_rows || !t->cube_perm || !t->tet_perm) {
      if (!t->perm_rows || (g->gd.cells.n_cubes && !t->cube_perm) || (g->gd.cells.n_tets && !t->tet_perm))
        return fail(B200_E_INVALID, "permutation rows given without the per-cell pair tables");
    }
    CU(g->data_pool.upload(t->perm_rows, (size_t)t->n_perm_rows * B, &d.perm_rows));
    CU(g->data_pool.upload(t->cube_perm, (size_t)g->gd.cells.n_cubes * 64, &d.cube_perm));
    CU(g->data_pool.upload(t->tet_perm, (size_t)g->gd.cells.n_tets * 16, &d.tet_perm));
  }
  const int G = g->h_bz.n_ops;
  d.n_ops = (uint32_t)G;
  d.n_atoms = t->n_atoms;
  const bool gamma = d.values.rot_kind >= 3 || d.vectors.rot_kind >= 3;
  if (gamma) {
    if (!t->n_atoms || !t->gamma_F0 || !t->gamma_vidx || !t->gamma_vectors)
      return fail(B200_E_INVALID, "RotatesLike::Gamma data needs the GammaTable (an atom basis)");
    auto check = [&](const InterpDev& i) {
      return i.rot_kind < 3 || (i.no1 <= t->n_atoms && (uint32_t)(std::sqrt((double)i.no2)) / 3u <= t->n_atoms);
    };
    if (!check(d.values) || !check(d.vectors)) return fail(B200_E_INVALID, "Attempting to access out of bounds mapping!");  // phonon.hpp:190-193
    for (size_t i = 0; i < (size_t)t->n_atoms * G; ++i)
      if (t->gamma_F0[i] >= t->n_atoms || t->gamma_vidx[i] >= t->n_gamma_vectors) return fail(B200_E_INVALID, "GammaTable index out of range");
    // The reference scatters the rotated vector of atom k to slot F0(k, R) of a work array of no1 slots (and matrix (n, m) to
    // slot (F0(n, R), F0(m, R^-1)) of Nmat x Nmat, interpolator_gamma.tpp:100-132): only defined when the atoms that carry
    // data are mapped among themselves.
    auto closed = [&](uint32_t count) {
      for (uint32_t k = 0; k < count; ++k)
        for (int r = 0; r < G; ++r)
          if (t->gamma_F0[(size_t)k * G + r] >= count) return false;
      return true;
    };
    for (const InterpDev* i : {&d.values, &d.vectors}) {
      if (i->rot_kind < 3) continue;
      const uint32_t Nmat = (uint32_t)(std::sqrt((double)i->no2)) / 3u;
      if (!closed(i->no1) || !closed(Nmat))
        return fail(B200_E_INVALID, "RotatesLike::Gamma data: the atoms that carry vectors / matrices are not mapped among themselves by the point group");
    }
    CU(g->data_pool.upload(t->gamma_F0, (size_t)t->n_atoms * G, &d.gamma_F0));
    CU(g->data_pool.upload(t->gamma_vidx, (size_t)t->n_atoms * G, &d.gamma_vidx));
    CU(g->data_pool.upload(t->gamma_vectors, (size_t)t->n_gamma_vectors * 3, &d.gamma_vectors));
  }
  if ((d.values.rot_kind == 4 || d.vectors.rot_kind == 4)) {
    if (!t->rot_cart) return fail(B200_E_INVALID, "LengthUnit::angstrom Gamma data needs rot_cart");
    CU(g->data_pool.upload(t->rot_cart, (size_t)G * 9, &d.rot_cart));
  }
  {  // rotations as doubles: [0,G) R, [G,2G) R^T ; determinants ; identity flags
    std::vector<double> ri((size_t)G * 18), det(G);
    std::vector<uint8_t> isid(G);
    for (int j = 0; j < G; ++j) {
      double R[9];
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) R[3 * a + b] = g->h_bz.Rt[j][3 * b + a];
      for (int i = 0; i < 9; ++i) {
        ri[9 * j + i] = R[i];
        ri[9 * (G + j) + i] = g->h_bz.Rt[j][i];
      }
      det[j] = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
      static const double E[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      isid[j] = std::memcmp(R, E, sizeof(E)) == 0;
    }
    CU(g->data_pool.upload(ri.data(), ri.size(), &d.rot_int));
    CU(g->data_pool.upload(det.data(), det.size(), &d.rot_det));
    CU(g->data_pool.upload(isid.data(), isid.size(), &d.rot_is_identity));
  }
  g->dd = d;
  g->vals_row_bytes = (size_t)B * d.values.span * (d.values.is_complex ? 16 : 8);
  g->vecs_row_bytes = (size_t)B * d.vectors.span * (d.vectors.is_complex ? 16 : 8);
  g->has_data = t->n_vertices > 0 && B > 0;
  return B200_OK;
}

// ----------------------------------------------------------------------------------------------------
// the path
// ----------------------------------------------------------------------------------------------------
static LocateIn as_input(const b200_grid* g, const LocateOut& lo) {
  LocateIn in{};
  in.q_ir = lo.q_ir; in.ridx = lo.ridx; in.invridx = lo.invridx; in.cell = lo.cell; in.tet = lo.tet;
  in.n_vert = lo.n_vert; in.vertex = lo.vertex; in.weight = lo.weight; in.slots = lo.slots; in.status = lo.status;
  in.node_type = g->gd.tr.node_type;
  in.node_index = g->gd.cells.node_index;
  return in;
}

static void note_time(b200_grid* g, const char* name, cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) == cudaSuccess) {
    auto& e = g->kernel_ms[name];
    e.first += ms;
    e.second += 1;
  }
}

// enqueue locate (+ interpolate) for n points that are already on the device; no synchronisation
static int enqueue(b200_grid* g, Workspace& ws, unsigned long long* d_fail, const double* dQ, size_t n, uint32_t mode,
                   bool interp, int ir, double* dvals, double* dvecs, cudaStream_t stream, size_t n_call, bool want_probe,
                   bool reset_fail = true, const SfFuse* fz = nullptr) {
  const uint32_t nb = g->gd.cells.n_cubes + g->gd.cells.n_tets + 1;
  // cell-batched path: worthwhile once the cells hold several points each; always correct when eligible
  bool cell = interp && g->interp_path != 1 && cell_path_eligible(g->dd) && n < 0xffffffffull;
  // (decided on the size of the whole call, not of the chunk, so that chunking never changes which kernel a point sees)
  if (cell && g->interp_path == 0 && n_call < 4 * (size_t)nb) cell = false;
  uint32_t mpp = 0, chunk = g->chunk;
  bool tma = false;
  if (cell && g->cell_kernel != 1) {
    // pipelined kernel: needs the per-cell records, built once per fill (synchronously: the two host-pipeline streams share it)
    const uint32_t ch = cell_tma_pick(g->dd, g->gd.cells.n_cubes > 0, g->chunk, (g->tile == 2 ? 72 : 108) * 1024, &mpp, fz != nullptr);
    if (ch && mpp) {
      if (g->cell_table && g->ct.mpp == mpp) {
        tma = true;
      } else if (!g->cell_table_refused) {
        if (g->cell_table) drop_cell_table(g);
        const CellTableDev ct = cell_table_layout(g->dd, g->gd.cells.n_cubes, g->gd.cells.n_tets, mpp);
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        if (ct.total_bytes <= free_b / 4 && cudaMalloc(&g->cell_table, ct.total_bytes) == cudaSuccess) {
          CU(launch_build_cell_table(g->dd, g->gd.cells.cube_vertices, g->gd.cells.tet_vertices, ct, g->cell_table, g->sm_count, stream));
          CU(cudaStreamSynchronize(stream));
          g->ct = ct;
          g->launches += 1;
          tma = true;
        } else {
          (void)cudaGetLastError();
          g->cell_table = nullptr;
          g->cell_table_refused = true;
        }
      }
      if (tma) chunk = ch;
    }
    if (!tma && g->cell_kernel == 2 && !fz) return fail(B200_E_CUDA, "the cell table of the pipelined kernel does not fit in device memory");
  }
  if (fz && !tma) return RC_NOT_FUSABLE;
  if (cell && !tma) {
    chunk = cell_pick_chunk(g->dd, g->gd.cells.n_cubes > 0, g->chunk, 100 * 1024, &mpp);
    if (chunk == 0 || mpp == 0) {
      cell = false;
      chunk = g->chunk;
    }
  }
  const uint32_t nsub = (uint32_t)g->h_bz.n_ops;
  // two-kernel trellis location (points regrouped by node between the halves): pays off once the nodes hold a few points each
  // (nest / mesh: spatial bins, coarsened to about 100 points of the call per bin: the lanes of a warp then share the candidate list of
  // the fast path; bins of 1000 points sort faster but the second kernel loses more than that; the work space is sized for the finest level)
  const bool trellis = g->gd.kind == B200_GRID_TRELLIS;
  int bin_shift = 0;
  if (!trellis && g->gd.bins.total)
    while (bin_shift < 4 && (size_t)bins_at_level(g->gd.bins, bin_shift) * g->bin_points > n_call) ++bin_shift;
  const uint32_t n_nodes_alloc = trellis ? g->gd.tr.n_nodes : g->gd.bins.total;
  const uint32_t n_nodes = trellis ? g->gd.tr.n_nodes : (g->gd.bins.total ? bins_at_level(g->gd.bins, bin_shift) : 0u);
  const bool split = cell && n_nodes && g->split_locate && !(mode & MODE_NO_LOCATE) && n_call >= 8 * (size_t)n_nodes;
  CU(ws.ensure(n, nb, chunk, 0u, nsub, split ? n_nodes_alloc : ws.n_nodes));
  if (reset_fail) CU(cudaMemsetAsync(d_fail, 0, N_FAIL * sizeof(unsigned long long), stream));
  LocateOut lo = ws.lo;
  if (want_probe) {  // x_ir and tau leave the kernel only for a caller who asked for them (36 bytes per point)
    lo.x_ir = ws.x_ir;
    lo.tau = ws.tau;
  }
  if (cell) {
    lo.key = ws.key;
    lo.rank = ws.rank;
    lo.cell_count = ws.cell_count;
    lo.sub = nsub;
    CU(cudaMemsetAsync(ws.cell_count, 0, (size_t)nb * nsub * sizeof(uint32_t), stream));
  }
  g->last_path = (split ? B200_PATH_SPLIT_LOCATE : 0u) | (interp && cell ? (tma ? B200_PATH_CELL_PIPELINED : B200_PATH_CELL_ONTHEFLY) : 0u) |
                 (interp && !cell ? B200_PATH_GENERAL : 0u) | (fz ? B200_PATH_SF_FUSED : 0u);
  if (g->timing) cudaEventRecord(g->ev[0], stream);
  if (split) {
    lo.parked = ws.parked;
    lo.lean = want_probe ? 0 : 1;
    lo.node_count = ws.node_count;
    lo.bin_shift = bin_shift;
    CU(cudaMemsetAsync(ws.node_count, 0, ((size_t)n_nodes + 1) * sizeof(uint32_t), stream));
    CU(launch_locate(g->d_bz, g->gd, dQ, n, mode | MODE_SPLIT_A, g->eps_w, g->eps_o, lo, d_fail, g->sm_count, stream));
    if (g->timing) cudaEventRecord(g->ev[6], stream);
    BucketDev nbk = ws.nbk;  // (allocated for the finest level of bins; this call uses n_nodes of them + the "no node" bucket)
    nbk.n_buckets = n_nodes + 1;
    CU(launch_bucket_sort(nbk, ws.key, ws.rank, n, g->sm_count, stream));
    if (g->timing) cudaEventRecord(g->ev[7], stream);
    CU(launch_locate_in_node(g->d_bz, g->gd, n, mode, lo, ws.nbk.order, d_fail, g->sm_count, stream, g->coop_locate != 0));
    if (g->coop_locate && trellis) g->last_path |= B200_PATH_COOP_LOCATE;
    g->launches += 6;
  } else {
    CU(launch_locate(g->d_bz, g->gd, dQ, n, mode, g->eps_w, g->eps_o, lo, d_fail, g->sm_count, stream));
    g->launches += 1;
  }
  if (g->timing) cudaEventRecord(g->ev[1], stream);
  if (interp && cell) {
    CU(launch_bucket_sort(ws.bk, ws.key, ws.rank, n, g->sm_count, stream, split ? ws.nbk.order : nullptr));
    g->launches += 5;
    if (g->timing) cudaEventRecord(g->ev[2], stream);
    CellArgs a{};
    a.dd = g->dd;
    a.cube_vertices = g->gd.cells.cube_vertices;
    a.tet_vertices = g->gd.cells.tet_vertices;
    a.n_cubes = g->gd.cells.n_cubes;
    a.bk = ws.bk;
    a.weight = lo.weight;
    a.q_ir = lo.q_ir;
    a.ridx = lo.ridx;
    a.invridx = lo.invridx;
    a.vals_out = dvals;
    a.vecs_out = dvecs;
    a.ir = ir;
    a.modes_per_pass = mpp;
    if (fz) {  // fused structure factor: |F|^2 instead of the eigenvectors
      a.Q = dQ;
      a.sf_out = fz->dsf;
      a.sf = g->sf;
      a.vecs_out = nullptr;
    }
    if (tma) CU(launch_interp_cell_tma(a, g->ct, g->cell_table, n, g->sm_count, stream, g->tile));
    else CU(launch_interp_cell(a, n, stream));
    if (tma && !fz && g->replay_stores && g->timing) {  // diagnostic: the store pattern on its own, then the real kernel once more
      cudaEventRecord(g->ev[4], stream);
      CU(launch_store_replay(a, n, g->sm_count, stream));
      cudaEventRecord(g->ev[5], stream);
      CU(launch_interp_cell_tma(a, g->ct, g->cell_table, n, g->sm_count, stream, g->tile));
      cudaEventSynchronize(g->ev[5]);
      note_time(g, "replay", g->ev[4], g->ev[5]);
    }
    // points that are not generic members of their cell (and failed points): general kernel over the last bucket
    if (fz) {  // ... into compact scratch rows, reduced by the list mode of the consumer kernel
      CU(launch_interp(g->dd, as_input(g, lo), n, ir, dvals, fz->gscratch, g->sm_count, stream, ws.bk.order, ws.bk.n_items, fz->gcap, d_fail + 3));
      CU(launch_structure_factor(g->sf, dQ, fz->gscratch, n, g->dd.vectors.branches, fz->dsf, g->sm_count, stream, ws.bk.order, ws.bk.n_items, fz->gcap));
      g->launches += 1;
    } else {
      CU(launch_interp(g->dd, as_input(g, lo), n, ir, dvals, dvecs, g->sm_count, stream, ws.bk.order, ws.bk.n_items));
    }
    g->launches += 2;
    if (g->timing) cudaEventRecord(g->ev[3], stream);
  } else if (interp) {
    if (g->timing) cudaEventRecord(g->ev[2], stream);
    CU(launch_interp(g->dd, as_input(g, lo), n, ir, dvals, dvecs, g->sm_count, stream, nullptr, nullptr));
    g->launches += 1;
    if (g->timing) cudaEventRecord(g->ev[3], stream);
  }
  if (g->timing) {
    cudaEventSynchronize(interp ? g->ev[3] : g->ev[1]);
    note_time(g, "locate", g->ev[0], g->ev[1]);
    if (split) {  // the parts of the two-kernel location
      note_time(g, "locate_a", g->ev[0], g->ev[6]);
      note_time(g, "locate_sort", g->ev[6], g->ev[7]);
      note_time(g, "locate_b", g->ev[7], g->ev[1]);
    }
    if (interp) {
      note_time(g, "sort", g->ev[1], g->ev[2]);
      note_time(g, "interpolate", g->ev[2], g->ev[3]);
    }
  }
  return B200_OK;
}

static int status_error(const unsigned long long* c, size_t nQ) {
  // order of the reference's checks: moveinto (bz_move.cpp:151-158), wedge (:288-294), then location
  if (c[0]) return fail(B200_E_OUTSIDE_BZ, "Not all points inside Brillouin zone (" + std::to_string(c[0]) + " of " + std::to_string(nQ) + " still outside)");
  if (c[1]) return fail(B200_E_OUTSIDE_WEDGE, std::to_string(c[1]) + " Q point(s) are outside of the irreducible BrillouinZone");
  if (c[2]) return fail(B200_E_NOT_FOUND, "interpolate_at failed to find " + std::to_string(c[2]) + (c[2] > 1 ? " points." : " point."));
  return B200_OK;
}

static int check_ready(b200_grid* g, bool need_data, int ir) {
  if (!g) return fail(B200_E_INVALID, "NULL grid");
  if (need_data && !g->has_data) return fail(B200_E_NODATA, "The interpolation data must be filled before interpolating.");
  if (need_data && ir) {
    if (g->values_rot_error || g->vectors_rot_error)
      return fail(B200_E_UNSUPPORTED, "LengthUnit, RotatesLike combination not implemented");
  }
  return B200_OK;
}

static int interpolate_device(b200_grid* g, const double* dQ, size_t nQ, uint32_t flags, int ir, void* dvals, void* dvecs,
                              b200_probe_t* dprobe, cudaStream_t stream, uint64_t* n_failed) {
  int rc = check_ready(g, true, ir);
  if (rc) return rc;
  // (point indices travel as 32-bit words through the sorts and the point records; the host-buffer entry points chunk on their own)
  if (nQ >= 0xfffffff0ull) return fail(B200_E_INVALID, "at most 4294967279 points per device-buffer call");
  DeviceGuard guard;
  CU(cudaSetDevice(g->device));
  if (g->timing) g->kernel_ms.clear();
  uint32_t mode = (flags & B200_FLAG_NO_MOVE) ? MODE_NO_MOVE : 0u;
  if (ir) mode |= MODE_IR;
  rc = enqueue(g, g->ws, g->d_fail, dQ, nQ, mode, true, ir, static_cast<double*>(dvals), static_cast<double*>(dvecs), stream, nQ, dprobe != nullptr);
  if (rc) return rc;
  if (dprobe) {
    const LocateOut& lo = g->ws.lo;
#define CP(dst, src, type, per) if (dprobe->dst) CU(cudaMemcpyAsync(dprobe->dst, src, nQ * (per) * sizeof(type), cudaMemcpyDeviceToDevice, stream));
    CP(q_ir, lo.q_ir, double, 3) CP(x_ir, g->ws.x_ir, double, 3) CP(tau, g->ws.tau, int32_t, 3) CP(ridx, lo.ridx, int32_t, 1)
    CP(invridx, lo.invridx, int32_t, 1) CP(cell, lo.cell, uint32_t, 1) CP(tet, lo.tet, int32_t, 1) CP(n_vert, lo.n_vert, int32_t, 1)
    CP(vertex, lo.vertex, uint32_t, 8) CP(status, lo.status, uint32_t, 1)
#undef CP
    if (dprobe->weight && nQ)  // the weight rows are the heads of the 96-byte point records
      CU(cudaMemcpy2DAsync(dprobe->weight, 8 * sizeof(double), lo.weight, REC_BYTES, 8 * sizeof(double), nQ, cudaMemcpyDeviceToDevice, stream));
  }
  if (n_failed) {
    unsigned long long c[3];
    CU(cudaMemcpyAsync(c, g->d_fail, sizeof(c), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    *n_failed = c[0] + c[1] + c[2];
    return status_error(c, nQ);
  }
  return B200_OK;
}

extern "C" int b200_ir_interpolate_at_device(b200_grid_t* g, const double* dQ, size_t nQ, uint32_t flags, void* dvals,
                                             void* dvecs, b200_probe_t* dprobe, void* stream, uint64_t* n_failed) {
  return interpolate_device(g, dQ, nQ, flags, 1, dvals, dvecs, dprobe, static_cast<cudaStream_t>(stream), n_failed);
}

// copy with several host threads (a fresh numpy array is first touched here: the page faults are spread over the cores)
static void parallel_copy(void* dst, const void* src, size_t bytes) {
  const size_t min_slice = (size_t)4 << 20;
  unsigned nt = std::thread::hardware_concurrency();
  nt = std::max(1u, std::min(nt, 16u));
  nt = (unsigned)std::min<size_t>(nt, std::max<size_t>(1, bytes / min_slice));
  if (nt <= 1) {
    std::memcpy(dst, src, bytes);
    return;
  }
  const size_t slice = ((bytes + nt - 1) / nt + 4095) / 4096 * 4096;
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) {
    const size_t lo = (size_t)t * slice;
    if (lo >= bytes) break;
    const size_t n = std::min(slice, bytes - lo);
    th.emplace_back([=] { std::memcpy(static_cast<char*>(dst) + lo, static_cast<const char*>(src) + lo, n); });
  }
  for (auto& x : th) x.join();
}
static bool is_pageable(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

// host-pointer pipeline: chunks alternate between two stages (stream + staging buffers) so that the H2D copy of
// chunk i+1 and the D2H copy of chunk i-1 overlap the kernels of chunk i.
// sf_out != nullptr: the structure-factor consumer -- fused into the pipelined cell kernel when `fuse` (the eigenvectors then
// exist nowhere; the staging buffer for them is not even allocated), otherwise reduced from the staging buffer by k_structure_factor.
static int host_pipeline_body(b200_grid* g, const double* Q, size_t nQ, uint32_t mode, bool interp, int ir, void* vals, void* vecs,
                              b200_probe_t* probe, double* sf_out, bool fuse);
// Every exit of the pipeline passes through here: on an error nothing may be left in flight -- kernels of the other stage and
// their asynchronous copies into the caller's arrays (or the bounce buffers) -- when the caller gets control back (and may free
// those arrays), and no chunk may be left waiting in a bounce buffer for the next call.
static int host_pipeline(b200_grid* g, const double* Q, size_t nQ, uint32_t mode, bool interp, int ir, void* vals, void* vecs,
                         b200_probe_t* probe, double* sf_out = nullptr, bool fuse = false) {
  const int rc = host_pipeline_body(g, Q, nQ, mode, interp, ir, vals, vecs, probe, sf_out, fuse);
  if (rc != B200_OK) {
    const std::string msg = g_err;
    for (int s = 0; s < 2; ++s) {
      if (g->stage[s].stream) cudaStreamSynchronize(g->stage[s].stream);
      g->stage[s].pend_n = 0;
    }
    (void)cudaGetLastError();
    g_err = msg;
  }
  return rc;
}
static int host_pipeline_body(b200_grid* g, const double* Q, size_t nQ, uint32_t mode, bool interp, int ir, void* vals, void* vecs,
                              b200_probe_t* probe, double* sf_out, bool fuse) {
  CU(cudaSetDevice(g->device));
  if (g->timing) g->kernel_ms.clear();
  const size_t sf_row = (size_t)g->dd.vectors.branches * sizeof(double);
  const size_t vecs_per_q = fuse ? g->vecs_row_bytes / 16 + 1 : g->vecs_row_bytes;  // fused: only the compact scratch (1/16 of the rows)
  const size_t per_q = 24 + (interp ? g->vals_row_bytes + vecs_per_q : 0) + (sf_out ? sf_row : 0) + 200;
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  size_t budget = std::min<size_t>(free_b / 4, (size_t)6 << 30);  // both stages together
  size_t chunk = std::max<size_t>(1024, budget / 2 / per_q);
  chunk = std::min(chunk, (size_t)1 << 22);
  {  // at least ~8 chunks per call, so that the copies of one chunk hide the kernels of the next (a call that is a single chunk is
     // H2D, kernels and D2H in series), but never so small that the cells hold only a handful of points per chunk
    const size_t nb = (size_t)g->gd.cells.n_cubes + g->gd.cells.n_tets + 1;
    chunk = std::min(chunk, std::max<size_t>(std::max<size_t>(32 * nb, (nQ + 7) / 8), 65536));
  }
  // Pageable destinations (plain numpy arrays): a device-to-pageable copy runs at a fraction of the PCIe rate and first-touches the
  // caller's fresh pages on one thread; instead the chunk goes to a page-locked bounce buffer at full rate and several host threads
  // move it on while the GPU works on the next chunk.
  const size_t second_row = sf_out ? sf_row : g->vecs_row_bytes;
  char* const second_dst = sf_out ? reinterpret_cast<char*>(sf_out) : static_cast<char*>(vecs);
  const bool bounce = interp && g->bounce && nQ * (g->vals_row_bytes + second_row) >= ((size_t)8 << 20) && (is_pageable(vals) || is_pageable(second_dst));
  if (bounce) chunk = std::min(chunk, std::max<size_t>(((size_t)640 << 20) / (g->vals_row_bytes + second_row), 1024));
  if (g->host_chunk) chunk = std::min(chunk, g->host_chunk);
  if (chunk > nQ) chunk = std::max<size_t>(nQ, 1);
  auto flush = [&](HostStage& h) {  // the finished chunk of this stage: bounce buffer -> caller's arrays
    if (!bounce || !h.pend_n) return;
    parallel_copy(static_cast<char*>(vals) + h.pend_lo * g->vals_row_bytes, h.hb, h.pend_n * g->vals_row_bytes);
    parallel_copy(second_dst + h.pend_lo * second_row, h.hb + h.pend_n * g->vals_row_bytes, h.pend_n * second_row);
    h.pend_n = 0;
  };
  unsigned long long total[N_FAIL] = {0, 0, 0, 0};
  bool pending[2] = {false, false};
  size_t nchunks = (nQ + chunk - 1) / chunk;
  for (size_t c = 0; c < nchunks; ++c) {
    HostStage& h = g->stage[c & 1];
    if (pending[c & 1]) {  // staging buffers of this slot are reused: wait for its previous chunk
      CU(cudaStreamSynchronize(h.stream));
      for (int k = 0; k < N_FAIL; ++k) total[k] += h.h_fail[k];
      pending[c & 1] = false;
      flush(h);
    }
    const size_t lo = c * chunk, n = std::min(chunk, nQ - lo);
    if (h.capacity < n) {
      for (double** p : {&h.dQ, &h.dvals, &h.dvecs, &h.dsf, &h.dgs}) {
        if (*p) cudaFree(*p);
        *p = nullptr;
      }
      h.capacity = 0;
      h.gcap = 0;
      h.vecs_capacity = 0;
      CU(cudaMalloc(&h.dQ, n * 3 * sizeof(double)));
      h.capacity = n;
    }
    if (interp && !h.dvals) CU(cudaMalloc(&h.dvals, std::max<size_t>(h.capacity * g->vals_row_bytes, 8)));
    if (interp && !fuse && h.vecs_capacity < n) {  // sized for this call's chunks, not for the (possibly larger) chunks of a fused call
      if (h.dvecs) cudaFree(h.dvecs);
      h.dvecs = nullptr;
      h.vecs_capacity = 0;
      CU(cudaMalloc(&h.dvecs, std::max<size_t>(std::min(chunk, h.capacity) * g->vecs_row_bytes, 8)));
      h.vecs_capacity = std::min(chunk, h.capacity);
    }
    if (sf_out && !h.dsf) CU(cudaMalloc(&h.dsf, std::max<size_t>(h.capacity * sf_row, 8)));
    if (fuse && !h.dgs) {
      h.gcap = (uint32_t)std::max<size_t>(1024, h.capacity / 16);
      CU(cudaMalloc(&h.dgs, (size_t)h.gcap * g->vecs_row_bytes));
    }
    CU(cudaMemcpyAsync(h.dQ, Q + 3 * lo, n * 3 * sizeof(double), cudaMemcpyHostToDevice, h.stream));
    SfFuse fz{h.dsf, h.dgs, h.gcap};
    int rc = enqueue(g, h.ws, h.d_fail, h.dQ, n, mode, interp, ir, h.dvals, h.dvecs, h.stream, nQ, probe != nullptr, true, fuse ? &fz : nullptr);
    if (rc == RC_NOT_FUSABLE) {  // (decided on the size of the call: always on the first chunk) start again without the fusion
      for (int st = 0; st < 2; ++st) CU(cudaStreamSynchronize(g->stage[st].stream));
      return host_pipeline(g, Q, nQ, mode, interp, ir, vals, vecs, probe, sf_out, false);
    }
    if (rc) return rc;
    if (interp) {
      char* vals_to = static_cast<char*>(vals) + lo * g->vals_row_bytes;
      char* second_to = second_dst + lo * second_row;
      if (bounce) {
        const size_t need = n * (g->vals_row_bytes + second_row);
        if (h.hb_bytes < need) {
          if (h.hb) cudaFreeHost(h.hb);
          h.hb = nullptr;
          h.hb_bytes = 0;
          const size_t want = std::min(chunk, nQ) * (g->vals_row_bytes + second_row);
          CU(cudaHostAlloc(reinterpret_cast<void**>(&h.hb), want, cudaHostAllocDefault));
          h.hb_bytes = want;
        }
        vals_to = h.hb;
        second_to = h.hb + n * g->vals_row_bytes;
        h.pend_lo = lo;
        h.pend_n = n;
      }
      CU(cudaMemcpyAsync(vals_to, h.dvals, n * g->vals_row_bytes, cudaMemcpyDeviceToHost, h.stream));
      if (sf_out) {  // only `modes` doubles per Q go back
        if (!fuse) {  // the eigenvectors of the chunk are reduced where they are
          CU(launch_structure_factor(g->sf, h.dQ, h.dvecs, n, g->dd.vectors.branches, h.dsf, g->sm_count, h.stream));
          g->launches += 1;
        }
        CU(cudaMemcpyAsync(second_to, h.dsf, n * sf_row, cudaMemcpyDeviceToHost, h.stream));
      } else {
        CU(cudaMemcpyAsync(second_to, h.dvecs, n * g->vecs_row_bytes, cudaMemcpyDeviceToHost, h.stream));
      }
    }
    if (probe) {
      const LocateOut& w = h.ws.lo;
#define CP(dst, src, type, per) if (probe->dst && src) CU(cudaMemcpyAsync(probe->dst + lo * (per), src, n * (per) * sizeof(type), cudaMemcpyDeviceToHost, h.stream));
      CP(q_ir, w.q_ir, double, 3) CP(x_ir, h.ws.x_ir, double, 3) CP(tau, h.ws.tau, int32_t, 3) CP(ridx, w.ridx, int32_t, 1)
      CP(invridx, w.invridx, int32_t, 1) CP(status, w.status, uint32_t, 1)
      if (!(mode & MODE_NO_LOCATE)) {
        CP(cell, w.cell, uint32_t, 1) CP(tet, w.tet, int32_t, 1) CP(n_vert, w.n_vert, int32_t, 1)
        CP(vertex, w.vertex, uint32_t, 8)
        if (probe->weight && n)
          CU(cudaMemcpy2DAsync(probe->weight + lo * 8, 8 * sizeof(double), w.weight, REC_BYTES, 8 * sizeof(double), n, cudaMemcpyDeviceToHost, h.stream));
      }
#undef CP
    }
    CU(cudaMemcpyAsync(h.h_fail, h.d_fail, N_FAIL * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h.stream));
    pending[c & 1] = true;
  }
  for (int s = 0; s < 2; ++s)
    if (pending[s]) {
      CU(cudaStreamSynchronize(g->stage[s].stream));
      for (int k = 0; k < N_FAIL; ++k) total[k] += g->stage[s].h_fail[k];
      flush(g->stage[s]);
    }
  int rc = status_error(total, nQ);
  if (rc == B200_OK && fuse && total[3])  // more general-kernel points than compact rows in some chunk (a degenerate point set)
    return host_pipeline(g, Q, nQ, mode, interp, ir, vals, vecs, probe, sf_out, false);
  return rc;
}

extern "C" int b200_ir_interpolate_at(b200_grid_t* g, const double* Q, size_t nQ, uint32_t flags, void* vals, void* vecs,
                                      b200_probe_t* probe) {
  int rc = check_ready(g, true, 1);
  if (rc) return rc;
  if (nQ && (!Q || !vals || !vecs)) return fail(B200_E_INVALID, "NULL buffer");
  if (nQ == 0) return B200_OK;
  uint32_t mode = MODE_IR | ((flags & B200_FLAG_NO_MOVE) ? MODE_NO_MOVE : 0u);
  return host_pipeline(g, Q, nQ, mode, true, 1, vals, vecs, probe);
}

extern "C" int b200_interpolate_at(b200_grid_t* g, const double* Q, size_t nQ, uint32_t flags, void* vals, void* vecs,
                                   b200_probe_t* probe) {
  int rc = check_ready(g, true, 0);
  if (rc) return rc;
  if (nQ && (!Q || !vals || !vecs)) return fail(B200_E_INVALID, "NULL buffer");
  if (nQ == 0) return B200_OK;
  uint32_t mode = (flags & B200_FLAG_NO_MOVE) ? MODE_NO_MOVE : 0u;
  return host_pipeline(g, Q, nQ, mode, true, 0, vals, vecs, probe);
}

extern "C" int b200_moveinto(b200_grid_t* g, const double* Q, size_t nQ, int ir, b200_probe_t* probe) {
  int rc = check_ready(g, false, 0);
  if (rc) return rc;
  if (nQ && (!Q || !probe)) return fail(B200_E_INVALID, "NULL buffer");
  if (nQ == 0) return B200_OK;
  uint32_t mode = MODE_NO_LOCATE;
  switch (ir) {
    case 0: break;                                          // moveinto
    case 1: mode |= MODE_IR; break;                         // ir_moveinto
    case 2: mode |= MODE_IR | MODE_NO_TAU; break;           // ir_moveinto_wedge
    case 3: mode |= MODE_NO_MOVE | MODE_ISINSIDE; break;    // isinside
    default: return fail(B200_E_INVALID, "b200_moveinto: ir must be 0 (moveinto), 1 (ir_moveinto), 2 (ir_moveinto_wedge) or 3 (isinside)");
  }
  return host_pipeline(g, Q, nQ, mode, false, ir, nullptr, nullptr, probe);
}

// ----------------------------------------------------------------------------------------------------
// device-resident consumer: structure factor (consumer.cu)
// ----------------------------------------------------------------------------------------------------
extern "C" int b200_grid_set_structure_factor(b200_grid_t* g, const b200_sf_config_t* c) {
  if (!g || !c) return fail(B200_E_INVALID, "NULL argument");
  if (!c->n_atoms || !c->coef) return fail(B200_E_INVALID, "structure factor: n_atoms and coef are required");
  CU(cudaSetDevice(g->device));
  CU(cudaDeviceSynchronize());
  g->sf_pool.release();
  g->has_sf = false;
  SFDev d{};
  d.n_atoms = c->n_atoms;
  CU(g->sf_pool.upload(c->coef, (size_t)c->n_atoms * 2, &d.coef));
  if (c->positions) CU(g->sf_pool.upload(c->positions, (size_t)c->n_atoms * 3, &d.pos));
  if (c->debye_waller) CU(g->sf_pool.upload(c->debye_waller, (size_t)c->n_atoms * 9, &d.dw));
  for (int i = 0; i < 9; ++i) d.T[i] = c->q_transform[i];
  d.conjugate = c->conjugate ? 1 : 0;
  g->sf = d;
  g->has_sf = true;
  return B200_OK;
}

static int check_sf_ready(b200_grid* g) {
  int rc = check_ready(g, true, 1);
  if (rc) return rc;
  if (!g->has_sf) return fail(B200_E_INVALID, "structure factor: call b200_grid_set_structure_factor first");
  const InterpDev& v = g->dd.vectors;
  if (!v.is_complex || v.no0 != 0 || v.no2 != 0 || v.no1 != g->sf.n_atoms)
    return fail(B200_E_UNSUPPORTED, "structure factor: the eigenvector data must be complex 3-vectors, one per atom (" +
                                        std::to_string(g->sf.n_atoms) + " atoms configured)");
  return B200_OK;
}

extern "C" int b200_ir_structure_factor(b200_grid_t* g, const double* Q, size_t nQ, uint32_t flags, void* vals, double* sf_out) {
  int rc = check_sf_ready(g);
  if (rc) return rc;
  if (nQ && (!Q || !vals || !sf_out)) return fail(B200_E_INVALID, "NULL buffer");
  if (nQ == 0) return B200_OK;
  uint32_t mode = MODE_IR | ((flags & B200_FLAG_NO_MOVE) ? MODE_NO_MOVE : 0u);
  return host_pipeline(g, Q, nQ, mode, true, 1, vals, nullptr, nullptr, sf_out, g->sf_fused && cell_sf_fusable(g->dd, g->sf));
}

extern "C" int b200_ir_structure_factor_device(b200_grid_t* g, const double* dQ, size_t nQ, uint32_t flags, void* dvals, double* dsf,
                                               void* dscratch, void* stream_, uint64_t* n_failed) {
  int rc = check_sf_ready(g);
  if (rc) return rc;
  if (nQ && (!dQ || !dvals || !dsf)) return fail(B200_E_INVALID, "NULL buffer");
  if (n_failed) *n_failed = 0;
  if (nQ == 0) return B200_OK;
  if (nQ >= 0xfffffff0ull) return fail(B200_E_INVALID, "at most 4294967279 points per device-buffer call");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard guard;
  CU(cudaSetDevice(g->device));
  if (g->timing) g->kernel_ms.clear();
  const uint32_t mode = MODE_IR | ((flags & B200_FLAG_NO_MOVE) ? MODE_NO_MOVE : 0u);
  // Fused into the pipelined cell kernel when the caller does not ask for the eigenvectors (no scratch of theirs) and lets the
  // call synchronise (n_failed): the rows-beyond-the-compact-scratch counter has to be read back before the result is final.
  if (g->sf_fused && !dscratch && n_failed && nQ < 0xffffffffull && cell_sf_fusable(g->dd, g->sf)) {
    const uint32_t want = (uint32_t)std::max<size_t>(1024, nQ / 16);
    if (g->sf_gcap < want) {
      CU(cudaStreamSynchronize(stream));
      if (g->sf_gscratch) cudaFree(g->sf_gscratch);
      g->sf_gscratch = nullptr;
      g->sf_gcap = 0;
      CU(cudaMalloc(&g->sf_gscratch, (size_t)want * g->vecs_row_bytes));
      g->sf_gcap = want;
    }
    const SfFuse fz{dsf, g->sf_gscratch, g->sf_gcap};
    rc = enqueue(g, g->ws, g->d_fail, dQ, nQ, mode, true, 1, static_cast<double*>(dvals), nullptr, stream, nQ, false, true, &fz);
    if (rc != RC_NOT_FUSABLE) {
      if (rc) return rc;
      unsigned long long c[N_FAIL];
      CU(cudaMemcpyAsync(c, g->d_fail, sizeof(c), cudaMemcpyDeviceToHost, stream));
      CU(cudaStreamSynchronize(stream));
      if (c[3] == 0) {
        *n_failed = c[0] + c[1] + c[2];
        return status_error(c, nQ);
      }
      // a degenerate point set (more general-kernel points than compact rows): once more, through the scratch
    }
  }
  size_t chunk = nQ;
  char* scratch = static_cast<char*>(dscratch);
  if (!scratch) {  // the library's own scratch: as many points as fit a third of the free memory (kept between calls)
    const size_t want = nQ * g->vecs_row_bytes;
    if (g->sf_scratch_bytes < want) {
      size_t free_b = 0, total_b = 0;
      CU(cudaMemGetInfo(&free_b, &total_b));
      const size_t cap = std::max<size_t>((free_b + g->sf_scratch_bytes) / 3, g->vecs_row_bytes);
      const size_t bytes = std::min(want, cap / g->vecs_row_bytes * g->vecs_row_bytes);
      if (bytes > g->sf_scratch_bytes) {
        CU(cudaStreamSynchronize(stream));
        if (g->sf_scratch) cudaFree(g->sf_scratch);
        g->sf_scratch = nullptr;
        g->sf_scratch_bytes = 0;
        CU(cudaMalloc(&g->sf_scratch, bytes));
        g->sf_scratch_bytes = bytes;
      }
    }
    scratch = static_cast<char*>(g->sf_scratch);
    chunk = std::min(nQ, g->sf_scratch_bytes / g->vecs_row_bytes);
  }
  const size_t M = g->dd.vectors.branches;
  for (size_t lo = 0; lo < nQ; lo += chunk) {
    const size_t n = std::min(chunk, nQ - lo);
    // (the failure counters accumulate over the chunks of a call)
    rc = enqueue(g, g->ws, g->d_fail, dQ + 3 * lo, n, mode, true, 1, reinterpret_cast<double*>(static_cast<char*>(dvals) + lo * g->vals_row_bytes),
                 reinterpret_cast<double*>(scratch), stream, nQ, false, lo == 0);
    if (rc) return rc;
    if (g->timing) cudaEventRecord(g->ev[4], stream);
    CU(launch_structure_factor(g->sf, dQ + 3 * lo, reinterpret_cast<const double*>(scratch), n, (uint32_t)M, dsf + lo * M, g->sm_count, stream));
    g->launches += 1;
    if (g->timing) {
      cudaEventRecord(g->ev[5], stream);
      cudaEventSynchronize(g->ev[5]);
      note_time(g, "consumer", g->ev[4], g->ev[5]);
    }
  }
  if (n_failed) {
    unsigned long long c[3];
    CU(cudaMemcpyAsync(c, g->d_fail, sizeof(c), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    *n_failed = c[0] + c[1] + c[2];
    return status_error(c, nQ);
  }
  return B200_OK;
}

// ----------------------------------------------------------------------------------------------------
// device-resident consumer: powder average (consumer.cu)
// ----------------------------------------------------------------------------------------------------
static int powder_config(b200_grid* g, const b200_powder_config_t* c, PowderDev* d) {
  if (!c) return fail(B200_E_INVALID, "NULL powder configuration");
  if (!c->n_qbins || !c->n_wbins || !(c->q_hi > c->q_lo) || !(c->w_hi > c->w_lo) || !(c->q_lo >= 0.0))
    return fail(B200_E_INVALID, "powder average: need n_qbins, n_wbins > 0, 0 <= q_lo < q_hi and w_lo < w_hi");
  if (c->weight != 0 && c->weight != 1) return fail(B200_E_INVALID, "powder average: weight must be 0 (|F|^2) or 1 (|F|^2 / omega)");
  if (g->dd.values.is_complex || g->dd.values.span < 1)
    return fail(B200_E_UNSUPPORTED, "powder average: the eigenvalue data must be real with at least one element per mode (the energy)");
  d->n_qbins = c->n_qbins;
  d->n_wbins = c->n_wbins;
  d->q_lo = c->q_lo;
  d->dq = (c->q_hi - c->q_lo) / c->n_qbins;
  d->inv_dq = c->n_qbins / (c->q_hi - c->q_lo);
  d->w_lo = c->w_lo;
  d->inv_dw = c->n_wbins / (c->w_hi - c->w_lo);
  d->weight = c->weight;
  const double* B = g->h_bz.to_xyz;
  for (int i = 0; i < 9; ++i) d->B[i] = B[i];
  const double det = B[0] * (B[4] * B[8] - B[5] * B[7]) - B[1] * (B[3] * B[8] - B[5] * B[6]) + B[2] * (B[3] * B[7] - B[4] * B[6]);
  if (det == 0.0) return fail(B200_E_INVALID, "singular reciprocal basis");
  const double inv[9] = {B[4] * B[8] - B[5] * B[7], B[2] * B[7] - B[1] * B[8], B[1] * B[5] - B[2] * B[4],
                         B[5] * B[6] - B[3] * B[8], B[0] * B[8] - B[2] * B[6], B[2] * B[3] - B[0] * B[5],
                         B[3] * B[7] - B[4] * B[6], B[1] * B[6] - B[0] * B[7], B[0] * B[4] - B[1] * B[3]};
  for (int i = 0; i < 9; ++i) d->Binv[i] = inv[i] / det;
  return B200_OK;
}

// the points of a call in chunks: generated on the device (Q == NULL) or copied from the host; the path with the structure factor
// fused into the pipelined cell kernel when possible (else through an eigenvector scratch); k_powder_bin; the histogram back
static int powder_run(b200_grid* g, const double* Q, size_t n_total, uint32_t flags, const PowderDev& pc, uint64_t n_dir_local, uint64_t dir_lo,
                      uint64_t n_dir, uint64_t seed, double* hist_out, double* counts_out, double* Q_out) {
  DeviceGuard guard;
  CU(cudaSetDevice(g->device));
  if (g->timing) g->kernel_ms.clear();
  cudaStream_t stream = g->stage[0].stream;
  PowderWS& w = g->pw;
  const size_t M = g->dd.vectors.branches, hist_elems = (size_t)pc.n_qbins * pc.n_wbins + pc.n_qbins;
  const bool points_only = Q_out != nullptr;
  bool fuse = !points_only && g->sf_fused && cell_sf_fusable(g->dd, g->sf);
  const size_t want_cap = std::min<size_t>(n_total, (size_t)1 << 22);
  if (w.cap < want_cap) {
    w.release();
    CU(cudaMalloc(&w.dQ, want_cap * 3 * sizeof(double)));
    if (!points_only) {
      CU(cudaMalloc(&w.dvals, std::max<size_t>(want_cap * g->vals_row_bytes, 8)));
      CU(cudaMalloc(&w.dsf, std::max<size_t>(want_cap * M * sizeof(double), 8)));
    }
    w.cap = want_cap;
  }
  if (!points_only && !w.dvals) {
    CU(cudaMalloc(&w.dvals, std::max<size_t>(w.cap * g->vals_row_bytes, 8)));
    CU(cudaMalloc(&w.dsf, std::max<size_t>(w.cap * M * sizeof(double), 8)));
  }
  if (!points_only && w.hist_elems < hist_elems) {
    if (w.dhist) cudaFree(w.dhist);
    w.dhist = nullptr;
    w.hist_elems = 0;
    CU(cudaMalloc(&w.dhist, hist_elems * sizeof(double)));
    w.hist_elems = hist_elems;
  }
  if (!points_only) CU(cudaMemsetAsync(w.dhist, 0, hist_elems * sizeof(double), stream));
  const uint32_t mode = MODE_IR | ((flags & B200_FLAG_NO_MOVE) ? MODE_NO_MOVE : 0u);
  unsigned long long total[3] = {0, 0, 0};
  for (size_t lo = 0; lo < n_total; lo += w.cap) {
    const size_t n = std::min(w.cap, n_total - lo);
    if (Q) CU(cudaMemcpyAsync(w.dQ, Q + 3 * lo, n * 3 * sizeof(double), cudaMemcpyHostToDevice, stream));
    else CU(launch_powder_q(w.dQ, lo, n, n_dir_local, dir_lo, n_dir, seed, pc, g->sm_count, stream));
    g->launches += Q ? 0 : 1;
    if (points_only) {
      CU(cudaMemcpyAsync(Q_out + 3 * lo, w.dQ, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, stream));
      CU(cudaStreamSynchronize(stream));
      continue;
    }
    bool done = false;
    if (fuse) {
      if (!w.dgs) {
        w.gcap = (uint32_t)std::max<size_t>(1024, w.cap / 16);
        CU(cudaMalloc(&w.dgs, (size_t)w.gcap * g->vecs_row_bytes));
      }
      const SfFuse fz{w.dsf, w.dgs, w.gcap};
      int rc = enqueue(g, g->ws, g->d_fail, w.dQ, n, mode, true, 1, w.dvals, nullptr, stream, n_total, false, true, &fz);
      if (rc == RC_NOT_FUSABLE) {
        fuse = false;
      } else {
        if (rc) return rc;
        unsigned long long c[N_FAIL];
        CU(cudaMemcpyAsync(c, g->d_fail, sizeof(c), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        if (c[3] == 0) {
          for (int k = 0; k < 3; ++k) total[k] += c[k];
          done = true;
        }  // else: more general-kernel points than compact rows in this chunk (a degenerate point set): once more, unfused
      }
    }
    if (!done) {
      if (!w.dvecs) CU(cudaMalloc(&w.dvecs, std::max<size_t>(w.cap * g->vecs_row_bytes, 8)));
      int rc = enqueue(g, g->ws, g->d_fail, w.dQ, n, mode, true, 1, w.dvals, w.dvecs, stream, n_total, false, true);
      if (rc) return rc;
      CU(launch_structure_factor(g->sf, w.dQ, w.dvecs, n, (uint32_t)M, w.dsf, g->sm_count, stream));
      g->launches += 1;
      unsigned long long c[3];
      CU(cudaMemcpyAsync(c, g->d_fail, sizeof(c), cudaMemcpyDeviceToHost, stream));
      CU(cudaStreamSynchronize(stream));
      for (int k = 0; k < 3; ++k) total[k] += c[k];
    }
    if (total[0] || total[1] || total[2]) return status_error(total, n_total);  // all-or-nothing, like ir_interpolate_at
    CU(launch_powder_bin(w.dQ, w.dvals, w.dsf, n, (uint32_t)M, g->dd.values.span, pc, w.dhist, w.dhist + (size_t)pc.n_qbins * pc.n_wbins, g->sm_count, stream));
    g->launches += 1;
  }
  if (!points_only) {
    std::vector<double> h(hist_elems);
    CU(cudaMemcpyAsync(h.data(), w.dhist, hist_elems * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    const size_t nh = (size_t)pc.n_qbins * pc.n_wbins;
    for (size_t i = 0; i < nh; ++i) hist_out[i] += h[i];
    for (size_t i = 0; i < pc.n_qbins; ++i) counts_out[i] += h[nh + i];
  }
  return B200_OK;
}

extern "C" int b200_ir_powder_bin(b200_grid_t* g, const double* Q, size_t nQ, uint32_t flags, const b200_powder_config_t* config, double* hist_out,
                                  double* counts_out) {
  int rc = check_sf_ready(g);
  if (rc) return rc;
  PowderDev pc;
  if ((rc = powder_config(g, config, &pc))) return rc;
  if (!hist_out || !counts_out || (nQ && !Q)) return fail(B200_E_INVALID, "NULL buffer");
  if (nQ == 0) return B200_OK;
  return powder_run(g, Q, nQ, flags, pc, 0, 0, 0, 0, hist_out, counts_out, nullptr);
}

static int sweep_args(const b200_powder_config_t* config, uint64_t n_dir, uint64_t dir_lo, uint64_t dir_hi, size_t* n_total) {
  if (!config) return fail(B200_E_INVALID, "NULL powder configuration");
  if (dir_lo > dir_hi || dir_hi > n_dir) return fail(B200_E_INVALID, "powder sweep: need dir_lo <= dir_hi <= n_dir");
  *n_total = (size_t)config->n_qbins * (size_t)(dir_hi - dir_lo);
  return B200_OK;
}
extern "C" int b200_ir_powder_sweep(b200_grid_t* g, const b200_powder_config_t* config, uint64_t n_dir, uint64_t seed, uint64_t dir_lo, uint64_t dir_hi,
                                    double* hist_out, double* counts_out) {
  int rc = check_sf_ready(g);
  if (rc) return rc;
  PowderDev pc;
  if ((rc = powder_config(g, config, &pc))) return rc;
  size_t n_total = 0;
  if ((rc = sweep_args(config, n_dir, dir_lo, dir_hi, &n_total))) return rc;
  if (!hist_out || !counts_out) return fail(B200_E_INVALID, "NULL buffer");
  if (n_total == 0) return B200_OK;
  return powder_run(g, nullptr, n_total, 0u, pc, dir_hi - dir_lo, dir_lo, n_dir, seed, hist_out, counts_out, nullptr);
}
extern "C" int b200_powder_points(b200_grid_t* g, const b200_powder_config_t* config, uint64_t n_dir, uint64_t seed, uint64_t dir_lo, uint64_t dir_hi,
                                  double* Q_out) {
  int rc = check_ready(g, true, 1);
  if (rc) return rc;
  PowderDev pc;
  if ((rc = powder_config(g, config, &pc))) return rc;
  size_t n_total = 0;
  if ((rc = sweep_args(config, n_dir, dir_lo, dir_hi, &n_total))) return rc;
  if (n_total && !Q_out) return fail(B200_E_INVALID, "NULL buffer");
  if (n_total == 0) return B200_OK;
  return powder_run(g, nullptr, n_total, 0u, pc, dir_hi - dir_lo, dir_lo, n_dir, seed, nullptr, nullptr, Q_out);
}

// ----------------------------------------------------------------------------------------------------
// introspection
// ----------------------------------------------------------------------------------------------------
extern "C" const char* b200_last_error(void) { return g_err.c_str(); }
extern "C" void* b200_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    g_err = "cudaHostAlloc failed";
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
extern "C" void b200_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}
extern "C" int b200_abi_version(void) { return B200_ABI_VERSION; }
extern "C" int b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
extern "C" uint64_t b200_grid_launch_count(const b200_grid_t* g) { return g ? g->launches : 0; }
extern "C" uint32_t b200_grid_last_path(const b200_grid_t* g) { return g ? g->last_path : 0u; }
extern "C" int b200_grid_sort_pairs(b200_grid_t* g, const uint32_t* pairs, size_t n_pairs, const b200_sort_config_t* cfg, int32_t* row_out,
                                    int32_t* col_out, double* cost_out) {
  if (!g || !cfg || (n_pairs && (!pairs || !row_out || !col_out))) return fail(B200_E_INVALID, "NULL argument");
  if (!g->has_data) return fail(B200_E_NODATA, "The interpolation data must be filled before sorting.");
  if (!g->dd.vectors.is_complex && g->dd.vectors.span)
    return fail(B200_E_UNSUPPORTED, "sort() of real-valued eigenvectors is not offloaded (the reference's anti-phase of real data is undefined)");
  for (size_t k = 0; k < n_pairs; ++k)
    if (pairs[2 * k] >= g->n_vertices || pairs[2 * k + 1] >= g->n_vertices) return fail(B200_E_INVALID, "vertex index out of range in pairs");
  CU(cudaSetDevice(g->device));
  CU(cudaDeviceSynchronize());
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  const size_t ws = std::min<size_t>((free_b + g->sort_ws.batch * 8 * g->sort_ws.branches * g->sort_ws.branches) / 4, (size_t)1 << 30);
  CU(run_sort_pairs(g->dd, cfg->values_costmult, cfg->values_vector_cost, cfg->vectors_costmult, cfg->vectors_vector_cost, pairs, n_pairs,
                    row_out, col_out, cost_out, g->sm_count, ws, g->sort_ws, &g->launches));
  return B200_OK;
}

extern "C" int b200_solve_assignments(const double* cost, size_t n, uint32_t modes, int32_t* row_out, int32_t* col_out, int device) {
  if (n && (!cost || !row_out || !col_out)) return fail(B200_E_INVALID, "NULL argument");
  if (modes == 0) return fail(B200_E_INVALID, "modes must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(B200_E_CUDA, "no CUDA device available; brille_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(B200_E_INVALID, "device index out of range");
  DeviceGuard guard;
  CU(cudaSetDevice(device));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  CU(run_match_only(cost, n, modes, row_out, col_out, sms));
  return B200_OK;
}

extern "C" int b200_grid_enable_timing(b200_grid_t* g, int on) {
  if (!g) return fail(B200_E_INVALID, "NULL grid");
  g->timing = on != 0;
  g->kernel_ms.clear();
  return B200_OK;
}
extern "C" double b200_grid_kernel_ms(const b200_grid_t* g, const char* name) {
  if (!g || !name) return -1.0;
  auto it = g->kernel_ms.find(name);
  if (it == g->kernel_ms.end() || it->second.second == 0) return -1.0;
  return it->second.first / it->second.second;
}
extern "C" int b200_grid_set_option(b200_grid_t* g, const char* name, double value) {
  if (!g || !name) return fail(B200_E_INVALID, "NULL argument");
  const std::string n(name);
  if (n == "interp_path") {
    if (value < 0 || value > 2) return fail(B200_E_INVALID, "interp_path must be 0 (auto), 1 (general) or 2 (cell-batched)");
    g->interp_path = (int)value;
  } else if (n == "split_locate") {
    g->split_locate = value != 0;
  } else if (n == "bin_points") {
    if (value < 1) return fail(B200_E_INVALID, "bin_points must be >= 1");
    g->bin_points = (size_t)value;
  } else if (n == "coop_locate") {
    g->coop_locate = value != 0;
  } else if (n == "tile") {
    if (value != 2 && value != 4) return fail(B200_E_INVALID, "tile must be 2 or 4");
    g->tile = (int)value;
  } else if (n == "cell_kernel") {
    if (value < 0 || value > 2) return fail(B200_E_INVALID, "cell_kernel must be 0 (auto), 1 (on-the-fly staging) or 2 (pipelined, cell table)");
    g->cell_kernel = (int)value;
  } else if (n == "chunk") {
    if (value < 32 || value > 256) return fail(B200_E_INVALID, "chunk must be in [32, 256]");
    g->chunk = ((uint32_t)value / 4u) * 4u;  // the weight tile is read with 16-byte loads
  } else if (n == "bounce") {
    g->bounce = value != 0;
  } else if (n == "replay_stores") {
    g->replay_stores = value != 0;
  } else if (n == "sf_fused") {
    g->sf_fused = value != 0;
  } else if (n == "host_chunk") {
    if (value < 0) return fail(B200_E_INVALID, "host_chunk must be >= 0");
    g->host_chunk = (size_t)value;
  } else {
    return fail(B200_E_INVALID, "unknown option " + n);
  }
  return B200_OK;
}
extern "C" int b200_grid_row_bytes(const b200_grid_t* g, size_t* v, size_t* w) {
  if (!g || !g->has_data) return fail(B200_E_NODATA, "The interpolation data must be filled before interpolating.");
  if (v) *v = g->vals_row_bytes;
  if (w) *w = g->vecs_row_bytes;
  return B200_OK;
}

// Triangle-mesh stage behind the extraction (include/d3h_mesh.h):
//
//   Mesh.get_edge   render/mesh.py:240-250   distinct undirected face edges in lexicographic order
//   auto_normals    render/mesh.py:418-446   area-weighted vertex normals, forward and adjoint
//
// Edges.  The reference sorts 3F (min,max) rows and runs torch.unique(dim=0) (a comparator sort of row indices plus a
// host sync for the size).  Here the rows are de-duplicated FIRST, in a hash set of packed 64-bit keys, which also
// counts the distinct edges per smaller endpoint; an exclusive scan of those counts gives every vertex its segment of the
// output, the set is poured into the segments and each segment (a vertex's larger neighbours, typically < 10) is sorted
// by one thread.  Lexicographic order = (segment order, order inside the segment), so no global sort is needed.  The
// edge count is published to mapped host memory by the scan kernel, before the rows are written.
//
//   insert (F threads)  ->  scan (V/4096 CTAs)  ->  fill (table slots)  ->  emit (V threads)  ->  big segments (CTAs)
//
// Normals.  One thread per face adds its normal to its three vertices with 16-byte vector atomics, one thread per
// vertex normalises.  The adjoint recomputes the per-vertex projection from the kept sums inside the per-face thread, so
// it is one kernel over the faces.  Work is O(faces): a few microseconds each, bounded by launch latency, not HBM.
#include <chrono>

#include "../../include/d3h_mesh.h"
#include "d3h_internal.cuh"

namespace d3h {

constexpr int kMeshThreads = 256;
constexpr int kMeshScanThreads = 1024;
constexpr int kMeshScanItems = 4;
constexpr int kMeshScanTile = kMeshScanThreads * kMeshScanItems;  // vertices per scan CTA
constexpr int kMeshTileShift = 12;                                // log2(kMeshScanTile)
constexpr int kSmallSegment = 32;                                 // neighbours one thread sorts in local memory
constexpr int kBigSegmentCtas = 148;
constexpr unsigned long long kEmptyKey = ~0ull;
static_assert((1 << kMeshTileShift) == kMeshScanTile, "tile shift");

enum MeshCtr { MC_NBIG = 0, MC_BAD = 1, MC_WORDS = 8 };

struct MeshWorkspace {
  unsigned long long* table;  // hash set of edge keys, cap_table slots (power of two, >= 2 * 3F)
  int64_t cap_table;
  int log2_cap;
  unsigned* ucount;    // V + 1: distinct edges whose smaller endpoint is v      } one zero-filled region
  unsigned* tile_tot;  // ntiles: the same, summed per scan tile                 }
  unsigned* ctr;       // MC_WORDS                                                }
  int64_t zero_bytes;
  unsigned* ustart;    // V + 1: exclusive scan of ucount = first output row of every vertex
  unsigned* his;       // 3F: larger endpoints, grouped by smaller endpoint, unsorted inside a group
  unsigned* big_list;  // vertices with more than kSmallSegment neighbours
  int64_t ntiles;
  int64_t total_bytes;
};

static MeshWorkspace carve_mesh_workspace(void* base, int64_t n_faces, int64_t n_verts) {
  MeshWorkspace w;
  const int64_t n_keys = 3 * n_faces;
  int lg = 10;
  while ((1ll << lg) < 2 * n_keys) ++lg;
  w.log2_cap = lg;
  w.cap_table = 1ll << lg;
  w.ntiles = (n_verts + kMeshScanTile - 1) / kMeshScanTile;
  char* p = static_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    char* r = p ? p + off : nullptr;
    off += (bytes + 255) & ~int64_t(255);
    return r;
  };
  w.table = reinterpret_cast<unsigned long long*>(take(w.cap_table * 8));
  const int64_t zero_words = (n_verts + 1) + w.ntiles + MC_WORDS;
  w.ucount = reinterpret_cast<unsigned*>(take(zero_words * 4));
  w.tile_tot = w.ucount ? w.ucount + (n_verts + 1) : nullptr;
  w.ctr = w.tile_tot ? w.tile_tot + w.ntiles : nullptr;
  w.zero_bytes = zero_words * 4;
  w.ustart = reinterpret_cast<unsigned*>(take((n_verts + 1) * 4));
  w.his = reinterpret_cast<unsigned*>(take((n_keys + 1) * 4));
  w.big_list = reinterpret_cast<unsigned*>(take((n_keys / (kSmallSegment + 1) + 1) * 4));
  w.total_bytes = off;
  return w;
}

// ---- edges -------------------------------------------------------------------------------------------------------
// One thread per face: the three (min,max) keys go into the hash set; the thread that claims an empty slot counts the
// edge for its smaller endpoint.
__global__ void __launch_bounds__(kMeshThreads)
mesh_edge_insert_kernel(const int64_t* __restrict__ faces, int64_t n_faces, int64_t n_verts,
                        unsigned long long* __restrict__ table, int log2_cap, unsigned* __restrict__ ucount,
                        unsigned* __restrict__ tile_tot, unsigned* __restrict__ ctr) {
  const int64_t f = (int64_t)blockIdx.x * kMeshThreads + threadIdx.x;
  if (f >= n_faces) return;
  long long v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = __ldg(faces + 3 * f + c);
  const unsigned long long mask = (1ull << log2_cap) - 1ull;
#pragma unroll
  for (int e = 0; e < 3; ++e) {  // (0,1) (1,2) (2,0): render/mesh.py:241-245
    const long long a = v[e], b = v[(e + 1) % 3];
    const long long lo = a < b ? a : b, hi = a < b ? b : a;
    if (lo < 0 || hi >= n_verts) {
      atomicOr(ctr + MC_BAD, 1u);
      continue;
    }
    const unsigned long long key = ((unsigned long long)lo << 32) | (unsigned long long)hi;
    unsigned long long h = (key * 0x9E3779B97F4A7C15ull) >> (64 - log2_cap);
    for (;;) {
      unsigned long long cur = __ldcg(table + h);
      if (cur == kEmptyKey) cur = atomicCAS(table + h, kEmptyKey, key);
      if (cur == kEmptyKey) {  // claimed: first sighting of this edge
        atomicAdd(ucount + lo, 1u);
        atomicAdd(tile_tot + (lo >> kMeshTileShift), 1u);
        break;
      }
      if (cur == key) break;
      h = (h + 1ull) & mask;
    }
  }
}

// One CTA per 4096 vertices: exclusive scan of the per-vertex edge counts.  The CTA's base is the sum of the tile
// totals before it (no chain between CTAs).  The last CTA knows E and publishes the sizes.
__global__ void __launch_bounds__(kMeshScanThreads)
mesh_edge_scan_kernel(const unsigned* __restrict__ ucount, const unsigned* __restrict__ tile_tot, int64_t n_verts,
                      unsigned* __restrict__ ustart, const unsigned* __restrict__ ctr, int64_t cap_edges,
                      d3h_mesh_counts* counts_dev, d3h_mesh_counts* counts_mapped, int64_t seq) {
  __shared__ unsigned s_part[32];
  __shared__ unsigned s_warp[32];
  const unsigned tile = blockIdx.x;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  unsigned part = 0;
  for (unsigned i = threadIdx.x; i < tile; i += kMeshScanThreads) part += __ldcg(tile_tot + i);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
  if (lane == 0) s_part[warp] = part;

  const int64_t v0 = (int64_t)tile * kMeshScanTile + (int64_t)threadIdx.x * kMeshScanItems;
  unsigned c[kMeshScanItems];
  unsigned mine = 0;
#pragma unroll
  for (int k = 0; k < kMeshScanItems; ++k) {
    c[k] = (v0 + k < n_verts) ? __ldcg(ucount + v0 + k) : 0u;
    mine += c[k];
  }
  unsigned incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
    if ((int)lane >= d) incl += up;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  unsigned base = 0, before = 0, total = 0;
  for (int w = 0; w < kMeshScanThreads / 32; ++w) {
    base += s_part[w];
    const unsigned t = s_warp[w];
    if (w < (int)warp) before += t;
    total += t;
  }
  unsigned run = base + before + incl - mine;
#pragma unroll
  for (int k = 0; k < kMeshScanItems; ++k) {
    if (v0 + k < n_verts) ustart[v0 + k] = run;
    run += c[k];
  }
  if (tile == gridDim.x - 1 && threadIdx.x == 0) {
    const long long n_edges = (long long)base + (long long)total;
    ustart[n_verts] = (unsigned)n_edges;
    d3h_mesh_counts r;
    r.n_edges = n_edges;
    r.bad_index = __ldcg(ctr + MC_BAD) ? 1 : 0;
    r.overflow = n_edges > cap_edges ? 1 : 0;
    r.seq = seq;
    *counts_dev = r;
    if (counts_mapped != nullptr) {  // `seq` last, after a system-scope fence (as d3h_counts in d3h_surface.cu)
      volatile int64_t* dst = reinterpret_cast<volatile int64_t*>(counts_mapped);
      dst[0] = r.n_edges;
      dst[1] = r.bad_index;
      dst[2] = r.overflow;
      __threadfence_system();
      dst[3] = seq;
    }
  }
}

// One thread per slot of the hash set: every distinct edge takes a place in its smaller endpoint's segment.  The counts
// are handed back one by one (they are zero again afterwards).
__global__ void __launch_bounds__(kMeshThreads)
mesh_edge_fill_kernel(const unsigned long long* __restrict__ table, int64_t cap_table,
                      const unsigned* __restrict__ ustart, unsigned* __restrict__ ucount, unsigned* __restrict__ his) {
  const int64_t i = (int64_t)blockIdx.x * kMeshThreads + threadIdx.x;
  if (i >= cap_table) return;
  const unsigned long long key = __ldcs(table + i);
  if (key == kEmptyKey) return;
  const unsigned lo = (unsigned)(key >> 32), hi = (unsigned)key;
  const unsigned slot = ustart[lo] + atomicSub(ucount + lo, 1u) - 1u;
  his[slot] = hi;
}

__device__ __forceinline__ void store_edge(int64_t* edges, int64_t cap_edges, int64_t row, unsigned lo, unsigned hi) {
  if (row >= cap_edges) return;
  longlong2 r;
  r.x = (long long)lo;
  r.y = (long long)hi;
  reinterpret_cast<longlong2*>(edges)[row] = r;
}

// One thread per vertex: sort the (few) larger neighbours and write the int64 rows.  Long segments are queued.
__global__ void __launch_bounds__(kMeshThreads)
mesh_edge_emit_kernel(const unsigned* __restrict__ ustart, const unsigned* __restrict__ his, int64_t n_verts,
                      int64_t* __restrict__ edges, int64_t cap_edges, unsigned* __restrict__ big_list,
                      unsigned* __restrict__ ctr) {
  const int64_t v = (int64_t)blockIdx.x * kMeshThreads + threadIdx.x;
  if (v >= n_verts) return;
  const unsigned s = ustart[v], d = ustart[v + 1] - s;
  if (d == 0u) return;
  if (d > (unsigned)kSmallSegment) {
    big_list[atomicAdd(ctr + MC_NBIG, 1u)] = (unsigned)v;
    return;
  }
  unsigned h[kSmallSegment];
  for (unsigned k = 0; k < d; ++k) {  // insertion sort
    const unsigned x = his[s + k];
    unsigned j = k;
    while (j > 0u && h[j - 1u] > x) {
      h[j] = h[j - 1u];
      --j;
    }
    h[j] = x;
  }
  for (unsigned k = 0; k < d; ++k) store_edge(edges, cap_edges, (int64_t)s + k, (unsigned)v, h[k]);
}

// Long segments (a hub vertex): one CTA per segment, every element finds its rank by counting the smaller ones (all
// elements of a segment are distinct).  Quadratic, but such vertices do not occur in extracted surfaces.
__global__ void __launch_bounds__(kMeshThreads)
mesh_edge_big_kernel(const unsigned* __restrict__ ustart, const unsigned* __restrict__ his,
                     const unsigned* __restrict__ big_list, const unsigned* __restrict__ ctr,
                     int64_t* __restrict__ edges, int64_t cap_edges) {
  const unsigned n_big = __ldcg(ctr + MC_NBIG);
  for (unsigned i = blockIdx.x; i < n_big; i += gridDim.x) {
    const unsigned v = big_list[i];
    const unsigned s = ustart[v], d = ustart[v + 1] - s;
    for (unsigned k = threadIdx.x; k < d; k += kMeshThreads) {
      const unsigned x = his[s + k];
      unsigned rank = 0;
      for (unsigned j = 0; j < d; ++j) rank += (__ldg(his + s + j) < x) ? 1u : 0u;
      store_edge(edges, cap_edges, (int64_t)s + rank, v, x);
    }
  }
}

// ---- normals -----------------------------------------------------------------------------------------------------
// torch.cross on CPU: component = fma(a_i, b_j, -fl(a_j * b_i)) (same restatement as d3h_surface.cu)
__device__ __forceinline__ float mesh_cross_comp(float ai, float bj, float aj, float bi) {
  return __fmaf_rn(ai, bj, -__fmul_rn(aj, bi));
}

__device__ __forceinline__ bool load_face(const int64_t* faces, int64_t f, int64_t n_verts, long long i[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) i[c] = __ldg(faces + 3 * f + c);
  return i[0] >= 0 && i[1] >= 0 && i[2] >= 0 && i[0] < n_verts && i[1] < n_verts && i[2] < n_verts;
}

__device__ __forceinline__ void face_sides(const float* pos, const long long i[3], float a[3], float b[3]) {
  float p[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int k = 0; k < 3; ++k) p[c][k] = __ldg(pos + 3 * i[c] + k);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    a[k] = __fsub_rn(p[1][k], p[0][k]);  // v1 - v0
    b[k] = __fsub_rn(p[2][k], p[0][k]);  // v2 - v0
  }
}

// render/mesh.py:431-437: face normal added to the three vertices of the face
__global__ void __launch_bounds__(kMeshThreads)
mesh_normal_splat_kernel(const float* __restrict__ pos, const int64_t* __restrict__ faces, int64_t n_verts,
                         int64_t n_faces, float* __restrict__ acc, int32_t* __restrict__ bad) {
  const int64_t f = (int64_t)blockIdx.x * kMeshThreads + threadIdx.x;
  if (f >= n_faces) return;
  long long i[3];
  if (!load_face(faces, f, n_verts, i)) {
    if (bad) *bad = 1;
    return;
  }
  float a[3], b[3];
  face_sides(pos, i, a, b);
  const float4 n = make_float4(mesh_cross_comp(a[1], b[2], a[2], b[1]), mesh_cross_comp(a[2], b[0], a[0], b[2]),
                               mesh_cross_comp(a[0], b[1], a[1], b[0]), 0.f);
#pragma unroll
  for (int c = 0; c < 3; ++c) atomicAdd(reinterpret_cast<float4*>(acc) + i[c], n);
}

// Exactly three faces: torch.cross without `dim` takes the FIRST axis of size 3, i.e. the face axis of the (3,3)
// operands (render/mesh.py:431).  Column c of the result is (a[:,c]) x (b[:,c]).  One thread.
__global__ void mesh_normal_splat3_kernel(const float* __restrict__ pos, const int64_t* __restrict__ faces,
                                          int64_t n_verts, float* __restrict__ acc, int32_t* __restrict__ bad) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long i[3][3];
  float a[3][3], b[3][3];
  for (int r = 0; r < 3; ++r) {
    if (!load_face(faces, r, n_verts, i[r])) {
      if (bad) *bad = 1;
      return;
    }
    face_sides(pos, i[r], a[r], b[r]);
  }
  for (int c = 0; c < 3; ++c) {
    float fn[3];
    fn[0] = mesh_cross_comp(a[1][c], b[2][c], a[2][c], b[1][c]);
    fn[1] = mesh_cross_comp(a[2][c], b[0][c], a[0][c], b[2][c]);
    fn[2] = mesh_cross_comp(a[0][c], b[1][c], a[1][c], b[0][c]);
    for (int r = 0; r < 3; ++r)
      for (int k = 0; k < 3; ++k) atomicAdd(acc + 4 * i[r][k] + c, fn[r]);
  }
}

// render/mesh.py:439-441 with render/util.py:19-29: degenerate sums become (0,0,1), then x / sqrt(max(x.x, 1e-20))
__global__ void __launch_bounds__(kMeshThreads)
mesh_normal_finish_kernel(const float* __restrict__ acc, int64_t n_verts, float* __restrict__ v_nrm) {
  const int64_t v = (int64_t)blockIdx.x * kMeshThreads + threadIdx.x;
  if (v >= n_verts) return;
  const float4 s = __ldcg(reinterpret_cast<const float4*>(acc) + v);
  float x = s.x, y = s.y, z = s.z;
  const float d = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  if (!(d > 1e-20f)) {
    x = 0.f;
    y = 0.f;
    z = 1.f;
  }
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  const float len = __fsqrt_rn(fmaxf(d2, 1e-20f));
  v_nrm[3 * v + 0] = __fdiv_rn(x, len);
  v_nrm[3 * v + 1] = __fdiv_rn(y, len);
  v_nrm[3 * v + 2] = __fdiv_rn(z, len);
}

// d loss / d (un-normalised sum) of one vertex: (g - n (n.g)) / |s| where the sum was kept, 0 where it was replaced
__device__ __forceinline__ void normal_sum_adjoint(const float* acc, const float* g_nrm, long long v, float out[3]) {
  const float4 s = __ldg(reinterpret_cast<const float4*>(acc) + v);
  const float d = s.x * s.x + s.y * s.y + s.z * s.z;
  if (!(d > 1e-20f)) {
    out[0] = out[1] = out[2] = 0.f;
    return;
  }
  const float inv = 1.f / sqrtf(d);
  const float nx = s.x * inv, ny = s.y * inv, nz = s.z * inv;
  const float gx = __ldg(g_nrm + 3 * v), gy = __ldg(g_nrm + 3 * v + 1), gz = __ldg(g_nrm + 3 * v + 2);
  const float dp = nx * gx + ny * gy + nz * gz;
  out[0] = (gx - nx * dp) * inv;
  out[1] = (gy - ny * dp) * inv;
  out[2] = (gz - nz * dp) * inv;
}

// n = a x b:  d/da = b x G,  d/db = G x a;  a = v1 - v0, b = v2 - v0
__device__ __forceinline__ void cross_adjoint(const float a[3], const float b[3], const float G[3], float ga[3],
                                              float gb[3]) {
  ga[0] = b[1] * G[2] - b[2] * G[1];
  ga[1] = b[2] * G[0] - b[0] * G[2];
  ga[2] = b[0] * G[1] - b[1] * G[0];
  gb[0] = G[1] * a[2] - G[2] * a[1];
  gb[1] = G[2] * a[0] - G[0] * a[2];
  gb[2] = G[0] * a[1] - G[1] * a[0];
}

__global__ void __launch_bounds__(kMeshThreads)
mesh_normal_adjoint_kernel(const float* __restrict__ pos, const int64_t* __restrict__ faces, int64_t n_verts,
                           int64_t n_faces, const float* __restrict__ acc, const float* __restrict__ g_nrm,
                           float* __restrict__ g_pos) {
  const int64_t f = (int64_t)blockIdx.x * kMeshThreads + threadIdx.x;
  if (f >= n_faces) return;
  long long i[3];
  if (!load_face(faces, f, n_verts, i)) return;
  float a[3], b[3], G[3] = {0.f, 0.f, 0.f};
  face_sides(pos, i, a, b);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float g[3];
    normal_sum_adjoint(acc, g_nrm, i[c], g);
    G[0] += g[0];
    G[1] += g[1];
    G[2] += g[2];
  }
  float ga[3], gb[3];
  cross_adjoint(a, b, G, ga, gb);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    atomicAdd(g_pos + 3 * i[1] + k, ga[k]);
    atomicAdd(g_pos + 3 * i[2] + k, gb[k]);
    atomicAdd(g_pos + 3 * i[0] + k, -(ga[k] + gb[k]));
  }
}

// adjoint of the three-face case: column c of the face-normal matrix is a[:,c] x b[:,c]
__global__ void mesh_normal_adjoint3_kernel(const float* __restrict__ pos, const int64_t* __restrict__ faces,
                                            int64_t n_verts, const float* __restrict__ acc,
                                            const float* __restrict__ g_nrm, float* __restrict__ g_pos) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long i[3][3];
  float a[3][3], b[3][3], G[3][3];  // G[r][c]: d loss / d face-normal component c of face r
  for (int r = 0; r < 3; ++r) {
    if (!load_face(faces, r, n_verts, i[r])) return;
    face_sides(pos, i[r], a[r], b[r]);
    G[r][0] = G[r][1] = G[r][2] = 0.f;
    for (int k = 0; k < 3; ++k) {
      float g[3];
      normal_sum_adjoint(acc, g_nrm, i[r][k], g);
      G[r][0] += g[0];
      G[r][1] += g[1];
      G[r][2] += g[2];
    }
  }
  for (int c = 0; c < 3; ++c) {
    const float ac[3] = {a[0][c], a[1][c], a[2][c]}, bc[3] = {b[0][c], b[1][c], b[2][c]};
    const float Gc[3] = {G[0][c], G[1][c], G[2][c]};
    float ga[3], gb[3];
    cross_adjoint(ac, bc, Gc, ga, gb);
    for (int r = 0; r < 3; ++r) {
      atomicAdd(g_pos + 3 * i[r][1] + c, ga[r]);
      atomicAdd(g_pos + 3 * i[r][2] + c, gb[r]);
      atomicAdd(g_pos + 3 * i[r][0] + c, -(ga[r] + gb[r]));
    }
  }
}

static inline unsigned ctas_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

static int finish_call(const char* who) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", who, cudaGetErrorString(e));
    return D3H_E_CUDA;
  }
  return D3H_OK;
}

}  // namespace d3h

using namespace d3h;

extern "C" int64_t d3h_mesh_edges_workspace_bytes(int64_t n_faces, int64_t n_verts) {
  if (n_faces < 0 || n_verts < 0 || 3 * n_faces >= (1ll << 31) || n_verts >= (1ll << 31)) return D3H_E_BADARG;
  return carve_mesh_workspace(nullptr, n_faces, n_verts).total_bytes;
}

extern "C" int d3h_mesh_edges(const int64_t* faces, int64_t n_faces, int64_t n_verts, int64_t* edges, int64_t cap_edges,
                              void* workspace, int64_t workspace_bytes, d3h_mesh_counts* counts_dev,
                              d3h_mesh_counts* counts_host, int64_t seq, d3h_stream_t s) {
  const char* who = "d3h_mesh_edges";
  if (n_faces < 0 || n_verts < 0 || cap_edges < 0 || 3 * n_faces >= (1ll << 31) || n_verts >= (1ll << 31)) {
    set_error("%s: n_faces=%lld / n_verts=%lld / cap_edges=%lld out of range", who, (long long)n_faces,
              (long long)n_verts, (long long)cap_edges);
    return D3H_E_BADARG;
  }
  if (!counts_dev || !workspace || (n_faces > 0 && !faces) || (cap_edges > 0 && !edges)) {
    set_error("%s: null pointer", who);
    return D3H_E_BADARG;
  }
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(edges) & 15) ||
      (reinterpret_cast<uintptr_t>(faces) & 7)) {
    set_error("%s: workspace must be 256-byte, edges 16-byte, faces 8-byte aligned", who);
    return D3H_E_BADARG;
  }
  const MeshWorkspace w = carve_mesh_workspace(workspace, n_faces, n_verts);
  if (workspace_bytes < w.total_bytes) {
    set_error("%s: workspace has %lld bytes, %lld needed", who, (long long)workspace_bytes, (long long)w.total_bytes);
    return D3H_E_SMALLWS;
  }
  cudaStream_t stream = (cudaStream_t)s;
  d3h_mesh_counts* mapped =
      reinterpret_cast<d3h_mesh_counts*>(mapped_counts_pointer(reinterpret_cast<d3h_counts*>(counts_host)));
  ProfScope ps(K_MESH_EDGES, stream);  // d3h_profile_*: the whole edge pipeline as one entry
  cudaMemsetAsync(w.table, 0xff, (size_t)w.cap_table * 8, stream);
  cudaMemsetAsync(w.ucount, 0, (size_t)w.zero_bytes, stream);
  if (n_faces > 0)
    launch_k(mesh_edge_insert_kernel, ctas_for(n_faces, kMeshThreads), kMeshThreads, stream, kLaunchLatency, faces,
             n_faces, n_verts, w.table, w.log2_cap, w.ucount, w.tile_tot, w.ctr);
  // a mesh without vertices still publishes its (zero) sizes: one scan CTA
  launch_k(mesh_edge_scan_kernel, (unsigned)(w.ntiles > 0 ? w.ntiles : 1), kMeshScanThreads, stream, kLaunchLatency,
           (const unsigned*)w.ucount, (const unsigned*)w.tile_tot, n_verts, w.ustart, (const unsigned*)w.ctr, cap_edges,
           counts_dev, mapped, seq);
  if (n_faces > 0 && n_verts > 0) {
    launch_k(mesh_edge_fill_kernel, ctas_for(w.cap_table, kMeshThreads), kMeshThreads, stream, kLaunchLatency,
             (const unsigned long long*)w.table, w.cap_table, (const unsigned*)w.ustart, w.ucount, w.his);
    launch_k(mesh_edge_emit_kernel, ctas_for(n_verts, kMeshThreads), kMeshThreads, stream, kLaunchLatency,
             (const unsigned*)w.ustart, (const unsigned*)w.his, n_verts, edges, cap_edges, w.big_list, w.ctr);
    launch_k(mesh_edge_big_kernel, (unsigned)kBigSegmentCtas, kMeshThreads, stream, kLaunchLatency,
             (const unsigned*)w.ustart, (const unsigned*)w.his, (const unsigned*)w.big_list, (const unsigned*)w.ctr,
             edges, cap_edges);
  }
  if (counts_host && mapped == nullptr)  // not device-mapped: copy at the end, the caller synchronises
    cudaMemcpyAsync(counts_host, counts_dev, sizeof(d3h_mesh_counts), cudaMemcpyDeviceToHost, stream);
  return finish_call(who);
}

extern "C" int d3h_mesh_wait_counts(const d3h_mesh_counts* counts_host, int64_t seq, int64_t timeout_us) {
  if (!counts_host) {
    set_error("d3h_mesh_wait_counts: null pointer");
    return D3H_E_BADARG;
  }
  const volatile int64_t* flag = &counts_host->seq;
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned spins = 0;; ++spins) {
    if (*flag == seq) {
      __atomic_thread_fence(__ATOMIC_ACQUIRE);
      return D3H_OK;
    }
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
    if (timeout_us > 0 && (spins & 1023u) == 1023u) {
      const auto dt =
          std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
      if (dt > timeout_us) {
        set_error("d3h_mesh_wait_counts: seq %lld not published after %lld us", (long long)seq, (long long)dt);
        return D3H_E_TIMEOUT;
      }
    }
  }
}

static int check_normal_args(const char* who, const float* pos, const int64_t* faces, int64_t n_verts, int64_t n_faces,
                             const void* out, const float* acc) {
  if (n_faces < 0 || n_verts < 0 || n_faces >= (1ll << 31) || n_verts >= (1ll << 31)) {
    set_error("%s: n_faces=%lld / n_verts=%lld out of range", who, (long long)n_faces, (long long)n_verts);
    return D3H_E_BADARG;
  }
  if (n_verts > 0 && (!pos || !out || !acc)) {
    set_error("%s: null pointer", who);
    return D3H_E_BADARG;
  }
  if (n_faces > 0 && !faces) {
    set_error("%s: null faces", who);
    return D3H_E_BADARG;
  }
  if ((reinterpret_cast<uintptr_t>(acc) & 15) || (reinterpret_cast<uintptr_t>(faces) & 7) ||
      (reinterpret_cast<uintptr_t>(pos) & 3) || (reinterpret_cast<uintptr_t>(out) & 3)) {
    set_error("%s: acc must be 16-byte, faces 8-byte, pos / outputs 4-byte aligned", who);
    return D3H_E_BADARG;
  }
  return D3H_OK;
}

extern "C" int d3h_mesh_normals_forward(const float* pos, const int64_t* faces, int64_t n_verts, int64_t n_faces,
                                        float* v_nrm, float* acc, int32_t* bad, d3h_stream_t s) {
  const char* who = "d3h_mesh_normals_forward";
  const int rc = check_normal_args(who, pos, faces, n_verts, n_faces, v_nrm, acc);
  if (rc) return rc;
  if (n_verts == 0) return D3H_OK;
  cudaStream_t stream = (cudaStream_t)s;
  ProfScope ps(K_MESH_NORMALS, stream);
  cudaMemsetAsync(acc, 0, (size_t)n_verts * 16, stream);
  if (n_faces == 3)
    launch_k(mesh_normal_splat3_kernel, 1u, 32u, stream, kLaunchLatency, pos, faces, n_verts, acc, bad);
  else if (n_faces > 0)
    launch_k(mesh_normal_splat_kernel, ctas_for(n_faces, kMeshThreads), kMeshThreads, stream, kLaunchLatency, pos, faces,
             n_verts, n_faces, acc, bad);
  launch_k(mesh_normal_finish_kernel, ctas_for(n_verts, kMeshThreads), kMeshThreads, stream, kLaunchLatency,
           (const float*)acc, n_verts, v_nrm);
  return finish_call(who);
}

extern "C" int d3h_mesh_normals_backward(const float* pos, const int64_t* faces, int64_t n_verts, int64_t n_faces,
                                         const float* acc, const float* g_nrm, float* g_pos, d3h_stream_t s) {
  const char* who = "d3h_mesh_normals_backward";
  const int rc = check_normal_args(who, pos, faces, n_verts, n_faces, g_pos, acc);
  if (rc) return rc;
  if (n_verts == 0) return D3H_OK;
  if (!g_nrm) {
    set_error("%s: null g_nrm", who);
    return D3H_E_BADARG;
  }
  cudaStream_t stream = (cudaStream_t)s;
  ProfScope ps(K_MESH_ADJOINT, stream);
  cudaMemsetAsync(g_pos, 0, (size_t)n_verts * 12, stream);
  if (n_faces == 3)
    launch_k(mesh_normal_adjoint3_kernel, 1u, 32u, stream, kLaunchLatency, pos, faces, n_verts, acc, g_nrm, g_pos);
  else if (n_faces > 0)
    launch_k(mesh_normal_adjoint_kernel, ctas_for(n_faces, kMeshThreads), kMeshThreads, stream, kLaunchLatency, pos,
             faces, n_verts, n_faces, acc, g_nrm, g_pos);
  return finish_call(who);
}

// Shared device helpers, case tables and the workspace layout of the sm_100a extraction kernels.
//
// Case tables are the reference's data (geometry/gshell_tets.py:91-190).  They are stored here in the
// forms the kernels index directly; tests/test_cabi.py checks them element-for-element against the
// oracle's copies through d3h_debug_table().
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/d3h_tets.h"

namespace d3h {

constexpr int kWarp = 32;
constexpr float kEps12 = 1e-12f;

// ------------------------------------------------------------------------------------------------
// tables
// ------------------------------------------------------------------------------------------------
// Each table is written once as an initializer macro so that the __constant__ copy the kernels index and the host
// copy returned by d3h_debug_table() (compared with the oracle in tests/test_cabi.py, no GPU needed) cannot diverge.
// number of watertight triangles per occupancy code (gshell_tets.py:186)
#define D3H_T_NUM_TRI {0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0}
// tet edge id -> local endpoints (gshell_tets.py:187)
#define D3H_T_EDGE_P {0, 0, 0, 1, 1, 2}
#define D3H_T_EDGE_Q {1, 2, 3, 2, 3, 3}
// polygon loop per code: tet edge ids of the loop corners, first 3 (tri) or 4 (quad) entries of
// mesh_edge_table (gshell_tets.py:110-127)
#define D3H_T_LOOP_EDGE                                                                              \
  {{-1, -1, -1, -1}, {1, 0, 2, -1}, {4, 0, 3, -1}, {1, 3, 4, 2}, {3, 1, 5, -1}, {2, 5, 3, 0},       \
   {1, 5, 4, 0},     {4, 2, 5, -1}, {4, 5, 2, -1}, {4, 5, 1, 0}, {3, 5, 2, 0},  {1, 3, 5, -1},      \
   {4, 3, 1, 2},     {3, 0, 4, -1}, {2, 0, 1, -1}, {-1, -1, -1, -1}}
// watertight triangles per code as tet edge ids (gshell_tets.py:91-108)
#define D3H_T_TRI_EDGE                                                                               \
  {{-1, -1, -1, -1, -1, -1}, {1, 0, 2, -1, -1, -1}, {4, 0, 3, -1, -1, -1}, {1, 4, 2, 1, 3, 4},      \
   {3, 1, 5, -1, -1, -1},    {2, 3, 0, 2, 5, 3},    {1, 4, 0, 1, 5, 4},    {4, 2, 5, -1, -1, -1},   \
   {4, 5, 2, -1, -1, -1},    {4, 1, 0, 4, 5, 1},    {3, 2, 0, 3, 5, 2},    {1, 3, 5, -1, -1, -1},   \
   {4, 1, 2, 4, 3, 1},       {3, 0, 4, -1, -1, -1}, {2, 0, 1, -1, -1, -1}, {-1, -1, -1, -1, -1, -1}}
// mSDF cut of a triangle polygon: locals 0-2 = corners, 3-5 = boundary vertices (gshell_tets.py:130-147)
#define D3H_T_CUT_TRI                                                                                \
  {{-1, -1, -1, -1, -1, -1}, {4, 2, 5, -1, -1, -1}, {3, 1, 4, -1, -1, -1}, {3, 1, 2, 3, 2, 5},      \
   {0, 3, 5, -1, -1, -1},    {0, 3, 4, 0, 4, 2},    {0, 1, 4, 0, 4, 5},    {0, 1, 2, -1, -1, -1}}
#define D3H_T_NUM_CUT_TRI {0, 1, 1, 2, 1, 2, 2, 1}  // gshell_tets.py:189
// mSDF cut of a quad polygon: locals 0-3 = corners, 4-7 = boundary vertices (gshell_tets.py:149-184)
#define D3H_T_CUT_QUAD                                                                                     \
  {{-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1}, {6, 3, 7, -1, -1, -1, -1, -1, -1, -1, -1, -1},       \
   {5, 2, 6, -1, -1, -1, -1, -1, -1, -1, -1, -1},    {5, 2, 7, 3, 7, 2, -1, -1, -1, -1, -1, -1},          \
   {4, 1, 5, -1, -1, -1, -1, -1, -1, -1, -1, -1},    {4, 1, 5, 4, 5, 7, 5, 6, 7, 7, 6, 3},                \
   {4, 1, 2, 6, 4, 2, -1, -1, -1, -1, -1, -1},       {4, 1, 2, 7, 4, 2, 7, 2, 3, -1, -1, -1},             \
   {0, 4, 7, -1, -1, -1, -1, -1, -1, -1, -1, -1},    {0, 4, 6, 3, 0, 6, -1, -1, -1, -1, -1, -1},          \
   {0, 4, 5, 0, 5, 2, 0, 2, 6, 0, 6, 7},             {0, 4, 5, 0, 5, 2, 0, 2, 3, -1, -1, -1},             \
   {0, 1, 5, 7, 0, 5, -1, -1, -1, -1, -1, -1},       {0, 1, 5, 0, 5, 6, 0, 6, 3, -1, -1, -1},             \
   {0, 1, 2, 0, 2, 6, 0, 6, 7, -1, -1, -1},          {0, 1, 2, 0, 2, 3, -1, -1, -1, -1, -1, -1}}
#define D3H_T_NUM_CUT_QUAD {0, 1, 1, 2, 1, 4, 2, 3, 1, 2, 4, 3, 2, 3, 3, 2}  // gshell_tets.py:190

static __device__ __constant__ int8_t c_num_tri[16] = D3H_T_NUM_TRI;
static __device__ __constant__ int8_t c_edge_p[6] = D3H_T_EDGE_P;
static __device__ __constant__ int8_t c_edge_q[6] = D3H_T_EDGE_Q;
static __device__ __constant__ int8_t c_loop_edge[16][4] = D3H_T_LOOP_EDGE;
static __device__ __constant__ int8_t c_tri_edge[16][6] = D3H_T_TRI_EDGE;
static __device__ __constant__ int8_t c_cut_tri[8][6] = D3H_T_CUT_TRI;
static __device__ __constant__ int8_t c_num_cut_tri[8] = D3H_T_NUM_CUT_TRI;
static __device__ __constant__ int8_t c_cut_quad[16][12] = D3H_T_CUT_QUAD;
static __device__ __constant__ int8_t c_num_cut_quad[16] = D3H_T_NUM_CUT_QUAD;

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
#ifndef D3H_CPU_EMU
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
// status words of the decoupled look-back scans are self-contained (flag + value in one word), so relaxed
// gpu-scope accesses are enough -- no fence, hence no L1 invalidation in the streaming kernels.
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// streaming 16-byte load that does not allocate in L1 (the tet index stream is read exactly once)
__device__ __forceinline__ int4 ld_stream_int4(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// the same for one 4-byte word (coalesced rows of the transposed edge list: read once, must not evict the sign bitmap)
__device__ __forceinline__ int ld_stream_s32(const int32_t* p) {
  int r;
  asm("ld.global.nc.L1::no_allocate.L2::256B.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#else
// tests/emu (functional CPU emulation of the kernels, test infrastructure only): no PTX
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << (threadIdx.x & 31u)) - 1u; }
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) { return emu_load_relaxed(p); }
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) { emu_store_relaxed(p, v); }
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) { return emu_load_relaxed(p); }
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) { emu_store_relaxed(p, v); }
__device__ __forceinline__ int4 ld_stream_int4(const int4* p) { return *p; }
__device__ __forceinline__ int ld_stream_s32(const int32_t* p) { return *p; }
__device__ __forceinline__ unsigned long long global_timer_ns() { return emu::now_ns(); }
#endif
// Several threads may store the SAME value to one address (every corner of a vertex writes the vertex's tangent rows on
// the static edge path).  A plain store; the race detector of the CPU emulation is told that it is deliberate.
__device__ __forceinline__ void store_same_value(float* p, float v) {
#ifdef D3H_CPU_EMU
  emu_store_relaxed(p, v);
#else
  *p = v;
#endif
}
// Diagnostics (d3h_trace_enable): per call and kernel kind, [0] = time block 0 started, [1] = latest block exit.
constexpr int kTraceFrames = 64;
constexpr int kTraceKinds = 24;
__device__ __forceinline__ unsigned long long* trace_begin(unsigned long long* table, unsigned frame, int kind) {
  if (table == nullptr) return nullptr;
  unsigned long long* slot = table + ((frame % kTraceFrames) * kTraceKinds + kind) * 2;
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) slot[0] = global_timer_ns();
  return slot;
}
__device__ __forceinline__ void trace_end(unsigned long long* slot) {
  if (slot != nullptr && threadIdx.x == 0) atomicMax(slot + 1, global_timer_ns());
}
// Programmatic dependent launch (kernels launched with launch_k_dep): the kernel may be scheduled while its predecessor
// in the stream is still running -- its blocks arrive, load their parameters and stop HERE until the predecessor has
// completed and flushed its writes; it lets its own successor do the same.  Must be the first statement of the kernel
// (nothing before it may touch global memory).  A no-op for a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_enter() {
#ifndef D3H_CPU_EMU
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
// bit v of a vertex bitmap (sdf > 0, or msdf > 0)
__device__ __forceinline__ unsigned occ_of(const unsigned* __restrict__ bits, int v) {
  return (__ldg(bits + (v >> 5)) >> (v & 31)) & 1u;
}
__device__ __forceinline__ float fsign(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

// fl(xa*wa) + fl(xb*wb): the reference multiplies and adds in separate kernels (no FMA), SURVEY A.4
__device__ __forceinline__ float lerp2(float xa, float wa, float xb, float wb) {
  return __fadd_rn(__fmul_rn(xa, wa), __fmul_rn(xb, wb));
}

// Zero-crossing weights of edge (a,b): gshell_tets.py:292-299.  w0 multiplies the a end, w1 the b end.
__device__ __forceinline__ void crossing_weights(float sa, float sb, float& w0, float& w1, float& dd) {
  float e0 = sa, e1 = -sb;
  float d = __fadd_rn(e0, e1);
  dd = __fmul_rn(fsign(d), __fadd_rn(fabsf(d), kEps12));
  if (dd == 0.f) dd = kEps12;
  w0 = __fdiv_rn(e1, dd);
  w1 = __fdiv_rn(e0, dd);
}

// Boundary-vertex weights on polygon edge i->j from the interpolated mSDF values: gshell_tets.py:353-367.
__device__ __forceinline__ bool boundary_weights(float mi, float mj, float& u0, float& u1, float& D) {
  bool nz = fabsf(fsign(mi) + fsign(mj)) != 2.f;
  float nmj = -mj;
  D = __fadd_rn(mi, nmj);
  nz = nz && (fabsf(D) > kEps12);
  u0 = nz ? __fdiv_rn(nmj, D) : 0.f;
  u1 = nz ? __fdiv_rn(mi, D) : 0.f;
  return nz;
}

// ------------------------------------------------------------------------------------------------
// device-side counters (one cache line of int32 words at the head of the workspace; reset per call)
// ------------------------------------------------------------------------------------------------
struct DevCounters {
  unsigned ticket_scan;      // dynamic CTA ids of the ordered kernels (bucket scan, vertex numbering)
  unsigned ticket_unique;
  unsigned classify_done;    // CTAs of the classification kernel that have flushed their tile counts
  unsigned poly_done;        // CTAs of the polygon kernel that have stored their bucket counts
  unsigned n_valid;          // Fv (true count, even when the record buffer overflowed)
  unsigned n_tri;            // T1
  unsigned n_quad;           // T2
  unsigned work_tri;         // T1 / T2 the surface stages operate on: equal to n_tri / n_quad, or 0 / 0 when
  unsigned work_quad;        //   Fv exceeded the record capacity (the caller re-runs with a larger workspace)
  unsigned n_verts;          // V
  unsigned bucket[6];        // polygons per faces_aug bucket
  unsigned poly_done2;       // the same two for the replayed mSDF cut of a cloth / body pair
  unsigned bucket2[6];
  unsigned pad[5];
  unsigned trace_frame;      // diagnostics (d3h_trace_*): row of the trace table this call writes to
  unsigned pad2;
  unsigned long long* trace; // diagnostics: device trace table or nullptr; set by prepare_kernel, not reset
};
constexpr int kCounterWordsReset = 28;  // words of DevCounters that prepare_kernel zeroes
static_assert(sizeof(DevCounters) == 128, "DevCounters is one 128-byte line");

}  // namespace d3h

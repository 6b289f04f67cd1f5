// C-ABI entry points of libd3h_tets.so (declared in include/d3h_tets.h) and the workspace carving.
#include <chrono>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstring>

#include "d3h_internal.cuh"

namespace d3h {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

// ---- profiler -------------------------------------------------------------------------------------
static bool g_prof_on = false;
struct ProfRec { int kind; cudaEvent_t a, b; };
static ProfRec g_prof[4096];
static int g_prof_n = 0;

ProfScope::ProfScope(int kind, cudaStream_t st) : slot(-1), stream(st) {
  if (!g_prof_on || g_prof_n >= 4096) return;
  slot = g_prof_n++;
  ProfRec& r = g_prof[slot];
  r.kind = kind;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, stream);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof[slot].b, stream);
}

static inline int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }

Workspace carve_workspace(void* base, int64_t n_tets, int64_t n_grid, int64_t cap_valid_tets) {
  Workspace ws;
  memset(&ws, 0, sizeof(ws));
  char* p = reinterpret_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t bytes) -> void* {
    void* r = p ? (p + off) : nullptr;
    off += align256(bytes > 0 ? bytes : 1);
    return r;
  };
  const int64_t cap = cap_valid_tets > 0 ? cap_valid_tets : 0;
  const int64_t capc = 4 * cap;
  ws.cap_tets = cap;
  ws.cap_corners = capc;
  const int64_t nchunks = (n_tets + kChunkTets - 1) / kChunkTets;
  const int64_t nwords_f = nchunks * kClassifyItems;
  ws.ntiles_compact = (n_tets + kTileTets - 1) / kTileTets;
  ws.ngroups = capc / kSortGroup + 1;
  ws.ntiles_poly = (cap + kPolyThreads - 1) / kPolyThreads;
  ws.msd_bins = ((n_grid - 1) >> msd_shift_for(n_grid)) + 1;
  ws.nscan_ctas = (ws.msd_bins + kScanThreads - 1) / kScanThreads;
  const int64_t nwords = (n_grid + 31) / 32 + 1;
  ws.ctr = reinterpret_cast<DevCounters*>(take(sizeof(DevCounters)));
  ws.counts = reinterpret_cast<d3h_counts*>(take(sizeof(d3h_counts)));
  ws.occ_bits = reinterpret_cast<unsigned*>(take(nwords * 4));
  ws.mocc_bits = reinterpret_cast<unsigned*>(take(nwords * 4));
  ws.m1_words = reinterpret_cast<unsigned*>(take((nwords_f + kCompactThreads) * 4));
  ws.m2_words = reinterpret_cast<unsigned*>(take((nwords_f + kCompactThreads) * 4));
  ws.tile_cnt = reinterpret_cast<unsigned*>(take((ws.ntiles_compact + 1) * 4));
  ws.records = reinterpret_cast<d3h_tet_record*>(take(cap * (int64_t)sizeof(d3h_tet_record)));
  ws.keys = reinterpret_cast<unsigned long long*>(take(capc * 8));
  ws.vals = reinterpret_cast<unsigned*>(take(capc * 4));
  ws.keys2 = reinterpret_cast<unsigned long long*>(take(capc * 8));
  ws.vals2 = reinterpret_cast<unsigned*>(take(capc * 4));
  ws.keys_scratch = reinterpret_cast<unsigned long long*>(take(2 * capc * 8));
  ws.vals_scratch = reinterpret_cast<unsigned*>(take(2 * capc * 4));
  ws.msd_hist = reinterpret_cast<unsigned*>(take((ws.msd_bins + 12) * 4));
  ws.msd_fill = reinterpret_cast<unsigned*>(take((ws.msd_bins + 8) * 4));
  ws.msd_base = reinterpret_cast<unsigned*>(take((ws.msd_bins + 8) * 4));
  ws.st_scan = reinterpret_cast<unsigned long long*>(take(ws.nscan_ctas * 8));
  ws.group_start = reinterpret_cast<unsigned*>(take((ws.ngroups + 2) * 4));
  ws.st_unique = reinterpret_cast<unsigned long long*>(take(ws.ngroups * 8));
  ws.st_ublock = reinterpret_cast<unsigned long long*>(take((ws.ngroups / 256 + 1) * 8));
  ws.poly_cnt = reinterpret_cast<unsigned*>(take((ws.ntiles_poly + 1) * 32));
  ws.poly_excl = reinterpret_cast<unsigned*>(take((ws.ntiles_poly + 1) * 32));
  ws.vert = reinterpret_cast<float4*>(take(capc * 16));
  ws.acc = reinterpret_cast<float*>(take(capc * 32));
  ws.owner = reinterpret_cast<int32_t*>(take(capc * 4));
  ws.total_bytes = off;
  return ws;
}

// Device-visible alias of the caller's pinned counts buffer, or nullptr when the pointer is not mapped host memory
// (then the counts are copied with cudaMemcpyAsync at the end of the call instead).  One driver query per new pointer.
d3h_counts* mapped_counts_pointer(d3h_counts* host) {
  static thread_local d3h_counts* last_host = nullptr;
  static thread_local d3h_counts* last_dev = nullptr;
  if (host == nullptr) return nullptr;
  if (host == last_host) return last_dev;
  cudaPointerAttributes at;
  d3h_counts* dev = nullptr;
  if (cudaPointerGetAttributes(&at, host) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer != nullptr)
    dev = reinterpret_cast<d3h_counts*>(at.devicePointer);
  else
    cudaGetLastError();  // plain pageable memory: clear the sticky error of the query
  last_host = host;
  last_dev = dev;
  return dev;
}

static int check_forward_args(const d3h_forward_args* a, const char* who) {
  if (!a) { set_error("%s: null argument struct", who); return D3H_E_BADARG; }
  if (a->n_grid <= 0 || a->n_grid >= (1ll << 31) || a->n_tets < 0 || a->n_tets >= (1ll << 31)) {
    set_error("%s: n_grid=%lld / n_tets=%lld outside [1,2^31) / [0,2^31)", who, (long long)a->n_grid, (long long)a->n_tets);
    return D3H_E_BADARG;
  }
  if (!a->pos || !a->sdf || !a->msdf || (!a->tets && a->n_tets > 0) || !a->workspace) {
    set_error("%s: null input / workspace pointer", who);
    return D3H_E_BADARG;
  }
  if (reinterpret_cast<uintptr_t>(a->tets) & 15) { set_error("%s: tets must be 16-byte aligned", who); return D3H_E_BADARG; }
  if (reinterpret_cast<uintptr_t>(a->workspace) & 255) { set_error("%s: workspace must be 256-byte aligned", who); return D3H_E_BADARG; }
  if (a->tet_begin < 0 || a->tet_end < a->tet_begin || a->tet_end > a->n_tets) {
    set_error("%s: tet range [%lld,%lld) outside [0,%lld)", who, (long long)a->tet_begin, (long long)a->tet_end, (long long)a->n_tets);
    return D3H_E_BADARG;
  }
  if (a->cap_valid_tets < 0 || a->cap_valid_tets > (1ll << 27) || a->cap_verts < 0 || a->cap_verts_aug < 0 ||
      a->cap_faces_wt < 0 || a->cap_faces_aug < 0) {
    set_error("%s: capacities must be >= 0 and cap_valid_tets <= 2^27", who);
    return D3H_E_BADARG;
  }
  if (a->cap_valid_tets > 0 && (!a->tape_corners || !a->tape_slots || !a->tape_runs)) {
    set_error("%s: tape_corners / tape_slots (4*cap_valid_tets int32) and tape_runs (cap_verts+1 int32) are required", who);
    return D3H_E_BADARG;
  }
  if ((reinterpret_cast<uintptr_t>(a->zero_g_pos) | reinterpret_cast<uintptr_t>(a->zero_g_sdf) |
       reinterpret_cast<uintptr_t>(a->zero_g_msdf)) & 15) {
    set_error("%s: zero_g_* buffers must be 16-byte aligned", who);
    return D3H_E_BADARG;
  }
  if ((a->cap_verts > 0 && (!a->verts_wt || !a->v_tng_wt || !a->msdf_wt || !a->tape_edges)) ||
      (a->cap_verts_aug > 0 && (!a->verts_aug || !a->v_tng_aug || !a->msdf_aug)) ||
      (a->cap_faces_wt > 0 && !a->faces_wt) || (a->cap_faces_aug > 0 && !a->faces_aug)) {
    set_error("%s: an output pointer is null while its capacity is > 0", who);
    return D3H_E_BADARG;
  }
  const int64_t need = carve_workspace(nullptr, a->n_tets, a->n_grid, a->cap_valid_tets).total_bytes;
  if (a->workspace_bytes < need) {
    set_error("%s: workspace has %lld bytes, %lld needed", who, (long long)a->workspace_bytes, (long long)need);
    return D3H_E_SMALLWS;
  }
  return D3H_OK;
}

static int finish(const char* who, const d3h_forward_args* a, const Workspace& ws, cudaStream_t stream) {
  if (a->zero_g_pos || a->zero_g_sdf || a->zero_g_msdf)
    launch_zero_grads(a->zero_g_pos, a->zero_g_sdf, a->zero_g_msdf, a->n_grid, stream);
  if (a->counts_host && mapped_counts_pointer(a->counts_host) == nullptr)  // not device-mapped: copy at the end
    cudaMemcpyAsync(a->counts_host, ws.counts, sizeof(d3h_counts), cudaMemcpyDeviceToHost, stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("%s: %s", who, cudaGetErrorString(e)); return D3H_E_CUDA; }
  return D3H_OK;
}

}  // namespace d3h

using namespace d3h;

extern "C" int d3h_version(void) { return D3H_VERSION; }
extern "C" const char* d3h_last_error_string(void) { return g_error; }

extern "C" int64_t d3h_workspace_bytes(int64_t n_tets, int64_t n_grid, int64_t cap_valid_tets) {
  if (n_tets < 0 || n_grid <= 0 || cap_valid_tets < 0) return D3H_E_BADARG;
  return carve_workspace(nullptr, n_tets, n_grid, cap_valid_tets).total_bytes;
}
extern "C" int64_t d3h_backward_workspace_bytes(int64_t n_verts) { (void)n_verts; return 0; }

extern "C" int d3h_wait_counts(const d3h_counts* counts_host, int64_t seq, int64_t timeout_us) {
  if (!counts_host) { set_error("d3h_wait_counts: null pointer"); return D3H_E_BADARG; }
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
      const auto dt = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
      if (dt > timeout_us) { set_error("d3h_wait_counts: seq %lld not published after %lld us", (long long)seq, (long long)dt); return D3H_E_TIMEOUT; }
    }
  }
}

extern "C" int d3h_extract_forward(const d3h_forward_args* a, d3h_stream_t s) {
  int rc = check_forward_args(a, "d3h_extract_forward");
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)s;
  Workspace ws = carve_workspace(a->workspace, a->n_tets, a->n_grid, a->cap_valid_tets);
  launch_prepare(*a, ws, stream);
  launch_classify(*a, ws, ws.records, ws.cap_tets, /*emit_keys=*/true, stream);
  launch_edge_sort(*a, ws, stream);
  launch_surface(*a, ws, ws.records, stream);
  return finish("d3h_extract_forward", a, ws, stream);
}

// stage 1 only: prepare + classify of [tet_begin, tet_end); the compact records land in `records_out`
// (class ranks are local to the range) and the counts so far (n_valid/n_tri/n_quad) in `counts_dev_out`.
__global__ void export_range_counts_kernel(const DevCounters* ctr, d3h_counts* out, int64_t seq) {
  memset(out, 0, sizeof(d3h_counts));
  out->n_valid_tets = ctr->n_valid;
  out->n_tri_tets = ctr->n_tri;
  out->n_quad_tets = ctr->n_quad;
  out->n_corners = 3ll * ctr->n_tri + 4ll * ctr->n_quad;
  out->overflow = (ctr->n_valid != ctr->work_tri + ctr->work_quad) ? 1 : 0;
  out->seq = seq;
}

extern "C" int d3h_classify_range(const d3h_forward_args* a, d3h_tet_record* records_out, int64_t cap_records,
                                  d3h_counts* counts_dev_out, d3h_stream_t s) {
  int rc = check_forward_args(a, "d3h_classify_range");
  if (rc) return rc;
  if ((!records_out && cap_records > 0) || cap_records < 0 || !counts_dev_out) {
    set_error("d3h_classify_range: null records / counts pointer");
    return D3H_E_BADARG;
  }
  cudaStream_t stream = (cudaStream_t)s;
  Workspace ws = carve_workspace(a->workspace, a->n_tets, a->n_grid, a->cap_valid_tets);
  launch_prepare(*a, ws, stream);
  launch_classify(*a, ws, records_out, cap_records, /*emit_keys=*/false, stream);
  export_range_counts_kernel<<<1, 1, 0, stream>>>(ws.ctr, counts_dev_out, a->seq);
  if (a->counts_host) cudaMemcpyAsync(a->counts_host, counts_dev_out, sizeof(d3h_counts), cudaMemcpyDeviceToHost, stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("d3h_classify_range: %s", cudaGetErrorString(e)); return D3H_E_CUDA; }
  return D3H_OK;
}

// stage 2 only: `records` is the rank-order concatenation of the shards' records (= global tet order);
// n_tri_tets + n_quad_tets records.  Class ranks are recomputed over the concatenation.
extern "C" int d3h_extract_from_records(const d3h_forward_args* a, const d3h_tet_record* records, int64_t n_tri_tets,
                                        int64_t n_quad_tets, d3h_stream_t s) {
  int rc = check_forward_args(a, "d3h_extract_from_records");
  if (rc) return rc;
  const int64_t n = n_tri_tets + n_quad_tets;
  if (n < 0 || n_tri_tets < 0 || n_quad_tets < 0 || n > a->cap_valid_tets || (!records && n > 0)) {
    set_error("d3h_extract_from_records: %lld records do not fit cap_valid_tets=%lld", (long long)n, (long long)a->cap_valid_tets);
    return D3H_E_BADARG;
  }
  cudaStream_t stream = (cudaStream_t)s;
  Workspace ws = carve_workspace(a->workspace, a->n_tets, a->n_grid, a->cap_valid_tets);
  d3h_forward_args b = *a;
  b.tet_begin = b.tet_end = 0;  // prepare still resets the scan state; no tets are classified here
  launch_prepare(b, ws, stream);
  if (n > 0) cudaMemcpyAsync(ws.records, records, n * sizeof(d3h_tet_record), cudaMemcpyDeviceToDevice, stream);
  launch_rank_records(*a, ws, ws.records, n, stream);
  launch_edge_sort(*a, ws, stream);
  launch_surface(*a, ws, ws.records, stream);
  return finish("d3h_extract_from_records", a, ws, stream);
}

extern "C" int d3h_extract_backward(const d3h_backward_args* a, d3h_stream_t s) {
  if (!a) { set_error("d3h_extract_backward: null argument struct"); return D3H_E_BADARG; }
  if (a->n_grid <= 0 || a->n_verts < 0 || a->n_tri_tets < 0 || a->n_quad_tets < 0 || !a->g_pos || !a->g_sdf ||
      !a->pos || !a->sdf || !a->msdf) {
    set_error("d3h_extract_backward: null pointer or negative size");
    return D3H_E_BADARG;
  }
  if (a->n_verts > 0 && (!a->tape_edges || !a->tape_corners || !a->tape_slots || !a->tape_runs || !a->verts_wt || !a->msdf_wt)) {
    set_error("d3h_extract_backward: tape / saved outputs missing");
    return D3H_E_BADARG;
  }
  if ((reinterpret_cast<uintptr_t>(a->g_pos) | reinterpret_cast<uintptr_t>(a->g_sdf) |
       reinterpret_cast<uintptr_t>(a->g_msdf)) & 15) {
    set_error("d3h_extract_backward: gradient buffers must be 16-byte aligned");
    return D3H_E_BADARG;
  }
  launch_backward(*a, (cudaStream_t)s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("d3h_extract_backward: %s", cudaGetErrorString(e)); return D3H_E_CUDA; }
  return D3H_OK;
}

// ---- diagnostics: per-kernel device time, measured with CUDA events on the launching stream -----------------------
static const char* kKernelNames[K_COUNT] = {"prepare", "classify", "compact", "bucket_scan", "partition", "unique",
                                            "poly_faces", "poly_cut", "zero", "adjoint", "rank_records"};
extern "C" int d3h_profile_enable(int on) {
  g_prof_on = on != 0;
  return D3H_OK;
}
extern "C" int d3h_profile_kinds(void) { return K_COUNT; }
extern "C" const char* d3h_profile_kernel_name(int k) { return (k >= 0 && k < K_COUNT) ? kKernelNames[k] : ""; }
// Synchronises the recorded events, adds their durations (ms) and launch counts per kernel kind, then clears the log.
extern "C" int d3h_profile_read(float* ms_by_kind, int* launches_by_kind) {
  if (!ms_by_kind || !launches_by_kind) { set_error("d3h_profile_read: null output"); return D3H_E_BADARG; }
  for (int k = 0; k < K_COUNT; ++k) { ms_by_kind[k] = 0.f; launches_by_kind[k] = 0; }
  int rc = D3H_OK;
  for (int i = 0; i < g_prof_n; ++i) {
    ProfRec& r = g_prof[i];
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.a, r.b);
    if (e != cudaSuccess) { set_error("d3h_profile_read: %s", cudaGetErrorString(e)); rc = D3H_E_CUDA; }
    ms_by_kind[r.kind] += ms;
    launches_by_kind[r.kind] += 1;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof_n = 0;
  return rc;
}

// ---- test hook: the case tables, from the same initializer macros the __constant__ copies are built from --------
// which: 0 num_tri[16], 1 loop_edge[16][4], 2 tri_edge[16][6], 3 cut_tri[8][6], 4 num_cut_tri[8],
//        5 cut_quad[16][12], 6 num_cut_quad[16], 7 edge_p[6], 8 edge_q[6].  Returns the element count.  Host only.
extern "C" int d3h_debug_table(int which, int8_t* out, int cap) {
  static const int8_t t0[16] = D3H_T_NUM_TRI;
  static const int8_t t1[16][4] = D3H_T_LOOP_EDGE;
  static const int8_t t2[16][6] = D3H_T_TRI_EDGE;
  static const int8_t t3[8][6] = D3H_T_CUT_TRI;
  static const int8_t t4[8] = D3H_T_NUM_CUT_TRI;
  static const int8_t t5[16][12] = D3H_T_CUT_QUAD;
  static const int8_t t6[16] = D3H_T_NUM_CUT_QUAD;
  static const int8_t t7[6] = D3H_T_EDGE_P;
  static const int8_t t8[6] = D3H_T_EDGE_Q;
  const void* src[9] = {t0, t1, t2, t3, t4, t5, t6, t7, t8};
  const int len[9] = {16, 64, 96, 48, 8, 192, 16, 6, 6};
  if (which < 0 || which > 8) { set_error("d3h_debug_table: unknown table %d", which); return D3H_E_BADARG; }
  if (cap < len[which] || !out) { set_error("d3h_debug_table: buffer too small"); return D3H_E_BADARG; }
  memcpy(out, src[which], len[which]);
  return len[which];
}

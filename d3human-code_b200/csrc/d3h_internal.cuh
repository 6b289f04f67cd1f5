// Internal (non-ABI) declarations shared by the translation units of libd3h_tets.so.
#pragma once

#include "d3h_common.cuh"

namespace d3h {

// ---- tile shapes ----------------------------------------------------------------------------------
constexpr int kClassifyThreads = 256;
constexpr int kClassifyItems = 8;  // tets per thread: 8 x 16 B loads in flight
constexpr int kClassifyTile = kClassifyThreads * kClassifyItems;

constexpr int kSortThreads = 256;
constexpr int kSortItems = 4;
constexpr int kSortTile = kSortThreads * kSortItems;
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kMaxPasses = 8;

constexpr int kRleThreads = 256;
constexpr int kRleItems = 4;
constexpr int kRleTile = kRleThreads * kRleItems;

constexpr int kPolyThreads = 256;  // one valid tet (= one polygon) per thread

// ---- workspace ------------------------------------------------------------------------------------
// Every region is 256-byte aligned.  Sizes depend on (F, N, cap_valid_tets) only.
struct Workspace {
  DevCounters* ctr;               // 128 B
  d3h_counts* counts;             // device copy of the public counts
  unsigned* occ_bits;             // ceil(N/32) words: sdf > 0
  unsigned* mocc_bits;            // ceil(N/32) words: (+-)msdf > 0 (open-mesh prefilter only)
  unsigned long long* st_classify;// one status word per classify tile
  d3h_tet_record* records;        // cap_valid_tets
  unsigned long long* keys[2];    // 4*cap_valid_tets each
  unsigned* vals[2];              // 4*cap_valid_tets each
  unsigned* radix_hist;           // kMaxPasses * 256
  unsigned* st_sort;              // kMaxPasses * ntiles_sort * 256
  unsigned long long* st_rle;     // ntiles_rle
  unsigned* st_poly;              // ntiles_poly * 8 (6 used)
  float4* vert;                   // (x,y,z,msdf) per watertight vertex, 4*cap_valid_tets
  float4* tng;                    // (tx,ty,tz,-) per watertight vertex
  float* acc;                     // 8 floats per watertight vertex: normal xyz, tangent xyz, count, pad
  unsigned* polyinfo;             // per valid tet: (bucket rank << 4) | mSDF case
  int64_t cap_tets, cap_corners;
  int64_t ntiles_classify, ntiles_sort, ntiles_rle, ntiles_poly;
  int64_t total_bytes;
};

// Carves `base` (may be nullptr when only the size is wanted).
Workspace carve_workspace(void* base, int64_t n_tets, int64_t n_grid, int64_t cap_valid_tets);

int key_bits_for(int64_t n_grid);  // bits per endpoint in the packed edge key

// ---- stage launchers (all asynchronous on `stream`) -------------------------------------------------
void launch_prepare(const d3h_forward_args& a, const Workspace& ws, cudaStream_t stream);
void launch_classify(const d3h_forward_args& a, const Workspace& ws, d3h_tet_record* records, int64_t cap_records,
                     cudaStream_t stream);
void launch_edge_sort(const d3h_forward_args& a, const Workspace& ws, const d3h_tet_record* records,
                      cudaStream_t stream);
void launch_surface(const d3h_forward_args& a, const Workspace& ws, const d3h_tet_record* records,
                    cudaStream_t stream);
void launch_rank_records(const Workspace& ws, d3h_tet_record* records, int64_t n_records, cudaStream_t stream);
void launch_backward(const d3h_backward_args& a, cudaStream_t stream);

void set_error(const char* fmt, ...);

// ---- optional per-kernel timing (d3h_profile_*): CUDA events recorded on the launching stream around each launch ----
enum KernelKind {
  K_PREPARE = 0, K_CLASSIFY, K_EMIT_KEYS, K_RADIX_PASS, K_RLE_INTERP, K_POLY_FACES, K_VERTEX_FRAME, K_POLY_CUT,
  K_ZERO, K_BOUNDARY_ADJ, K_CROSSING_ADJ, K_RANK_RECORDS, K_COUNT
};
struct ProfScope {
  ProfScope(int kind, cudaStream_t stream);
  ~ProfScope();
  int slot;
  cudaStream_t stream;
};

}  // namespace d3h

// Stage 2: edge de-duplication and vertex numbering.
//
// Replaces gshell_tets.py:277-287: `sort_edges` (:209-217), `torch.unique(all_edges, dim=0, return_inverse=True)`
// (:279), the crossing mask (:282) and `mapping`/`idx_map`/`interp_v` (:283-287) -- and, fused into the run-length
// kernel, the zero-crossing interpolation of :291-303.
//
// Only *crossing* edges are sorted: a non-crossing edge maps to -1 in the reference and is never read again, and the
// rank of a crossing edge among crossing unique edges in (min,max) lexicographic order does not depend on the others.
// Each valid tet contributes its polygon corners (3 or 4 crossing edges, in mesh_edge_table order), so the inverse map
// of the sort *is* the polygon corner array, laid out [3*T1 | 4*T2] like the boundary vertices (:406-407).
//
// The sort is an MSD radix sort with a block-local finish (keys are (min << bits) | max, 2*bits <= 62):
//   partition_kernel  one radix pass on the top <= 17 bits of `min` (32-64 vertex ids per bucket; histogram, bases and groups come from
//                     the compaction kernel).  The pass need not be stable (equal keys are merged afterwards), so slots
//                     are claimed with warp-aggregated atomics -- no inter-tile dependency.
//   local_sort_kernel one CTA per group of whole buckets (<= 4096 keys): bitonic sort of (key, value) in shared memory.
//                     A bucket too large for shared memory (surface concentrated in a few thousand consecutive vertex
//                     ids, e.g. an axis-aligned plane) is sorted by its CTA in global memory instead: slower, still exact.
//   rle_interp_kernel head flags + scan (decoupled look-back) = vertex ids in sorted order; scatters the ids to the
//                     corner array, writes the (a,b) tape and interpolates position / mSDF of every new vertex.
//
// v1 used six chained 8-bit LSD passes (look-back per digit and tile): 11 us per pass under ncu for 192k keys
// (profiles/r01a_launches_v1.csv) -- the chains, not the data volume, set the time.
#include "d3h_internal.cuh"

namespace d3h {

int key_bits_for(int64_t n_grid) {
  int b = 1;
  while ((1ll << b) < n_grid) ++b;
  return b;
}
int msd_shift_for(int64_t n_grid) {
  const int b = key_bits_for(n_grid);
  return b > kMsdBits ? b - kMsdBits : 0;
}

// ------------------------------------------------------------------------------------------------
// MSD partition
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
partition_kernel(const unsigned long long* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
                 unsigned long long* __restrict__ keys_out, unsigned* __restrict__ vals_out,
                 const DevCounters* __restrict__ ctr, const unsigned* __restrict__ base, unsigned* __restrict__ fill,
                 int digit_shift) {
  const int64_t ncorn = 3ll * ctr->work_tri + 4ll * ctr->work_quad;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = i < ncorn;
  unsigned long long key = 0;
  unsigned val = 0;
  int bin = -1 - (int)lane_id();  // idle lanes match nobody
  if (ok) {
    key = keys_in[i];
    val = vals_in[i];
    bin = (int)(key >> digit_shift);
  }
  const unsigned peers = __match_any_sync(0xffffffffu, bin);
  if (!ok) return;
  const int leader = __ffs(peers) - 1;
  unsigned pos = 0;
  if ((int)lane_id() == leader) pos = __ldg(base + bin) + atomicAdd(&fill[bin], (unsigned)__popc(peers));
  pos = __shfl_sync(peers, pos, leader) + __popc(peers & lanemask_lt());
  keys_out[pos] = key;
  vals_out[pos] = val;
}

// ------------------------------------------------------------------------------------------------
// block-local finish
// ------------------------------------------------------------------------------------------------
// Group g owns key positions [snap(g*G), snap((g+1)*G)) where snap(x) = start of the bucket that contains position x:
// groups are unions of whole buckets, tile [0,P) exactly, and hold < G + (largest bucket) keys.
template <typename KeyPtr, typename ValPtr>
__device__ __forceinline__ void bitonic_sort_block(KeyPtr k, ValPtr v, unsigned npow2) {
  for (unsigned size = 2; size <= npow2; size <<= 1) {
    for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
      for (unsigned t = threadIdx.x; t < (npow2 >> 1); t += blockDim.x) {
        const unsigned lo = 2 * t - (t & (stride - 1));
        const unsigned hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long ka = k[lo], kb = k[hi];
        if ((ka > kb) == up) {
          const unsigned va = v[lo], vb = v[hi];
          k[lo] = kb; k[hi] = ka;
          v[lo] = vb; v[hi] = va;
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kLocalSortThreads)
local_sort_kernel(unsigned long long* keys, unsigned* vals, unsigned long long* scratch_keys, unsigned* scratch_vals,
                  const DevCounters* __restrict__ ctr, const unsigned* __restrict__ group_start) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* s_key = reinterpret_cast<unsigned long long*>(smem_raw);
  unsigned* s_val = reinterpret_cast<unsigned*>(s_key + kLocalSortCap);

  const int64_t ncorn = 3ll * ctr->work_tri + 4ll * ctr->work_quad;
  const int64_t g = blockIdx.x;
  if (g * kSortGroup >= ncorn) return;
  const unsigned lo = __ldg(group_start + g);
  const unsigned hi = ((g + 1) * kSortGroup >= ncorn) ? (unsigned)ncorn : __ldg(group_start + g + 1);
  if (hi <= lo) return;  // this group's positions belong to a bucket that started in an earlier group
  const unsigned n = hi - lo;
  unsigned npow2 = 2;
  while (npow2 < n) npow2 <<= 1;

  if (n <= (unsigned)kLocalSortCap) {
    for (unsigned i = threadIdx.x; i < npow2; i += kLocalSortThreads) {
      s_key[i] = (i < n) ? keys[lo + i] : ~0ull;
      s_val[i] = (i < n) ? vals[lo + i] : 0u;
    }
    __syncthreads();
    bitonic_sort_block(s_key, s_val, npow2);
    for (unsigned i = threadIdx.x; i < n; i += kLocalSortThreads) {
      keys[lo + i] = s_key[i];
      vals[lo + i] = s_val[i];
    }
  } else {
    // oversized bucket: same network on a padded copy in global scratch (exact, slower; only degenerate inputs)
    unsigned long long* gk = scratch_keys + 2ull * lo;  // padded copies of disjoint ranges cannot overlap at 2*lo
    unsigned* gv = scratch_vals + 2ull * lo;
    for (unsigned i = threadIdx.x; i < npow2; i += kLocalSortThreads) {
      gk[i] = (i < n) ? keys[lo + i] : ~0ull;
      gv[i] = (i < n) ? vals[lo + i] : 0u;
    }
    __syncthreads();
    bitonic_sort_block(gk, gv, npow2);
    for (unsigned i = threadIdx.x; i < n; i += kLocalSortThreads) {
      keys[lo + i] = gk[i];
      vals[lo + i] = gv[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// run-length + numbering + zero-crossing interpolation
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long kRFlagAgg = 1ull << 62, kRFlagInc = 2ull << 62, kRValMask = (1ull << 62) - 1;

__global__ void __launch_bounds__(kRleThreads)
rle_interp_kernel(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ vals,
                  DevCounters* __restrict__ ctr, unsigned long long* __restrict__ status, int key_bits,
                  const float* __restrict__ pos, const float* __restrict__ sdf, const float* __restrict__ msdf,
                  int msdf_negate, int32_t* __restrict__ tape_corners, int32_t* __restrict__ tape_edges,
                  int64_t cap_verts, int64_t cap_verts_aug, float4* __restrict__ w_vert, float4* __restrict__ w_acc,
                  float* __restrict__ verts_wt, float* __restrict__ msdf_wt, float* __restrict__ verts_aug,
                  float* __restrict__ msdf_aug) {
  constexpr int WARPS = kRleThreads / 32;
  __shared__ unsigned s_tile;
  __shared__ unsigned s_wsum[WARPS];
  __shared__ unsigned long long s_excl;

  const unsigned t1 = ctr->work_tri;
  const int64_t ncorn = 3ll * t1 + 4ll * ctr->work_quad;
  const int64_t ntiles = (ncorn + kRleTile - 1) / kRleTile;
  if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->ticket_rle, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  if ((int64_t)tile >= ntiles) return;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;

  // blocked: thread owns kRleItems consecutive sorted keys
  const int64_t i0 = (int64_t)tile * kRleTile + (int64_t)threadIdx.x * kRleItems;
  unsigned long long k[kRleItems];
  unsigned long long prev = (i0 > 0 && i0 - 1 < ncorn) ? keys[i0 - 1] : ~0ull;
  bool head[kRleItems];
  unsigned nhead = 0;
#pragma unroll
  for (int j = 0; j < kRleItems; ++j) {
    const int64_t idx = i0 + j;
    const bool ok = idx < ncorn;
    k[j] = ok ? keys[idx] : ~0ull;
    head[j] = ok && (idx == 0 || k[j] != prev);
    prev = k[j];
    nhead += head[j];
  }
  unsigned incl = nhead;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (unsigned)o) incl += n;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  unsigned wpre = 0, total = 0;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) {
    if (w < (int)warp) wpre += s_wsum[w];
    total += s_wsum[w];
  }
  if (warp == 0) {
    unsigned long long excl_tiles = 0ull;
    if (tile == 0) {
      if (lane == 0) st_relaxed_u64(status, kRFlagInc | total);
    } else {
      if (lane == 0) st_relaxed_u64(status + tile, kRFlagAgg | total);
      int64_t look = (int64_t)tile - 1;
      while (true) {
        const int64_t idx = look - lane;
        unsigned long long w = kRFlagInc;
        if (idx >= 0) {
          do { w = ld_relaxed_u64(status + idx); } while ((w >> 62) == 0ull);
        }
        const unsigned inc_mask = __ballot_sync(0xffffffffu, (w >> 62) == 2ull);
        const int first = inc_mask ? (__ffs(inc_mask) - 1) : 32;
        unsigned long long contrib = ((int)lane <= first) ? (w & kRValMask) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl_tiles += contrib;
        if (inc_mask) break;
        look -= 32;
      }
      if (lane == 0) st_relaxed_u64(status + tile, kRFlagInc | (excl_tiles + total));
    }
    if (lane == 0) {
      s_excl = excl_tiles;
      if ((int64_t)tile == ntiles - 1) ctr->n_verts = (unsigned)(excl_tiles + total);
    }
  }
  __syncthreads();

  int64_t vid = (int64_t)s_excl + wpre + (incl - nhead) - 1;  // id of the run that precedes this thread's keys
  const unsigned long long bmask = (1ull << key_bits) - 1;
#pragma unroll
  for (int j = 0; j < kRleItems; ++j) {
    const int64_t idx = i0 + j;
    if (idx >= ncorn) break;
    if (head[j]) {
      ++vid;
      const int a = (int)(k[j] >> key_bits), b = (int)(k[j] & bmask);
      // zero-crossing interpolation, gshell_tets.py:291-303 (op order: SURVEY A.4)
      float w0, w1, dd;
      crossing_weights(__ldg(sdf + a), __ldg(sdf + b), w0, w1, dd);
      float ma = __ldg(msdf + a), mb = __ldg(msdf + b);
      if (msdf_negate) { ma = -ma; mb = -mb; }
      const float x = lerp2(__ldg(pos + 3ll * a + 0), w0, __ldg(pos + 3ll * b + 0), w1);
      const float y = lerp2(__ldg(pos + 3ll * a + 1), w0, __ldg(pos + 3ll * b + 1), w1);
      const float z = lerp2(__ldg(pos + 3ll * a + 2), w0, __ldg(pos + 3ll * b + 2), w1);
      const float m = lerp2(ma, w0, mb, w1);
      w_vert[vid] = make_float4(x, y, z, m);
      w_acc[2 * vid] = make_float4(0.f, 0.f, 0.f, 0.f);
      w_acc[2 * vid + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vid < cap_verts) {
        tape_edges[2 * vid] = a;
        tape_edges[2 * vid + 1] = b;
        verts_wt[3 * vid] = x; verts_wt[3 * vid + 1] = y; verts_wt[3 * vid + 2] = z;
        msdf_wt[vid] = m;
      }
      if (vid < cap_verts_aug) {
        // rows of verts_aug not referenced by faces_aug are zero (gshell_tets.py:423-427); a watertight vertex is
        // referenced iff its mSDF is positive (every cut case keeps exactly the positive corners)
        const bool used = m > 0.f;
        verts_aug[3 * vid] = used ? x : 0.f;
        verts_aug[3 * vid + 1] = used ? y : 0.f;
        verts_aug[3 * vid + 2] = used ? z : 0.f;
        msdf_aug[vid] = m;
      }
    }
    // value = (class, 4*class_rank + k) -> corner slot in the [3*T1 | 4*T2] layout
    const unsigned val = vals[idx];
    const unsigned r4 = val & 0x7fffffffu;
    const int64_t slot = (val >> 31) ? (3ll * t1 + r4) : (3ll * (r4 >> 2) + (r4 & 3u));
    tape_corners[slot] = (int32_t)vid;
  }
}

// ------------------------------------------------------------------------------------------------
void launch_edge_sort(const d3h_forward_args& a, const Workspace& ws, cudaStream_t stream) {
  const int key_bits = key_bits_for(a.n_grid);
  const int64_t capc = ws.cap_corners;
  if (capc <= 0) return;
  {
    ProfScope ps(K_PARTITION, stream);
    partition_kernel<<<(unsigned)((capc + 255) / 256), 256, 0, stream>>>(ws.keys, ws.vals, ws.keys2, ws.vals2, ws.ctr,
                                                                          ws.msd_base, ws.msd_fill,
                                                                          key_bits + msd_shift_for(a.n_grid));
  }
  {
    static bool attr_set = false;
    const int smem = kLocalSortCap * 12;
    if (!attr_set) {
      cudaFuncSetAttribute(local_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      attr_set = true;
    }
    ProfScope ps(K_LOCAL_SORT, stream);
    local_sort_kernel<<<(unsigned)(capc / kSortGroup + 1), kLocalSortThreads, smem, stream>>>(
        ws.keys2, ws.vals2, ws.keys_scratch, ws.vals_scratch, ws.ctr, ws.group_start);
  }
  ProfScope ps(K_RLE_INTERP, stream);
  rle_interp_kernel<<<(unsigned)ws.ntiles_rle, kRleThreads, 0, stream>>>(
      ws.keys2, ws.vals2, ws.ctr, ws.st_rle, key_bits, a.pos, a.sdf, a.msdf, a.msdf_negate, a.tape_corners, a.tape_edges,
      a.cap_verts, a.cap_verts_aug, ws.vert, reinterpret_cast<float4*>(ws.acc), a.verts_wt, a.msdf_wt, a.verts_aug,
      a.msdf_aug);
}

}  // namespace d3h

// Stage 2: edge de-duplication and vertex numbering.
//
// Two paths produce the same vertex ids (rank of a crossing edge in (min,max) lexicographic order):
//   general  this file's radix sort + run-length scan of the crossing-edge keys of the call (first call on a grid,
//            tet-range shards, D3H_STATIC_EDGES=0);
//   static   edge_emit_kernel below: the grid's tet array does not change during a run (hmsdf.py:207-212), so the sorted
//            list of ALL tet edges is built once and a call only marks its crossing edges in a bitmap over that list.
//
// Replaces gshell_tets.py:277-287: `sort_edges` (:209-217), `torch.unique(all_edges, dim=0, return_inverse=True)`
// (:279), the crossing mask (:282) and `mapping`/`idx_map`/`interp_v` (:283-287) -- and, fused into the same kernel,
// the zero-crossing interpolation of :291-303.
//
// Only *crossing* edges are sorted: a non-crossing edge maps to -1 in the reference and is never read again, and the
// rank of a crossing edge among crossing unique edges in (min,max) lexicographic order does not depend on the others.
// Each valid tet contributes its polygon corners (3 or 4 crossing edges, in mesh_edge_table order), so the inverse map
// of the sort *is* the polygon corner array, laid out [3*T1 | 4*T2] like the boundary vertices (:406-407).
//
// The sort is an MSD radix sort with a block-local finish (keys are (min << bits) | max, 2*bits <= 62):
//   bucket_scan_kernel  exclusive scan of the MSD histogram (<= 2^17 buckets of 16-64 vertex ids, filled by the
//                       compaction kernel): one bucket per thread, CTA aggregates chained through status words; also
//                       snaps the sort groups to bucket boundaries.
//   partition_kernel    one radix pass: scatters (key, value) to its bucket.  The pass need not be stable (equal keys
//                       are merged afterwards), so slots are claimed with warp-aggregated atomics.
//   group_sort_kernel   one CTA per group of whole buckets (~256 keys): bitonic sort of (key, value) in registers /
//                       shuffles / shared memory, in place; counts the group's distinct keys.  A group too large for
//                       shared memory (surface concentrated in a few consecutive vertex ids) is sorted in global
//                       memory instead: slower, still exact.
//   vertex_emit_kernel  head flags + block scan + (sum of the earlier groups' counts) = vertex ids in sorted order;
//                       scatters the ids to the corner array, writes the (a,b) tape and the per-vertex corner runs the
//                       backward pass gathers over, and interpolates position / mSDF of every new vertex.
//
// v1 used six chained 8-bit LSD passes (11 us per pass for 192k keys: the chains set the time); v2 sorted 1024-4096 key
// groups with 512 threads (40 us) and ran the run-length pass as a separate look-back kernel (13 us); v3a fused sort
// and numbering in one kernel whose CTAs waited on each other's counts (look-back, then direct polling): 40 % of its
// samples sat on that wait, so the count hand-over is now a kernel boundary (0.7 us inside a CUDA graph).
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

constexpr unsigned long long kFlagAgg = 1ull << 62, kValMask = (1ull << 62) - 1;

// ------------------------------------------------------------------------------------------------
// bucket scan
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads)
bucket_scan_kernel(const unsigned* __restrict__ hist, unsigned* __restrict__ base, unsigned* __restrict__ fill,
                   unsigned* __restrict__ group_start, int nbins, DevCounters* __restrict__ ctr,
                   unsigned long long* __restrict__ status) {
  __shared__ unsigned s_w[32];
  __shared__ unsigned s_cta;
  __shared__ unsigned s_excl;
  unsigned long long* tr = trace_begin(ctr->trace, ctr->trace_frame, K_BUCKET_SCAN);
  if (threadIdx.x == 0) s_cta = atomicAdd(&ctr->ticket_scan, 1u);  // CTAs are numbered in start order: no deadlock
  __syncthreads();
  const unsigned cta = s_cta;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  const int bin = (int)(cta * kScanThreads + threadIdx.x);
  const unsigned c = (bin < nbins) ? __ldcg(hist + bin) : 0u;
  unsigned incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (unsigned)o) incl += n;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  unsigned wpre = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    if (w < (int)warp) wpre += s_w[w];
    total += s_w[w];
  }
  if (threadIdx.x == 0) st_relaxed_u64(status + cta, kFlagAgg | total);
  // sum of the aggregates of all earlier CTAs (<= 128 of them: one per thread, no chain)
  unsigned prev = 0;
  for (unsigned j = threadIdx.x; j < cta; j += kScanThreads) {
    unsigned long long w;
    do { w = ld_relaxed_u64(status + j); } while ((w >> 62) == 0ull);
    prev += (unsigned)(w & kValMask);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) prev += __shfl_xor_sync(0xffffffffu, prev, o);
  __syncthreads();  // s_w is reused
  if (lane == 0) s_w[warp] = prev;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned e = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) e += s_w[w];
    s_excl = e;
  }
  __syncthreads();
  const unsigned run = s_excl + wpre + incl - c;
  if (bin < nbins) {
    base[bin] = run;
    fill[bin] = run;
    if (c) {
      for (unsigned g = (run + kSortGroup - 1) / kSortGroup; (uint64_t)g * kSortGroup < (uint64_t)run + c; ++g)
        group_start[g] = run;
    }
    if (bin == nbins - 1) {
      base[nbins] = run + c;  // = P
      group_start[(run + c + kSortGroup - 1) / kSortGroup] = run + c;
    }
  }
  trace_end(tr);
}

// ------------------------------------------------------------------------------------------------
// MSD partition
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
partition_kernel(const unsigned long long* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
                 unsigned long long* __restrict__ keys_out, unsigned* __restrict__ vals_out,
                 const DevCounters* __restrict__ ctr, unsigned* __restrict__ fill, int digit_shift) {
  const int64_t ncorn = 3ll * ctr->work_tri + 4ll * ctr->work_quad;
  unsigned long long* tr = trace_begin(ctr->trace, ctr->trace_frame, K_PARTITION);
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
  if ((int)lane_id() == leader) pos = atomicAdd(&fill[bin], (unsigned)__popc(peers));  // cursor starts at the bucket base
  pos = __shfl_sync(peers, pos, leader) + __popc(peers & lanemask_lt());
  keys_out[pos] = key;
  vals_out[pos] = val;
  trace_end(tr);
}

// ------------------------------------------------------------------------------------------------
// block-local finish: sort + run-length + numbering + zero-crossing interpolation
// ------------------------------------------------------------------------------------------------
// Classic shared/global-memory bitonic network (one barrier per step): only used for oversized groups in global scratch.
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

// Bitonic sort of 256*E (key, value) pairs held E per thread in registers, blocked layout (element e = thread*E + r).
// Strides below E are compare-exchanges between a thread's own registers, strides E..16E are warp shuffles, and only
// the three top strides of the three last stages cross warps through shared memory: 9 barriers instead of one per
// step (55 for 1024 keys), which is where v2/v3a of this kernel spent 42 % of their warp samples.
template <int E>
__device__ __forceinline__ void sort_group_regs(const unsigned long long* __restrict__ gkeys,
                                                const unsigned* __restrict__ gvals, unsigned n,
                                                unsigned long long* s_key, unsigned* s_val) {
  constexpr unsigned NP = kUniqueThreads * E;
  const unsigned t = threadIdx.x;
  unsigned long long k[E];
  unsigned v[E];
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const unsigned e = t * E + r;
    k[r] = (e < n) ? gkeys[e] : ~0ull;
    v[r] = (e < n) ? gvals[e] : 0u;
  }
#pragma unroll 1
  for (unsigned size = 2; size <= NP; size <<= 1) {
    unsigned stride = size >> 1;
    if (stride >= 32u * E) {  // cross-warp strides: through shared memory
#pragma unroll
      for (int r = 0; r < E; ++r) { s_key[t * E + r] = k[r]; s_val[t * E + r] = v[r]; }
      __syncthreads();
      for (; stride >= 32u * E; stride >>= 1) {
        for (unsigned q = t; q < NP / 2; q += kUniqueThreads) {
          const unsigned lo = 2 * q - (q & (stride - 1));
          const unsigned hi = lo + stride;
          const bool up = (lo & size) == 0;
          const unsigned long long ka = s_key[lo], kb = s_key[hi];
          if ((ka > kb) == up) {
            const unsigned va = s_val[lo], vb = s_val[hi];
            s_key[lo] = kb; s_key[hi] = ka;
            s_val[lo] = vb; s_val[hi] = va;
          }
        }
        __syncthreads();
      }
#pragma unroll
      for (int r = 0; r < E; ++r) { k[r] = s_key[t * E + r]; v[r] = s_val[t * E + r]; }
    }
    for (; stride >= (unsigned)E; stride >>= 1) {  // partner = same register of lane ^ (stride / E)
      const unsigned m = stride / E;
      const bool lower = (t & m) == 0;
#pragma unroll
      for (int r = 0; r < E; ++r) {
        const unsigned long long pk = __shfl_xor_sync(0xffffffffu, k[r], m);
        const unsigned pv = __shfl_xor_sync(0xffffffffu, v[r], m);
        const bool up = ((t * E + r) & size) == 0;
        const bool take = (lower == up) ? (pk < k[r]) : (pk > k[r]);  // ties keep their own pair on both sides
        if (take) { k[r] = pk; v[r] = pv; }
      }
    }
#pragma unroll
    for (int sr = E / 2; sr >= 1; sr >>= 1) {  // in-register strides
      if ((unsigned)sr < size) {
#pragma unroll
        for (int r = 0; r < E; ++r) {
          if ((r & sr) == 0) {
            const bool up = ((t * E + r) & size) == 0;
            if ((k[r] > k[r + sr]) == up) {
              const unsigned long long tk = k[r]; k[r] = k[r + sr]; k[r + sr] = tk;
              const unsigned tv = v[r]; v[r] = v[r + sr]; v[r + sr] = tv;
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < E; ++r) { s_key[t * E + r] = k[r]; s_val[t * E + r] = v[r]; }
  __syncthreads();
}

// ---- kernel A: sort every group in place, count its distinct keys ------------------------------------------
__global__ void __launch_bounds__(kUniqueThreads)
group_sort_kernel(unsigned long long* __restrict__ keys, unsigned* __restrict__ vals,
                  unsigned long long* __restrict__ scratch_keys, unsigned* __restrict__ scratch_vals,
                  const DevCounters* __restrict__ ctr, const unsigned* __restrict__ group_start,
                  unsigned* __restrict__ group_heads, unsigned* __restrict__ gblock_heads) {
  __shared__ __align__(16) unsigned long long s_key[kLocalSortCap];
  __shared__ unsigned s_val[kLocalSortCap];
  __shared__ unsigned s_cnt[kUniqueThreads / 32];

  const int64_t ncorn = 3ll * ctr->work_tri + 4ll * ctr->work_quad;
  const unsigned ngroups_used = (unsigned)((ncorn + kSortGroup - 1) / kSortGroup);
  const unsigned g = blockIdx.x;
  unsigned long long* tr = trace_begin(ctr->trace, ctr->trace_frame, K_GROUP_SORT);
  if (g >= ngroups_used) return;
  const unsigned lo = __ldcg(group_start + g);
  const unsigned hi = (g + 1 == ngroups_used) ? (unsigned)ncorn : __ldcg(group_start + g + 1);
  const unsigned n = hi > lo ? hi - lo : 0u;  // 0: this group's positions belong to a bucket that started earlier
  if (n == 0u) {
    if (threadIdx.x == 0) group_heads[g] = 0u;
    return;
  }
  unsigned nhead = 0;
  if (n <= (unsigned)kLocalSortCap) {
    if (n > 4u * kUniqueThreads) sort_group_regs<8>(keys + lo, vals + lo, n, s_key, s_val);
    else if (n > 2u * kUniqueThreads) sort_group_regs<4>(keys + lo, vals + lo, n, s_key, s_val);
    else sort_group_regs<2>(keys + lo, vals + lo, n, s_key, s_val);
    for (unsigned i = threadIdx.x; i < n; i += kUniqueThreads) {
      const unsigned long long k = s_key[i];
      keys[lo + i] = k;
      vals[lo + i] = s_val[i];
      nhead += (i == 0 || k != s_key[i - 1]);  // a group starts on a bucket boundary: i == 0 is always a head
    }
  } else {
    // oversized group: classic network on a padded copy in global scratch (exact, slower; only degenerate inputs)
    unsigned npow2 = 2;
    while (npow2 < n) npow2 <<= 1;
    unsigned long long* gk = scratch_keys + 2ull * lo;  // padded copies of disjoint ranges cannot overlap at 2*lo
    unsigned* gv = scratch_vals + 2ull * lo;
    for (unsigned i = threadIdx.x; i < npow2; i += kUniqueThreads) {
      gk[i] = (i < n) ? keys[lo + i] : ~0ull;
      gv[i] = (i < n) ? vals[lo + i] : 0u;
    }
    __syncthreads();
    bitonic_sort_block(gk, gv, npow2);
    for (unsigned i = threadIdx.x; i < n; i += kUniqueThreads) {
      const unsigned long long k = gk[i];
      keys[lo + i] = k;
      vals[lo + i] = gv[i];
      nhead += (i == 0 || k != gk[i - 1]);
    }
  }
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) nhead += __shfl_xor_sync(0xffffffffu, nhead, of);
  if (lane_id() == 0) s_cnt[threadIdx.x >> 5] = nhead;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tot = 0;
#pragma unroll
    for (int w = 0; w < kUniqueThreads / 32; ++w) tot += s_cnt[w];
    group_heads[g] = tot;
    if (tot) atomicAdd(gblock_heads + (g >> 8), tot);
  }
  trace_end(tr);
}

// ---- kernel B: number the runs (vertex ids), emit everything that hangs off a vertex -----------------------
__global__ void __launch_bounds__(256)
vertex_emit_kernel(const FwdBlock* __restrict__ blk, const unsigned long long* __restrict__ keys,
                   const unsigned* __restrict__ vals, DevCounters* __restrict__ ctr,
                   const unsigned* __restrict__ group_start, const unsigned* __restrict__ group_heads,
                   const unsigned* __restrict__ gblock_heads, int key_bits, float4* __restrict__ w_vert,
                   float4* __restrict__ w_acc, int32_t* __restrict__ owner) {
  constexpr int WARPS = 256 / 32;
  __shared__ unsigned s_w[WARPS];
  __shared__ unsigned s_part[WARPS];

  const unsigned t1 = ctr->work_tri;
  const int64_t ncorn = 3ll * t1 + 4ll * ctr->work_quad;
  const unsigned ngroups_used = (unsigned)((ncorn + kSortGroup - 1) / kSortGroup);
  const unsigned g = blockIdx.x;
  unsigned long long* tr = trace_begin(ctr->trace, ctr->trace_frame, K_VERTEX_EMIT);
  if (g >= ngroups_used) return;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  const unsigned lo = __ldcg(group_start + g);
  const unsigned hi = (g + 1 == ngroups_used) ? (unsigned)ncorn : __ldcg(group_start + g + 1);
  const unsigned n = hi > lo ? hi - lo : 0u;
  const bool is_last = g + 1 == ngroups_used;
  if (n == 0u && !is_last) return;

  // vertices of all earlier groups: the earlier groups of this 256-block + the earlier blocks (plain loads: the sort
  // kernel has completed)
  const unsigned blk256 = g >> 8, first = blk256 << 8;
  unsigned part = 0;
  if (first + threadIdx.x < g) part = __ldcg(group_heads + first + threadIdx.x);
  for (unsigned bq = threadIdx.x; bq < blk256; bq += 256) part += __ldcg(gblock_heads + bq);

  // heads of this thread's keys: blocked, ipt consecutive keys per thread
  const unsigned ipt = (n + 255) / 256;
  const unsigned i0 = threadIdx.x * ipt;
  const unsigned long long* __restrict__ k = keys + lo;
  unsigned long long prev = (i0 > 0 && i0 - 1 < n) ? k[i0 - 1] : 0ull;
  unsigned nhead = 0;
  for (unsigned j = 0; j < ipt; ++j) {
    const unsigned i = i0 + j;
    if (i < n) {
      const unsigned long long key = k[i];
      nhead += (i == 0 || key != prev);
      prev = key;
    }
  }
  unsigned incl = nhead;
#pragma unroll
  for (int of = 1; of < 32; of <<= 1) {
    const unsigned nb = __shfl_up_sync(0xffffffffu, incl, of);
    if (lane >= (unsigned)of) incl += nb;
  }
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) part += __shfl_xor_sync(0xffffffffu, part, of);
  if (lane == 31) s_w[warp] = incl;
  if (lane == 0) s_part[warp] = part;
  __syncthreads();
  unsigned wpre = 0, total = 0, excl = 0;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) {
    if (w < (int)warp) wpre += s_w[w];
    total += s_w[w];
    excl += s_part[w];
  }
  const d3h_forward_args& a = blk->a;
  const int64_t cap_verts = a.cap_verts, cap_verts_aug = a.cap_verts_aug;
  if (is_last && threadIdx.x == 0) {
    const unsigned nv = excl + total;
    ctr->n_verts = nv;
    if ((int64_t)nv <= cap_verts) a.tape_runs[nv] = (int32_t)ncorn;
  }
  if (n == 0u) return;

  const float* __restrict__ pos = a.pos;
  const float* __restrict__ sdf = a.sdf;
  const float* __restrict__ msdf = a.msdf;
  const int msdf_negate = a.msdf_negate;
  int32_t* __restrict__ tape_corners = a.tape_corners;
  int32_t* __restrict__ tape_slots = a.tape_slots;
  int64_t vid = (int64_t)excl + wpre + (incl - nhead) - 1;  // id of the run that precedes this thread's keys
  const unsigned long long bmask = (1ull << key_bits) - 1;
  prev = (i0 > 0 && i0 - 1 < n) ? k[i0 - 1] : 0ull;
  for (unsigned j = 0; j < ipt; ++j) {
    const unsigned i = i0 + j;
    if (i >= n) break;
    const unsigned long long key = k[i];
    // value = (class, 4*class_rank + corner) -> corner slot in the [3*T1 | 4*T2] layout
    const unsigned val = vals[lo + i];
    const unsigned r4 = val & 0x7fffffffu;
    const int64_t slot = (val >> 31) ? (3ll * t1 + r4) : (3ll * (r4 >> 2) + (r4 & 3u));
    if (i == 0 || key != prev) {
      ++vid;
      const int ea = (int)(key >> key_bits), eb = (int)(key & bmask);
      // zero-crossing interpolation, gshell_tets.py:291-303 (op order: SURVEY A.4)
      float w0, w1, dd;
      crossing_weights(__ldg(sdf + ea), __ldg(sdf + eb), w0, w1, dd);
      float ma = __ldg(msdf + ea), mb = __ldg(msdf + eb);
      if (msdf_negate) { ma = -ma; mb = -mb; }
      const float x = lerp2(__ldg(pos + 3ll * ea + 0), w0, __ldg(pos + 3ll * eb + 0), w1);
      const float y = lerp2(__ldg(pos + 3ll * ea + 1), w0, __ldg(pos + 3ll * eb + 1), w1);
      const float z = lerp2(__ldg(pos + 3ll * ea + 2), w0, __ldg(pos + 3ll * eb + 2), w1);
      const float m = lerp2(ma, w0, mb, w1);
      w_vert[vid] = make_float4(x, y, z, m);
      w_acc[2 * vid] = make_float4(0.f, 0.f, 0.f, 0.f);
      w_acc[2 * vid + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
      owner[vid] = (int32_t)slot;
      if (vid < cap_verts) {
        a.tape_edges[2 * vid] = ea;
        a.tape_edges[2 * vid + 1] = eb;
        a.tape_runs[vid] = (int32_t)(lo + i);
        a.verts_wt[3 * vid] = x; a.verts_wt[3 * vid + 1] = y; a.verts_wt[3 * vid + 2] = z;
        a.msdf_wt[vid] = m;
      }
      if (vid < cap_verts_aug) {
        // rows of verts_aug not referenced by faces_aug are zero (gshell_tets.py:423-427); a watertight vertex is
        // referenced iff its mSDF is positive (every cut case keeps exactly the positive corners)
        const bool used = m > 0.f;
        a.verts_aug[3 * vid] = used ? x : 0.f;
        a.verts_aug[3 * vid + 1] = used ? y : 0.f;
        a.verts_aug[3 * vid + 2] = used ? z : 0.f;
        a.msdf_aug[vid] = m;
      }
    }
    prev = key;
    tape_corners[slot] = (int32_t)vid;
    tape_slots[lo + i] = (int32_t)slot;
  }
  trace_end(tr);
}

// ------------------------------------------------------------------------------------------------
// static edge table path: vertex numbering without a sort
// ------------------------------------------------------------------------------------------------
// The compaction kernel has marked the crossing edges of this call in a bitmap over the static, lexicographically
// sorted list of all tet edges.  The rank of a marked edge among the marked edges IS its vertex id (the order of
// torch.unique(dim=0), gshell_tets.py:279-287).  One CTA per 8192 edges: CTAs of blocks without a mark (most) exit on
// the block counter, the others sum the counters of the earlier blocks themselves (as compact_kernel does), scan the
// popcounts of their 256 words and emit one vertex per set bit.
__global__ void __launch_bounds__(256)
edge_emit_kernel(const FwdBlock* __restrict__ blk, const unsigned* __restrict__ edge_bits,
                 const unsigned* __restrict__ eblock_cnt, unsigned* __restrict__ word_prefix, int64_t n_eblocks,
                 DevCounters* __restrict__ ctr, float4* __restrict__ w_vert, float4* __restrict__ w_acc) {
  constexpr int WARPS = 256 / 32;
  __shared__ unsigned s_w[WARPS];
  __shared__ unsigned s_part[WARPS];
  __shared__ unsigned s_pre[256];   // marked edges of this block before word t
  __shared__ unsigned s_bits[256];
  unsigned long long* tr = trace_begin(ctr->trace, ctr->trace_frame, K_EDGE_EMIT);
  const unsigned b = blockIdx.x;
  const unsigned mine = __ldcg(eblock_cnt + b);
  const bool is_last = (int64_t)b == n_eblocks - 1;
  // an overflowed record buffer skips the surface stages (work_* = 0): leave n_verts = 0 like the general path
  const bool skip = (ctr->work_tri + ctr->work_quad) == 0u;
  if ((mine == 0u && !is_last) || skip) return;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  unsigned part = 0;
#pragma unroll 4
  for (unsigned i = threadIdx.x; i < b; i += 256) part += __ldcg(eblock_cnt + i);
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) part += __shfl_xor_sync(0xffffffffu, part, of);
  const int64_t w = (int64_t)b * 256 + threadIdx.x;
  const unsigned bits = __ldcg(edge_bits + w);
  const unsigned c = __popc(bits);
  unsigned incl = c;
#pragma unroll
  for (int of = 1; of < 32; of <<= 1) {
    const unsigned nb = __shfl_up_sync(0xffffffffu, incl, of);
    if (lane >= (unsigned)of) incl += nb;
  }
  if (lane == 31) s_w[warp] = incl;
  if (lane == 0) s_part[warp] = part;
  s_bits[threadIdx.x] = bits;
  __syncthreads();
  unsigned wpre = 0, total = 0, excl = 0;
#pragma unroll
  for (int q = 0; q < WARPS; ++q) {
    if (q < (int)warp) wpre += s_w[q];
    total += s_w[q];
    excl += s_part[q];
  }
  s_pre[threadIdx.x] = wpre + incl - c;
  if (c) word_prefix[w] = excl + wpre + incl - c;
  const d3h_forward_args& a = blk->a;
  const int64_t cap_verts = a.cap_verts, cap_verts_aug = a.cap_verts_aug;
  if (is_last && threadIdx.x == 0) ctr->n_verts = excl + total;
  if (total == 0u) return;
  __syncthreads();
  const int2* __restrict__ edge_ab = reinterpret_cast<const int2*>(a.edge_ab);
  const float* __restrict__ pos = a.pos;
  const float* __restrict__ sdf = a.sdf;
  const float* __restrict__ msdf = a.msdf;
  const int msdf_negate = a.msdf_negate;
  float4* __restrict__ vacc = reinterpret_cast<float4*>(a.vacc);
  // one marked edge (= one new vertex) per thread and trip, whatever word it sits in
  for (unsigned i = threadIdx.x; i < total; i += 256) {
    int lo = 0, hi = 255;  // word: last t with s_pre[t] <= i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_pre[mid] <= i) lo = mid; else hi = mid - 1;
    }
    const int bit = (int)__fns(s_bits[lo], 0, (int)(i - s_pre[lo]) + 1);
    const int64_t vid = (int64_t)excl + i;
    const int2 ab = __ldg(edge_ab + (((int64_t)b * 256 + lo) * 32 + bit));
    const int ea = ab.x, eb = ab.y;
    // zero-crossing interpolation, gshell_tets.py:291-303 (op order: SURVEY A.4)
    float w0, w1, dd;
    crossing_weights(__ldg(sdf + ea), __ldg(sdf + eb), w0, w1, dd);
    float ma = __ldg(msdf + ea), mb = __ldg(msdf + eb);
    if (msdf_negate) { ma = -ma; mb = -mb; }
    const float x = lerp2(__ldg(pos + 3ll * ea + 0), w0, __ldg(pos + 3ll * eb + 0), w1);
    const float y = lerp2(__ldg(pos + 3ll * ea + 1), w0, __ldg(pos + 3ll * eb + 1), w1);
    const float z = lerp2(__ldg(pos + 3ll * ea + 2), w0, __ldg(pos + 3ll * eb + 2), w1);
    const float m = lerp2(ma, w0, mb, w1);
    w_vert[vid] = make_float4(x, y, z, m);
    w_acc[2 * vid] = make_float4(0.f, 0.f, 0.f, 0.f);
    w_acc[2 * vid + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vid < cap_verts) {
      reinterpret_cast<int2*>(a.tape_edges)[vid] = make_int2(ea, eb);
      a.verts_wt[3 * vid] = x; a.verts_wt[3 * vid + 1] = y; a.verts_wt[3 * vid + 2] = z;
      a.msdf_wt[vid] = m;
      vacc[2 * vid] = make_float4(0.f, 0.f, 0.f, 0.f);
      vacc[2 * vid + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (vid < cap_verts_aug) {
      // rows of verts_aug not referenced by faces_aug are zero (gshell_tets.py:423-427); a watertight vertex is
      // referenced iff its mSDF is positive (every cut case keeps exactly the positive corners)
      const bool used = m > 0.f;
      a.verts_aug[3 * vid] = used ? x : 0.f;
      a.verts_aug[3 * vid + 1] = used ? y : 0.f;
      a.verts_aug[3 * vid + 2] = used ? z : 0.f;
      a.msdf_aug[vid] = m;
    }
  }
  trace_end(tr);
}

void launch_edge_emit(const d3h_forward_args& a, const Workspace& ws, cudaStream_t stream) {
  if (ws.cap_corners <= 0 || ws.n_eblocks <= 0) return;
  ProfScope ps(K_EDGE_EMIT, stream);
  launch_k(edge_emit_kernel, (unsigned)ws.n_eblocks, 256u, stream, kLaunchLatency, ws.blk, ws.edge_bits, ws.eblock_cnt,
           ws.word_prefix, ws.n_eblocks, ws.ctr, ws.vert, reinterpret_cast<float4*>(ws.acc));
}

// ------------------------------------------------------------------------------------------------
void launch_edge_sort(const d3h_forward_args& a, const Workspace& ws, cudaStream_t stream) {
  const int key_bits = key_bits_for(a.n_grid);
  const int64_t capc = ws.cap_corners;
  if (capc <= 0) return;
  {
    ProfScope ps(K_BUCKET_SCAN, stream);
    launch_k(bucket_scan_kernel, (unsigned)ws.nscan_ctas, (unsigned)kScanThreads, stream, kLaunchLatency, ws.msd_hist,
             ws.msd_base, ws.msd_fill, ws.group_start, (int)ws.msd_bins, ws.ctr, ws.st_scan);
  }
  {
    ProfScope ps(K_PARTITION, stream);
    launch_k(partition_kernel, (unsigned)((capc + 255) / 256), 256u, stream, kLaunchLatency, ws.keys, ws.vals, ws.keys2,
             ws.vals2, ws.ctr, ws.msd_fill, key_bits + msd_shift_for(a.n_grid));
  }
  {
    ProfScope ps(K_GROUP_SORT, stream);
    launch_k(group_sort_kernel, (unsigned)ws.ngroups, (unsigned)kUniqueThreads, stream, kLaunchLatency, ws.keys2,
             ws.vals2, ws.keys_scratch, ws.vals_scratch, ws.ctr, ws.group_start, ws.group_heads, ws.gblock_heads);
  }
  ProfScope ps(K_VERTEX_EMIT, stream);
  launch_k(vertex_emit_kernel, (unsigned)ws.ngroups, 256u, stream, kLaunchLatency, ws.blk, ws.keys2, ws.vals2, ws.ctr,
           ws.group_start, ws.group_heads, ws.gblock_heads, key_bits, ws.vert, reinterpret_cast<float4*>(ws.acc),
           ws.owner);
}

}  // namespace d3h

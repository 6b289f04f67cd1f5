// Stage 1 of the extraction: occupancy bitmap, the only O(F) kernel (streaming classification) and the ordered
// compaction of the valid tets.
//
// Replaces gshell_tets.py:260-275 (occ_n, occ_fx4, occ_sum, valid_tets), :307-309 (tetindex, num_triangles)
// and the boolean-mask compactions `tet_fx4[valid_tets]`, `idx_map[num_triangles == k]` (:277, :323-324).
//
//   prepare_kernel   N-sized: sign bitmap of sdf (N/8 bytes: 268 KB at 128^3, L1/L2 resident) + reset of scan state.
//   classify_kernel  F-sized, pure stream: every warp loads 8 x 32 tets with 16-byte no-allocate loads, looks the four
//                    signs up in the bitmap and writes two ballot words per 32 tets (tet yields 1 / 2 triangles).
//                    No shared memory, no barrier, no atomics: the kernel is bound by the 16 B/tet HBM stream.
//   compact_kernel   scans the two class bitmaps (F/4 bytes, L2 resident) with a decoupled look-back, then visits only
//                    the valid tets (~1 %) in tet order: re-reads their indices, writes the compact records and, in the
//                    fused single-GPU path, the sort keys of their crossing edges + the MSD histogram.
//
// v1 of this file classified and compacted in one kernel (ticket + block scan + look-back per 2048-tet tile): ncu showed
// 45 % of the warp samples parked on the barrier behind the ticket atomic and 16 % DRAM utilisation
// (profiles/r01a_launches_v1.csv); splitting the ordered part off removes every dependency from the stream.
#include "d3h_internal.cuh"

namespace d3h {

// ------------------------------------------------------------------------------------------------
// K0: occupancy bitmaps + reset of all per-call scan state
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 load4_guarded(const float* __restrict__ p, int64_t q, int64_t n) {
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (4 * q + 3 < n) {
    s = __ldg(reinterpret_cast<const float4*>(p) + q);
  } else if (4 * q < n) {
    s.x = p[4 * q];
    if (4 * q + 1 < n) s.y = p[4 * q + 1];
    if (4 * q + 2 < n) s.z = p[4 * q + 2];
  }
  return s;
}

__global__ void __launch_bounds__(256) prepare_kernel(const float* __restrict__ sdf, const float* __restrict__ msdf,
                                                      int64_t n_grid, int msdf_negate, int want_mocc,
                                                      unsigned* __restrict__ occ_bits,
                                                      unsigned* __restrict__ mocc_bits, Workspace ws) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  if (tid < (int64_t)(sizeof(DevCounters) / 4)) reinterpret_cast<unsigned*>(ws.ctr)[tid] = 0u;
  for (int64_t i = tid; i < ws.ntiles_compact; i += nthreads) ws.st_compact[i] = 0ull;
  for (int64_t i = tid; i < ws.ntiles_rle; i += nthreads) ws.st_rle[i] = 0ull;
  for (int64_t i = tid; i < ws.ntiles_poly * 3; i += nthreads) ws.st_poly[i] = 0ull;
  for (int64_t i = tid; i < ws.msd_bins + 8; i += nthreads) {
    ws.msd_hist[i] = 0u;
    ws.msd_fill[i] = 0u;
  }

  // bitmaps: each lane takes 4 consecutive vertices (one 16-byte load), 8 lanes make one 32-bit word
  const int64_t nquads = (n_grid + 3) / 4;
  const unsigned lane = lane_id();
  for (int64_t q0 = tid - lane; q0 < nquads; q0 += nthreads) {  // warp-uniform trip count
    const int64_t q = q0 + lane;
    const float4 s = load4_guarded(sdf, q, n_grid);
    unsigned word = ((s.x > 0.f) | ((s.y > 0.f) << 1) | ((s.z > 0.f) << 2) | ((s.w > 0.f) << 3)) << (4 * (lane & 7));
    word |= __shfl_xor_sync(0xffffffffu, word, 1);
    word |= __shfl_xor_sync(0xffffffffu, word, 2);
    word |= __shfl_xor_sync(0xffffffffu, word, 4);
    if ((lane & 7) == 0 && 4 * q < n_grid) occ_bits[q >> 3] = word;
    if (want_mocc) {
      float4 m = load4_guarded(msdf, q, n_grid);
      if (msdf_negate) { m.x = -m.x; m.y = -m.y; m.z = -m.z; m.w = -m.w; }
      unsigned mw = ((m.x > 0.f) | ((m.y > 0.f) << 1) | ((m.z > 0.f) << 2) | ((m.w > 0.f) << 3)) << (4 * (lane & 7));
      mw |= __shfl_xor_sync(0xffffffffu, mw, 1);
      mw |= __shfl_xor_sync(0xffffffffu, mw, 2);
      mw |= __shfl_xor_sync(0xffffffffu, mw, 4);
      if ((lane & 7) == 0 && 4 * q < n_grid) mocc_bits[q >> 3] = mw;
    }
  }
}

void launch_prepare(const d3h_forward_args& a, const Workspace& ws, cudaStream_t stream) {
  const int64_t nquads = (a.n_grid + 3) / 4;
  int64_t blocks = (nquads + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ProfScope ps(K_PREPARE, stream);
  prepare_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a.sdf, a.msdf, a.n_grid, a.msdf_negate,
                                                         a.watertight_template ? 0 : 1, ws.occ_bits, ws.mocc_bits, ws);
}

// ------------------------------------------------------------------------------------------------
// K1: streaming classification
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned occ_of(const unsigned* __restrict__ bits, int v) {
  return (__ldg(bits + (v >> 5)) >> (v & 31)) & 1u;
}

__global__ void __launch_bounds__(kClassifyThreads)
classify_kernel(const int4* __restrict__ tets, int64_t tet_begin, int64_t tet_end,
                const unsigned* __restrict__ occ_bits, const unsigned* __restrict__ mocc_bits,
                unsigned* __restrict__ m1_words, unsigned* __restrict__ m2_words) {
  const unsigned lane = lane_id();
  const int64_t warp = ((int64_t)blockIdx.x * kClassifyThreads + threadIdx.x) >> 5;
  const int64_t base = tet_begin + warp * (32 * kClassifyItems);
  if (base >= tet_end) return;

  int4 t[kClassifyItems];
#pragma unroll
  for (int j = 0; j < kClassifyItems; ++j) {
    const int64_t idx = base + j * 32 + lane;
    t[j] = (idx < tet_end) ? ld_stream_int4(tets + idx) : make_int4(0, 0, 0, 0);
  }
  unsigned w1 = 0, w2 = 0;  // lane j keeps the ballot words of item j
#pragma unroll
  for (int j = 0; j < kClassifyItems; ++j) {
    const int64_t idx = base + j * 32 + lane;
    const unsigned c = occ_of(occ_bits, t[j].x) + occ_of(occ_bits, t[j].y) + occ_of(occ_bits, t[j].z) +
                       occ_of(occ_bits, t[j].w);
    bool valid = (c != 0u) && (c != 4u) && (idx < tet_end);
    if (mocc_bits != nullptr && valid) {  // open-mesh prefilter, gshell_tets.py:275
      valid = (occ_of(mocc_bits, t[j].x) | occ_of(mocc_bits, t[j].y) | occ_of(mocc_bits, t[j].z) |
               occ_of(mocc_bits, t[j].w)) != 0u;
    }
    const unsigned b1 = __ballot_sync(0xffffffffu, valid && (c != 2u));  // 1 or 3 inside -> one triangle
    const unsigned b2 = __ballot_sync(0xffffffffu, valid && (c == 2u));  // 2 inside      -> two triangles
    if (lane == (unsigned)j) { w1 = b1; w2 = b2; }
  }
  if (lane < (unsigned)kClassifyItems) {
    const int64_t w = ((base - tet_begin) >> 5) + lane;
    m1_words[w] = w1;
    m2_words[w] = w2;
  }
}

// ------------------------------------------------------------------------------------------------
// K1b: ordered compaction of the valid tets (+ fused key emission)
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long kFlagAgg = 1ull << 62;  // tile aggregate published
constexpr unsigned long long kFlagInc = 2ull << 62;  // inclusive prefix published
constexpr unsigned long long kValMask = (1ull << 62) - 1;
// value = count(T1 class) in bits [0,31) | count(T2 class) in bits [31,62)

// Writes the sort keys of one valid tet: one key per polygon corner (= crossing edge, in mesh_edge_table order).
// Slots are dense in valid-tet order (3 per tri tet, 4 per quad tet); the value encodes the final corner slot
// [3*T1 | 4*T2] as (class, 4*class_rank + k) because T1 is not known yet.
__device__ __forceinline__ void emit_polygon_keys(const int4 v4, int code, bool quad, unsigned class_rank,
                                                  unsigned other_before, int key_bits, int msd_shift,
                                                  unsigned long long* __restrict__ keys, unsigned* __restrict__ vals,
                                                  unsigned* __restrict__ msd_hist) {
  const int vv[4] = {v4.x, v4.y, v4.z, v4.w};
  const int n = quad ? 4 : 3;
  const int64_t slot0 = quad ? (4ll * class_rank + 3ll * other_before) : (3ll * class_rank + 4ll * other_before);
  for (int k = 0; k < n; ++k) {
    const int e = c_loop_edge[code][k];
    const int p = vv[c_edge_p[e]], q = vv[c_edge_q[e]];
    const unsigned long long a = (unsigned)min(p, q), b = (unsigned)max(p, q);
    keys[slot0 + k] = (a << key_bits) | b;
    vals[slot0 + k] = (quad ? 0x80000000u : 0u) | (4u * class_rank + (unsigned)k);
    atomicAdd(&msd_hist[(unsigned)(a >> msd_shift)], 1u);  // global RED; buckets are fine-grained, contention is low
  }
}

// Last-arriving CTA: exclusive scan of the MSD histogram -> bucket bases, and the first key of every sort group.
// Group g owns key positions [snap(g*G), snap((g+1)*G)), snap(x) = base of the bucket that contains position x: groups
// are unions of whole buckets, tile [0,P) exactly and hold fewer than G + (largest bucket) keys.
__device__ void msd_scan_epilogue(const unsigned* __restrict__ hist, unsigned* __restrict__ base,
                                  unsigned* __restrict__ group_start, int nbins, unsigned* s_tmp /* >= 32 words */) {
  const int per = ((nbins + (int)blockDim.x - 1) / (int)blockDim.x + 3) & ~3;  // bins per thread, multiple of 4
  const int b0 = threadIdx.x * per;
  unsigned sum = 0;
  for (int i = 0; i < per; i += 4) {
    if (b0 + i < nbins) {  // hist is padded by 8 zeroed words: the vector load may run past nbins
      const uint4 h = __ldcg(reinterpret_cast<const uint4*>(hist + b0 + i));
      sum += h.x + h.y + h.z + h.w;
    }
  }
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  unsigned incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (unsigned)o) incl += n;
  }
  __syncthreads();  // s_tmp may alias a buffer the caller used before
  if (lane == 31) s_tmp[warp] = incl;
  __syncthreads();
  unsigned wpre = 0;
  for (unsigned w = 0; w < warp; ++w) wpre += s_tmp[w];
  unsigned run = wpre + incl - sum;
  for (int i = 0; i < per; ++i) {
    const int b = b0 + i;
    if (b < nbins) {
      const unsigned c = __ldcg(hist + b);
      base[b] = run;
      if (c) {
        for (unsigned g = (run + kSortGroup - 1) / kSortGroup; (uint64_t)g * kSortGroup < (uint64_t)run + c; ++g)
          group_start[g] = run;
      }
      run += c;
    }
  }
  if (threadIdx.x == blockDim.x - 1) {
    base[nbins] = run;  // = P
    group_start[(run + kSortGroup - 1) / kSortGroup] = run;
  }
}

template <bool EMIT_KEYS>
__global__ void __launch_bounds__(kCompactThreads)
compact_kernel(const unsigned* __restrict__ m1_words, const unsigned* __restrict__ m2_words, int64_t nwords,
               const int4* __restrict__ tets, int64_t tet_begin, const unsigned* __restrict__ occ_bits,
               unsigned long long* __restrict__ status, DevCounters* __restrict__ ctr,
               d3h_tet_record* __restrict__ records, int64_t cap_records, int64_t ntiles, int key_bits, int msd_shift,
               unsigned long long* __restrict__ keys, unsigned* __restrict__ vals, unsigned* __restrict__ msd_hist,
               unsigned* __restrict__ msd_base, unsigned* __restrict__ group_start, int msd_bins) {
  constexpr int WARPS = kCompactThreads / 32;
  __shared__ unsigned s_pre1[kCompactThreads], s_pre2[kCompactThreads];  // exclusive per-thread prefixes in the tile
  __shared__ unsigned s_m1[kCompactThreads * kCompactWords], s_m2[kCompactThreads * kCompactWords];
  __shared__ unsigned s_w1[32], s_w2[32];
  __shared__ unsigned long long s_excl;
  __shared__ unsigned s_total[2];
  __shared__ unsigned s_tile, s_last;

  if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->ticket_compact, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;

  // each thread owns kCompactWords consecutive words of both bitmaps (= 32*kCompactWords consecutive tets)
  const int64_t w0 = ((int64_t)tile * kCompactThreads + threadIdx.x) * kCompactWords;
  unsigned c1 = 0, c2 = 0;
#pragma unroll
  for (int j = 0; j < kCompactWords; ++j) {
    const unsigned a1 = (w0 + j < nwords) ? __ldcg(m1_words + w0 + j) : 0u;
    const unsigned a2 = (w0 + j < nwords) ? __ldcg(m2_words + w0 + j) : 0u;
    s_m1[threadIdx.x * kCompactWords + j] = a1;
    s_m2[threadIdx.x * kCompactWords + j] = a2;
    c1 += __popc(a1);
    c2 += __popc(a2);
  }
  // block exclusive scan of (c1, c2)
  unsigned i1 = c1, i2 = c2;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned n1 = __shfl_up_sync(0xffffffffu, i1, o), n2 = __shfl_up_sync(0xffffffffu, i2, o);
    if (lane >= (unsigned)o) { i1 += n1; i2 += n2; }
  }
  if (lane == 31) { s_w1[warp] = i1; s_w2[warp] = i2; }
  __syncthreads();
  unsigned p1 = 0, p2 = 0, tot1 = 0, tot2 = 0;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) {
    if (w < (int)warp) { p1 += s_w1[w]; p2 += s_w2[w]; }
    tot1 += s_w1[w];
    tot2 += s_w2[w];
  }
  s_pre1[threadIdx.x] = p1 + i1 - c1;
  s_pre2[threadIdx.x] = p2 + i2 - c2;

  if (warp == 0) {
    const unsigned long long agg = (unsigned long long)tot1 | ((unsigned long long)tot2 << 31);
    unsigned long long excl_tiles = 0ull;
    if (tile == 0) {
      if (lane == 0) st_relaxed_u64(status, kFlagInc | agg);
    } else {
      if (lane == 0) st_relaxed_u64(status + tile, kFlagAgg | agg);
      int64_t look = (int64_t)tile - 1;
      while (true) {
        const int64_t idx = look - lane;
        unsigned long long w = kFlagInc;  // virtual tile -1: inclusive prefix 0
        if (idx >= 0) {
          do { w = ld_relaxed_u64(status + idx); } while ((w >> 62) == 0ull);
        }
        const unsigned inc_mask = __ballot_sync(0xffffffffu, (w >> 62) == 2ull);
        const int first = inc_mask ? (__ffs(inc_mask) - 1) : 32;
        unsigned long long contrib = ((int)lane <= first) ? (w & kValMask) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl_tiles += contrib;
        if (inc_mask) break;
        look -= 32;
      }
      if (lane == 0) st_relaxed_u64(status + tile, kFlagInc | (excl_tiles + agg));
    }
    if (lane == 0) {
      s_excl = excl_tiles;
      s_total[0] = tot1;
      s_total[1] = tot2;
      if ((int64_t)tile == ntiles - 1) {  // grid totals
        const unsigned long long incl_all = excl_tiles + agg;
        const unsigned t1 = (unsigned)(incl_all & 0x7fffffffull), t2 = (unsigned)(incl_all >> 31);
        ctr->n_tri = t1;
        ctr->n_quad = t2;
        ctr->n_valid = t1 + t2;
        const bool fits = (int64_t)t1 + t2 <= cap_records;
        ctr->work_tri = fits ? t1 : 0u;
        ctr->work_quad = fits ? t2 : 0u;
      }
    }
  }
  __syncthreads();

  // ---- visit the tile's valid tets, one per thread per round, in tet order ----
  const unsigned e1 = (unsigned)(s_excl & 0x7fffffffull), e2 = (unsigned)(s_excl >> 31);
  const unsigned nvalid_tile = s_total[0] + s_total[1];
  for (unsigned i = threadIdx.x; i < nvalid_tile; i += kCompactThreads) {
    int lo = 0, hi = kCompactThreads - 1;  // owner thread: last th with pre1[th] + pre2[th] <= i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_pre1[mid] + s_pre2[mid] <= i) lo = mid; else hi = mid - 1;
    }
    const int th = lo;
    unsigned r1 = s_pre1[th], r2 = s_pre2[th];
    unsigned rem = i - (r1 + r2);
    const int64_t wb = ((int64_t)tile * kCompactThreads + th) * kCompactWords;
    unsigned b1 = 0, b2 = 0;
    int64_t word = wb;
#pragma unroll
    for (int j = 0; j < kCompactWords; ++j) {
      b1 = s_m1[th * kCompactWords + j];
      b2 = s_m2[th * kCompactWords + j];
      word = wb + j;
      const unsigned cnt = __popc(b1 | b2);
      if (rem < cnt) break;
      rem -= cnt;
      r1 += __popc(b1);
      r2 += __popc(b2);
    }
    const unsigned both = b1 | b2;
    const int bit = (int)__fns(both, 0, (int)rem + 1);  // position of the (rem+1)-th set bit
    const unsigned below = (1u << bit) - 1u;
    r1 += __popc(b1 & below);
    r2 += __popc(b2 & below);
    const bool quad = (b2 >> bit) & 1u;
    const unsigned g1 = e1 + r1, g2 = e2 + r2;  // tri / quad valid tets before this one, grid-wide
    const int64_t slot = (int64_t)g1 + g2;
    const int64_t tet = tet_begin + word * 32 + bit;
    if (slot < cap_records) {
      const int4 v4 = __ldg(tets + tet);
      const int code = (int)(occ_of(occ_bits, v4.x) | (occ_of(occ_bits, v4.y) << 1) | (occ_of(occ_bits, v4.z) << 2) |
                             (occ_of(occ_bits, v4.w) << 3));
      int4* out = reinterpret_cast<int4*>(records + slot);
      out[0] = v4;
      out[1] = make_int4(code, (int)(quad ? g2 : g1), (int)tet, (int)(quad ? g1 : g2));
      if (EMIT_KEYS)
        emit_polygon_keys(v4, code, quad, quad ? g2 : g1, quad ? g1 : g2, key_bits, msd_shift, keys, vals, msd_hist);
    }
  }
  if (EMIT_KEYS) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&ctr->compact_done, 1u) == (unsigned)(ntiles - 1));
    __syncthreads();
    if (s_last) {
      __threadfence();
      msd_scan_epilogue(msd_hist, msd_base, group_start, msd_bins, s_w1);
    }
  }
}

void launch_classify(const d3h_forward_args& a, const Workspace& ws, d3h_tet_record* records, int64_t cap_records,
                     bool emit_keys, cudaStream_t stream) {
  const int64_t n = a.tet_end - a.tet_begin;
  if (n <= 0) return;
  const int64_t nwarps = (n + 32 * kClassifyItems - 1) / (32 * kClassifyItems);
  const int64_t nblocks = (nwarps * 32 + kClassifyThreads - 1) / kClassifyThreads;
  {
    ProfScope ps(K_CLASSIFY, stream);
    classify_kernel<<<(unsigned)nblocks, kClassifyThreads, 0, stream>>>(
        reinterpret_cast<const int4*>(a.tets), a.tet_begin, a.tet_end, ws.occ_bits,
        a.watertight_template ? nullptr : ws.mocc_bits, ws.m1_words, ws.m2_words);
  }
  const int64_t nwords = nwarps * kClassifyItems;  // every word of a launched warp is written
  const int64_t ntiles = (nwords + kCompactThreads * kCompactWords - 1) / (kCompactThreads * kCompactWords);
  const int key_bits = key_bits_for(a.n_grid);
  const int msd_shift = msd_shift_for(a.n_grid);
  ProfScope ps(K_COMPACT, stream);
  if (emit_keys)
    compact_kernel<true><<<(unsigned)ntiles, kCompactThreads, 0, stream>>>(
        ws.m1_words, ws.m2_words, nwords, reinterpret_cast<const int4*>(a.tets), a.tet_begin, ws.occ_bits, ws.st_compact,
        ws.ctr, records, cap_records, ntiles, key_bits, msd_shift, ws.keys, ws.vals, ws.msd_hist, ws.msd_base,
        ws.group_start, (int)ws.msd_bins);
  else
    compact_kernel<false><<<(unsigned)ntiles, kCompactThreads, 0, stream>>>(
        ws.m1_words, ws.m2_words, nwords, reinterpret_cast<const int4*>(a.tets), a.tet_begin, ws.occ_bits, ws.st_compact,
        ws.ctr, records, cap_records, ntiles, key_bits, msd_shift, nullptr, nullptr, nullptr, nullptr, nullptr, 0);
}

// ------------------------------------------------------------------------------------------------
// records gathered from several shards: recompute the ranks over the concatenation, then emit keys
// (single block for the ranks; the merged list is O(surface))
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) rank_records_kernel(d3h_tet_record* __restrict__ records, int64_t n,
                                                            DevCounters* __restrict__ ctr) {
  __shared__ unsigned s_w1[32], s_w2[32];
  __shared__ unsigned s_run1, s_run2;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_run1 = 0; s_run2 = 0; }
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    int cls = 0;
    if (i < n) cls = (__popc((unsigned)records[i].code) == 2) ? 2 : 1;
    const unsigned b1 = __ballot_sync(0xffffffffu, cls == 1), b2 = __ballot_sync(0xffffffffu, cls == 2);
    if (lane == 0) { s_w1[warp] = __popc(b1); s_w2[warp] = __popc(b2); }
    __syncthreads();
    unsigned p1 = 0, p2 = 0, tot1 = 0, tot2 = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < (int)warp) { p1 += s_w1[w]; p2 += s_w2[w]; }
      tot1 += s_w1[w]; tot2 += s_w2[w];
    }
    const unsigned lt = lanemask_lt();
    const unsigned g1 = s_run1 + p1 + __popc(b1 & lt), g2 = s_run2 + p2 + __popc(b2 & lt);
    if (cls == 1) { records[i].class_rank = (int)g1; records[i].other_before = (int)g2; }
    if (cls == 2) { records[i].class_rank = (int)g2; records[i].other_before = (int)g1; }
    __syncthreads();
    if (threadIdx.x == 0) { s_run1 += tot1; s_run2 += tot2; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ctr->n_tri = s_run1;
    ctr->n_quad = s_run2;
    ctr->n_valid = s_run1 + s_run2;
    ctr->work_tri = s_run1;
    ctr->work_quad = s_run2;
  }
}

__global__ void __launch_bounds__(256)
keys_from_records_kernel(const d3h_tet_record* __restrict__ records, DevCounters* __restrict__ ctr, int key_bits,
                         int msd_shift, unsigned long long* __restrict__ keys, unsigned* __restrict__ vals,
                         unsigned* __restrict__ msd_hist, unsigned* __restrict__ msd_base,
                         unsigned* __restrict__ group_start, int msd_bins) {
  __shared__ unsigned s_tmp[32];
  __shared__ unsigned s_last;
  const int64_t n = (int64_t)ctr->work_tri + ctr->work_quad;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int4 v4 = reinterpret_cast<const int4*>(records + i)[0];
    const int4 meta = reinterpret_cast<const int4*>(records + i)[1];
    emit_polygon_keys(v4, meta.x, __popc((unsigned)meta.x) == 2, (unsigned)meta.y, (unsigned)meta.w, key_bits, msd_shift,
                      keys, vals, msd_hist);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ctr->compact_done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    msd_scan_epilogue(msd_hist, msd_base, group_start, msd_bins, s_tmp);
  }
}

void launch_rank_records(const d3h_forward_args& a, const Workspace& ws, d3h_tet_record* records, int64_t n_records,
                         cudaStream_t stream) {
  {
    ProfScope ps(K_RANK_RECORDS, stream);
    rank_records_kernel<<<1, 1024, 0, stream>>>(records, n_records, ws.ctr);
  }
  int64_t blocks = (n_records + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 2) blocks = 148 * 2;
  ProfScope ps(K_COMPACT, stream);
  keys_from_records_kernel<<<(unsigned)blocks, 256, 0, stream>>>(records, ws.ctr, key_bits_for(a.n_grid),
                                                                 msd_shift_for(a.n_grid), ws.keys, ws.vals, ws.msd_hist,
                                                                 ws.msd_base, ws.group_start, (int)ws.msd_bins);
}

// ------------------------------------------------------------------------------------------------
// one-time packing / validation of the static tet index array
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_tets_i64_kernel(const longlong2* __restrict__ in, int64_t n_tets,
                                                            int64_t n_grid, int4* __restrict__ out,
                                                            unsigned long long* __restrict__ bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned nbad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_tets; i += stride) {
    const longlong2 lo = in[2 * i], hi = in[2 * i + 1];
    nbad += (lo.x < 0 || lo.x >= n_grid) + (lo.y < 0 || lo.y >= n_grid) + (hi.x < 0 || hi.x >= n_grid) +
            (hi.y < 0 || hi.y >= n_grid);
    out[i] = make_int4((int)lo.x, (int)lo.y, (int)hi.x, (int)hi.y);
  }
  if (nbad) atomicAdd(bad, (unsigned long long)nbad);
}

__global__ void __launch_bounds__(256) check_tets_i32_kernel(const int4* __restrict__ in, int64_t n_tets,
                                                             int64_t n_grid, unsigned long long* __restrict__ bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned nbad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_tets; i += stride) {
    const int4 t = in[i];
    nbad += (t.x < 0 || t.x >= n_grid) + (t.y < 0 || t.y >= n_grid) + (t.z < 0 || t.z >= n_grid) +
            (t.w < 0 || t.w >= n_grid);
  }
  if (nbad) atomicAdd(bad, (unsigned long long)nbad);
}

}  // namespace d3h

using namespace d3h;

extern "C" int d3h_pack_tets_i64(const int64_t* tets, int64_t n_tets, int64_t n_grid, int32_t* out_tets,
                                 int64_t* bad_count_dev, d3h_stream_t stream) {
  if (!tets || !out_tets || !bad_count_dev || n_tets < 0 || n_grid <= 0 || n_grid >= (1ll << 31) ||
      n_tets >= (1ll << 31) || (reinterpret_cast<uintptr_t>(out_tets) & 15) || (reinterpret_cast<uintptr_t>(tets) & 15)) {
    set_error("d3h_pack_tets_i64: bad argument (null/misaligned pointer, or N/F outside [0, 2^31))");
    return D3H_E_BADARG;
  }
  if (n_tets == 0) return D3H_OK;
  int64_t blocks = (n_tets + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_tets_i64_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const longlong2*>(tets), n_tets, n_grid, reinterpret_cast<int4*>(out_tets),
      reinterpret_cast<unsigned long long*>(bad_count_dev));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("d3h_pack_tets_i64: %s", cudaGetErrorString(e)); return D3H_E_CUDA; }
  return D3H_OK;
}

extern "C" int d3h_check_tets_i32(const int32_t* tets, int64_t n_tets, int64_t n_grid, int64_t* bad_count_dev,
                                  d3h_stream_t stream) {
  if (!tets || !bad_count_dev || n_tets < 0 || n_grid <= 0 || n_grid >= (1ll << 31) || n_tets >= (1ll << 31) ||
      (reinterpret_cast<uintptr_t>(tets) & 15)) {
    set_error("d3h_check_tets_i32: bad argument (null/misaligned pointer, or N/F outside [0, 2^31))");
    return D3H_E_BADARG;
  }
  if (n_tets == 0) return D3H_OK;
  int64_t blocks = (n_tets + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  check_tets_i32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int4*>(tets), n_tets, n_grid, reinterpret_cast<unsigned long long*>(bad_count_dev));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("d3h_check_tets_i32: %s", cudaGetErrorString(e)); return D3H_E_CUDA; }
  return D3H_OK;
}

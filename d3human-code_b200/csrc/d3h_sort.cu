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
//   emit_keys     : key = (min << bits) | max, value = corner slot; all radix histograms in the same pass
//   radix_pass    : LSD, 8-bit digits, one kernel per digit ("onesweep": per-tile digit counts chained by decoupled
//                   look-back; stable ranks from warp match_any + per-warp counters)
//   rle_interp    : head flags + scan (look-back) = vertex ids in sorted order; scatters ids to the corner array,
//                   writes the (a,b) tape and interpolates position / mSDF of every new vertex.
#include "d3h_internal.cuh"

namespace d3h {

int key_bits_for(int64_t n_grid) {
  int b = 1;
  while ((1ll << b) < n_grid) ++b;
  return b;
}

// ------------------------------------------------------------------------------------------------
// K2: keys + histograms
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
emit_keys_kernel(const d3h_tet_record* __restrict__ records, const DevCounters* __restrict__ ctr,
                 int64_t cap_records, int key_bits, int npass, unsigned long long* __restrict__ keys,
                 unsigned* __restrict__ vals, unsigned* __restrict__ radix_hist, unsigned* __restrict__ st_sort,
                 int64_t st_sort_pass_stride) {
  __shared__ unsigned s_hist[kMaxPasses * kRadix];
  for (int i = threadIdx.x; i < npass * kRadix; i += blockDim.x) s_hist[i] = 0u;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  // zero the look-back state of the radix passes that will actually run
  const unsigned t1 = ctr->work_tri, t2 = ctr->work_quad;
  {
    const int64_t ncorn = 3ll * t1 + 4ll * t2;
    const int64_t per_pass = ((ncorn + kSortTile - 1) / kSortTile) * kRadix;
    for (int64_t i = tid; i < per_pass * npass; i += nthreads)
      st_sort[(i / per_pass) * st_sort_pass_stride + (i % per_pass)] = 0u;
  }
  __syncthreads();
  const int64_t nvalid = (int64_t)t1 + t2;
  for (int64_t i = tid; i < nvalid; i += nthreads) {
    const int4 v4 = reinterpret_cast<const int4*>(records + i)[0];
    const int4 meta = reinterpret_cast<const int4*>(records + i)[1];
    const int code = meta.x, rank = meta.y;
    const int vv[4] = {v4.x, v4.y, v4.z, v4.w};
    const bool quad = __popc((unsigned)code) == 2;
    const int n = quad ? 4 : 3;
    const int64_t p0 = quad ? (3ll * t1 + 4ll * rank) : 3ll * rank;
    for (int k = 0; k < n; ++k) {
      const int e = c_loop_edge[code][k];
      const int p = vv[c_edge_p[e]], q = vv[c_edge_q[e]];
      const unsigned long long a = (unsigned)min(p, q), b = (unsigned)max(p, q);
      const unsigned long long key = (a << key_bits) | b;
      keys[p0 + k] = key;
      vals[p0 + k] = (unsigned)(p0 + k);
      for (int ps = 0; ps < npass; ++ps) atomicAdd(&s_hist[ps * kRadix + (unsigned)((key >> (8 * ps)) & 0xffu)], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npass * kRadix; i += blockDim.x) {
    const unsigned c = s_hist[i];
    if (c) atomicAdd(&radix_hist[i], c);
  }
}

// ------------------------------------------------------------------------------------------------
// radix pass (stable LSD, 8-bit digit)
// ------------------------------------------------------------------------------------------------
constexpr unsigned kSFlagAgg = 1u << 30, kSFlagInc = 2u << 30, kSValMask = (1u << 30) - 1;

__global__ void __launch_bounds__(kSortThreads)
radix_pass_kernel(const unsigned long long* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
                  unsigned long long* __restrict__ keys_out, unsigned* __restrict__ vals_out,
                  DevCounters* __restrict__ ctr, const unsigned* __restrict__ hist /* this pass, 256 */,
                  unsigned* __restrict__ status /* this pass: ntiles x 256 */, int pass, int shift) {
  constexpr int WARPS = kSortThreads / 32;
  __shared__ unsigned s_tile;
  __shared__ unsigned s_whist[WARPS][kRadix];  // per-warp digit counters, then exclusive offsets over warps
  __shared__ unsigned s_base[kRadix];          // global base of each digit for this tile
  __shared__ unsigned s_scan[WARPS];

  const int64_t ncorn = 3ll * ctr->work_tri + 4ll * ctr->work_quad;
  const int64_t ntiles = (ncorn + kSortTile - 1) / kSortTile;
  if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->ticket_sort[pass], 1u);
  for (int i = threadIdx.x; i < WARPS * kRadix; i += kSortThreads) (&s_whist[0][0])[i] = 0u;
  __syncthreads();
  const unsigned tile = s_tile;
  if ((int64_t)tile >= ntiles) return;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  const unsigned lt = lanemask_lt();

  // warp-striped tile: warp w owns keys [w*32*ITEMS, (w+1)*32*ITEMS), item j of lane l is at j*32 + l
  const int64_t wbase = (int64_t)tile * kSortTile + (int64_t)warp * (32 * kSortItems);
  unsigned long long key[kSortItems];
  unsigned val[kSortItems], rank[kSortItems];
  int digit[kSortItems];
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const int64_t idx = wbase + j * 32 + lane;
    const bool ok = idx < ncorn;
    key[j] = ok ? keys_in[idx] : ~0ull;
    val[j] = ok ? vals_in[idx] : 0u;
    digit[j] = ok ? (int)((key[j] >> shift) & 0xffu) : -1;
  }
  // stable rank inside the warp's chunk
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const int dmatch = digit[j] >= 0 ? digit[j] : (kRadix + (int)lane);  // padding lanes match nobody
    const unsigned peers = __match_any_sync(0xffffffffu, dmatch);
    const int leader = __ffs(peers) - 1;
    unsigned prev = 0;
    if (digit[j] >= 0 && (int)lane == leader) {
      prev = s_whist[warp][digit[j]];
      s_whist[warp][digit[j]] = prev + __popc(peers);
    }
    prev = __shfl_sync(0xffffffffu, prev, leader);
    rank[j] = prev + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();

  // thread d owns digit d: offsets over warps, tile count, global exclusive scan of the pass histogram, look-back
  {
    const int d = threadIdx.x;
    unsigned run = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      const unsigned c = s_whist[w][d];
      s_whist[w][d] = run;
      run += c;
    }
    const unsigned tile_count = run;
    // exclusive scan over digits of the global histogram (block scan of 256 values)
    const unsigned h = hist[d];
    unsigned incl = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (unsigned)o) incl += n;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    unsigned wpre = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w)
      if (w < (int)warp) wpre += s_scan[w];
    const unsigned digit_base = wpre + incl - h;

    unsigned excl = 0;
    unsigned* my = status + (int64_t)tile * kRadix + d;
    if (tile == 0) {
      st_relaxed_u32(my, kSFlagInc | tile_count);
    } else {
      st_relaxed_u32(my, kSFlagAgg | tile_count);
      int64_t look = (int64_t)tile - 1;
      while (true) {
        unsigned w;
        do { w = ld_relaxed_u32(status + look * kRadix + d); } while ((w >> 30) == 0u);
        excl += w & kSValMask;
        if ((w >> 30) == 2u || look == 0) break;
        --look;
      }
      st_relaxed_u32(my, kSFlagInc | (excl + tile_count));
    }
    s_base[d] = digit_base + excl;
  }
  __syncthreads();

#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    if (digit[j] < 0) continue;
    const unsigned pos = s_base[digit[j]] + s_whist[warp][digit[j]] + rank[j];
    keys_out[pos] = key[j];
    vals_out[pos] = val[j];
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
                  int64_t cap_verts, int64_t cap_verts_aug, float4* __restrict__ w_vert, float* __restrict__ w_acc,
                  float* __restrict__ verts_wt, float* __restrict__ msdf_wt, float* __restrict__ verts_aug,
                  float* __restrict__ msdf_aug) {
  constexpr int WARPS = kRleThreads / 32;
  __shared__ unsigned s_tile;
  __shared__ unsigned s_wsum[WARPS];
  __shared__ unsigned long long s_excl;

  const int64_t ncorn = 3ll * ctr->work_tri + 4ll * ctr->work_quad;
  const int64_t ntiles = (ncorn + kRleTile - 1) / kRleTile;
  if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->ticket_rle, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  if ((int64_t)tile >= ntiles) return;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;

  // blocked: thread owns 4 consecutive sorted keys
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
  // block exclusive scan of nhead
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
      float4* acc4 = reinterpret_cast<float4*>(w_acc + 8 * vid);
      acc4[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      acc4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
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
    tape_corners[vals[idx]] = (int32_t)vid;
  }
}

// ------------------------------------------------------------------------------------------------
void launch_edge_sort(const d3h_forward_args& a, const Workspace& ws, const d3h_tet_record* records,
                      cudaStream_t stream) {
  const int key_bits = key_bits_for(a.n_grid);
  const int npass = (2 * key_bits + kRadixBits - 1) / kRadixBits;
  const int64_t cap = ws.cap_tets;
  if (cap <= 0) return;
  {
    int64_t blocks = (cap + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    ProfScope ps(K_EMIT_KEYS, stream);
    emit_keys_kernel<<<(unsigned)blocks, 256, 0, stream>>>(records, ws.ctr, cap, key_bits, npass, ws.keys[0], ws.vals[0],
                                                           ws.radix_hist, ws.st_sort,
                                                           ws.ntiles_sort * (int64_t)kRadix);
  }
  int cur = 0;
  for (int p = 0; p < npass; ++p) {
    ProfScope ps(K_RADIX_PASS, stream);
    radix_pass_kernel<<<(unsigned)ws.ntiles_sort, kSortThreads, 0, stream>>>(
        ws.keys[cur], ws.vals[cur], ws.keys[cur ^ 1], ws.vals[cur ^ 1], ws.ctr, ws.radix_hist + p * kRadix,
        ws.st_sort + (int64_t)p * ws.ntiles_sort * kRadix, p, p * kRadixBits);
    cur ^= 1;
  }
  ProfScope ps(K_RLE_INTERP, stream);
  rle_interp_kernel<<<(unsigned)ws.ntiles_rle, kRleThreads, 0, stream>>>(
      ws.keys[cur], ws.vals[cur], ws.ctr, ws.st_rle, key_bits, a.pos, a.sdf, a.msdf, a.msdf_negate, a.tape_corners,
      a.tape_edges, a.cap_verts, a.cap_verts_aug, ws.vert, ws.acc, a.verts_wt, a.msdf_wt, a.verts_aug, a.msdf_aug);
}

}  // namespace d3h

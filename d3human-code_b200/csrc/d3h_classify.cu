// Stage 1 of the extraction: occupancy bitmap, the only O(F) kernel (streaming classification) and the ordered
// compaction of the valid tets.
//
// Replaces gshell_tets.py:260-275 (occ_n, occ_fx4, occ_sum, valid_tets), :307-309 (tetindex, num_triangles)
// and the boolean-mask compactions `tet_fx4[valid_tets]`, `idx_map[num_triangles == k]` (:277, :323-324).
//
//   prepare_kernel   N-sized: sign bitmap of sdf (N/8 bytes: 268 KB at 128^3, L1/L2 resident) + reset of scan state.
//   classify_kernel  F-sized, pure stream, one warp per 256-tet chunk: every warp loads 8 x 32 tets with 16-byte no-allocate
//                    loads, looks the four signs up in the bitmap, writes two ballot words per 32 tets (tet yields
//                    1 / 2 triangles) and adds its (T1,T2) count to the counter of its 8192-tet tile (one RED per
//                    non-empty warp chunk).  No barrier, no fence, no tail.
//   compact_kernel   one CTA per tile, no inter-CTA dependency: empty tiles (most of the grid) exit on the tile
//                    counter; the others sum the counters of the earlier tiles (1536 words at 128^3, from L2), rank
//                    their valid tets inside the tile and write the compact records and, in the fused single-GPU
//                    path, either the sort keys of their crossing edges + the MSD histogram (general path) or the marks
//                    of those edges in the bitmap over the grid's static edge table (mark_polygon_edges).
//
// History (profiles/): v1 classified and compacted in one kernel (ticket + block scan + look-back per 2048-tet tile):
// 45 % of the warp samples parked on the barrier behind the ticket atomic, 16 % DRAM utilisation.  v2 split the stream
// from a look-back compaction over 1536 tiles whose last CTA scanned the 131072-bucket MSD histogram serially: 128 us
// for 5 MB of traffic.  v3 (this file) has no look-back and no serial epilogue larger than the tile counters.
#include "d3h_internal.cuh"

namespace d3h {

int persistent_grid(const void* kernel, int threads, size_t dyn_smem) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess || per_sm <= 0)
    per_sm = 1;
  return sms * per_sm;
}

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

// block_only: a frame that takes its topology from the first frame of the launch needs its argument block and fresh
// counters, no bitmaps and no scan state
__device__ __forceinline__ void prepare_body(const FwdBlock& src, const Workspace& ws, bool block_only = false) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  if (block_only) {   // (one CTA)
    if (threadIdx.x == 0) *ws.blk = src;
    if (threadIdx.x < kCounterWordsReset) reinterpret_cast<unsigned*>(ws.ctr)[threadIdx.x] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) { ws.ctr->trace = src.trace; ws.ctr->trace_frame = (unsigned)src.a.seq; }
    return;
  }
  if (tid == 0) *ws.blk = src;  // the argument block of this call, for every later kernel
  const float* __restrict__ sdf = src.a.sdf;
  const float* __restrict__ msdf = src.a.msdf;
  const int64_t n_grid = src.a.n_grid;
  const int msdf_negate = src.a.msdf_negate;
  const int want_mocc = src.a.watertight_template ? 0 : 1;
  unsigned* __restrict__ occ_bits = ws.occ_bits;
  unsigned* __restrict__ mocc_bits = ws.mocc_bits;
  if (tid < (int64_t)kCounterWordsReset) reinterpret_cast<unsigned*>(ws.ctr)[tid] = 0u;
  if (tid == 0) { ws.ctr->trace = src.trace; ws.ctr->trace_frame = (unsigned)src.a.seq; }
  unsigned long long* tr = trace_begin(src.trace, (unsigned)src.a.seq, K_PREPARE);
  for (int64_t i = tid; i < ws.ntiles_compact; i += nthreads) ws.tile_cnt[i] = 0u;
  for (int64_t i = tid; i < ws.nscan_ctas; i += nthreads) ws.st_scan[i] = 0ull;
  for (int64_t i = tid; i < ws.ngroups / 256 + 1; i += nthreads) ws.gblock_heads[i] = 0u;
  if (ws.n_edges > 0 && src.a.edge_off != nullptr) {
    // static edge table path: clear the edge bitmap (n_edges / 8 bytes: 1.9 MB at 128^3) and the per-block counts
    const int64_t nw4 = ws.n_eblocks * (kEdgeBlock / 32) / 4;
    for (int64_t i = tid; i < nw4; i += nthreads) reinterpret_cast<uint4*>(ws.edge_bits)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t i = tid; i < ws.n_eblocks + 1; i += nthreads) ws.eblock_cnt[i] = 0u;
    if (src.a.etets != nullptr) {
      if (tid < 4 * kQueues) ws.q_cnt[kQStride * tid] = 0u;
      // edge-scan path: the tet bitmaps are MARKED (atomicOr) instead of written by the classification stream
      const int64_t nt4 = ws.nwords_tet / 4;   // nwords_tet is a multiple of 8
      for (int64_t i = tid; i < nt4; i += nthreads) {
        reinterpret_cast<uint4*>(ws.m1_words)[i] = make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4*>(ws.m2_words)[i] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  } else {
    for (int64_t i = tid; i < (ws.msd_bins + 8 + 3) / 4; i += nthreads)  // msd_hist: 256-byte aligned region, padded by 8
      reinterpret_cast<uint4*>(ws.msd_hist)[i] = make_uint4(0u, 0u, 0u, 0u);
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
  trace_end(tr);
}

__global__ void __launch_bounds__(256) prepare_kernel(FwdBlock src, Workspace ws) { prepare_body(src, ws); }

// several frames in one launch: frame blockIdx.y, its argument block from the set, its workspace by offset
__global__ void __launch_bounds__(256) prepare_frames_kernel(const __grid_constant__ FwdBlockSet set, Workspace ws,
                                                             const __grid_constant__ FrameSet fs, int shared_topology) {
  if (shared_topology < 0) {
    // grid.y == 1, -shared_topology frames: the topology of the launch before still stands in the first frame's workspace
    // (BatchCtx.reuse_topology).  CTA f stores the argument block of frame f, resets its counters and takes the totals from
    // the counts that launch published in that workspace (its own frame there had the same totals).
    if ((int)blockIdx.x < -shared_topology) {
      const d3h_counts* __restrict__ tc = ws.counts;
      Workspace w2 = ws;
      shift_workspace(w2, fs.off[blockIdx.x]);
      prepare_body(set.f[blockIdx.x], w2, true);
      if (threadIdx.x == 0) {
        const bool ok = tc->overflow == 0;
        w2.ctr->n_valid = (unsigned)tc->n_valid_tets;
        w2.ctr->n_tri = (unsigned)tc->n_tri_tets;
        w2.ctr->n_quad = (unsigned)tc->n_quad_tets;
        w2.ctr->work_tri = ok ? (unsigned)tc->n_tri_tets : 0u;
        w2.ctr->work_quad = ok ? (unsigned)tc->n_quad_tets : 0u;
        w2.ctr->n_verts = (unsigned)tc->n_verts;
      }
    }
    return;
  }
  if (shared_topology) {
    // grid.y == 1: the first frame has the bitmaps to build and the scan state to clear; CTA f also stores the argument
    // block of frame f and resets its counters (all that a frame needs which takes its topology from the first)
    if (blockIdx.x >= 1 && (int)blockIdx.x < shared_topology) {
      Workspace w2 = ws;
      shift_workspace(w2, fs.off[blockIdx.x]);
      prepare_body(set.f[blockIdx.x], w2, true);
    }
    prepare_body(set.f[0], ws);
    return;
  }
  shift_workspace(ws, fs.off[blockIdx.y]);
  prepare_body(set.f[blockIdx.y], ws);
}

const void* prepare_kernel_address() { return reinterpret_cast<const void*>(prepare_kernel); }

void launch_prepare(const d3h_forward_args& a, const Workspace& ws, cudaStream_t stream) {
  const int64_t nquads = (a.n_grid + 3) / 4;
  int64_t blocks = (nquads + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  FwdBlock blk;
  blk.a = a;
  blk.counts_mapped = mapped_counts_pointer(a.counts_host);
  blk.trace = trace_table();
  ProfScope ps(K_PREPARE, stream);
  launch_k(prepare_kernel, (unsigned)blocks, 256, stream, kLaunchLatency, blk, ws);
}

// the frames of batch_ctx() in one launch; `a` are their argument structs, `ws` is the workspace of the first
void launch_prepare_frames(const d3h_forward_args* a, const Workspace& ws, cudaStream_t stream) {
  const int frames = batch_ctx().frames;
  const int64_t nquads = (a[0].n_grid + 3) / 4;
  int64_t blocks = (nquads + 255) / 256;
  if (blocks < 1) blocks = 1;
  // with a shared topology only the first frame has bitmaps to build and state to clear: it keeps the whole grid (the
  // CTAs of the other frames store their argument block and leave)
  const bool shared = batch_ctx().topo_frames == 1 && frames > 1;
  if (!shared && blocks > 148 * 8 / frames) blocks = 148 * 8 / frames;
  static thread_local FwdBlockSet set;
  for (int f = 0; f < frames; ++f) {
    set.f[f].a = a[f];
    set.f[f].counts_mapped = mapped_counts_pointer(a[f].counts_host);
    set.f[f].trace = trace_table();
  }
  ProfScope ps(K_PREPARE, stream);
  if (shared && batch_ctx().reuse_topology) {   // a CTA per frame
    batch_ctx().frames = 1;
    launch_k(prepare_frames_kernel, (unsigned)frames, 256, stream, kLaunchLatency, set, ws, batch_ctx().fs, -frames);
    batch_ctx().frames = frames;
  } else if (shared) {   // one frame's worth of CTAs (at least one per frame), shared_topology = number of frames
    if (blocks < frames) blocks = frames;
    batch_ctx().frames = 1;
    launch_k(prepare_frames_kernel, (unsigned)blocks, 256, stream, kLaunchLatency, set, ws, batch_ctx().fs, frames);
    batch_ctx().frames = frames;
  } else {
    launch_k(prepare_frames_kernel, (unsigned)blocks, 256, stream, kLaunchLatency, set, ws, batch_ctx().fs, 0);
  }
}

// ------------------------------------------------------------------------------------------------
// K1: streaming classification (persistent grid)
// ------------------------------------------------------------------------------------------------

// One warp classifies kChunkTets consecutive tets per loop trip.  MOCC adds the open-mesh prefilter (gshell_tets.py:275).
// A tet past the end of the range is loaded as (0,0,0,0): four equal vertices are never a sign change, so the tail needs
// no per-tet bounds test.  c = number of occupied vertices: c odd -> one triangle, c == 2 -> two, c in {0,4} -> none.
template <bool MOCC>
__global__ void __launch_bounds__(kClassifyThreads)
classify_kernel(const FwdBlock* __restrict__ blk, int64_t tet_begin, int64_t tet_end,
                const unsigned* __restrict__ occ_bits, const unsigned* __restrict__ mocc_bits,
                unsigned* __restrict__ m1_words, unsigned* __restrict__ m2_words, unsigned* __restrict__ tile_cnt,
                int64_t nchunks) {
  const int4* __restrict__ tets = reinterpret_cast<const int4*>(blk->a.tets);
  unsigned long long* tr = trace_begin(blk->trace, (unsigned)blk->a.seq, K_CLASSIFY);
  const unsigned lane = lane_id();
  const int64_t warps_total = (int64_t)gridDim.x * (kClassifyThreads / 32);
  for (int64_t chunk = ((int64_t)blockIdx.x * kClassifyThreads + threadIdx.x) >> 5; chunk < nchunks;
       chunk += warps_total) {
    const int64_t base = tet_begin + chunk * kChunkTets;
    const int4* p = tets + base + lane;
    int4 t[kClassifyItems];
    if (base + kChunkTets <= tet_end) {
#pragma unroll
      for (int j = 0; j < kClassifyItems; ++j) t[j] = ld_stream_int4(p + j * 32);
    } else {
#pragma unroll
      for (int j = 0; j < kClassifyItems; ++j)
        t[j] = (base + j * 32 + lane < tet_end) ? ld_stream_int4(p + j * 32) : make_int4(0, 0, 0, 0);
    }
    unsigned w1 = 0, w2 = 0;  // lane j keeps the ballot words of item j
#pragma unroll
    for (int j = 0; j < kClassifyItems; ++j) {
      const unsigned c = occ_of(occ_bits, t[j].x) + occ_of(occ_bits, t[j].y) + occ_of(occ_bits, t[j].z) +
                         occ_of(occ_bits, t[j].w);
      bool tri = (c & 1u) != 0u, quad = (c == 2u);
      if (MOCC) {
        const bool keep = (occ_of(mocc_bits, t[j].x) | occ_of(mocc_bits, t[j].y) | occ_of(mocc_bits, t[j].z) |
                           occ_of(mocc_bits, t[j].w)) != 0u;
        tri = tri && keep;
        quad = quad && keep;
      }
      const unsigned b1 = __ballot_sync(0xffffffffu, tri);
      const unsigned b2 = __ballot_sync(0xffffffffu, quad);
      if (lane == (unsigned)j) { w1 = b1; w2 = b2; }
    }
    if (lane < (unsigned)kClassifyItems) {
      const int64_t w = chunk * kClassifyItems + lane;
      m1_words[w] = w1;
      m2_words[w] = w2;
    }
    unsigned cnt = __popc(w1) | (__popc(w2) << 16);  // lanes >= kClassifyItems hold 0
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
    if (lane == 0 && cnt != 0u) atomicAdd(tile_cnt + (chunk * kChunkTets) / kTileTets, cnt);
  }
  trace_end(tr);
}

// ------------------------------------------------------------------------------------------------
// K1b: ordered compaction of the valid tets (+ fused key emission)
// ------------------------------------------------------------------------------------------------
// Writes the sort keys of one valid tet: one key per polygon corner (= crossing edge, in mesh_edge_table order).
// Slots are dense in valid-tet order (3 per tri tet, 4 per quad tet); the value encodes the final corner slot
// [3*T1 | 4*T2] as (class, 4*class_rank + k) because T1 is not known to every CTA yet.
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

// Static edge table path: the rank of edge (a,b) in the sorted list of all tet edges of the grid is found by bisecting
// the (at most ~14) larger neighbours of a; the edge is marked in the bitmap and the first marker counts it for its
// 8192-edge block, so that the numbering kernel needs no scan over the blocks.
// Variant with the per-tet rank table (d3h_forward_args.tet_edge_rank): no bisection, two 16-byte loads per valid tet.
__device__ __forceinline__ void mark_polygon_edges_ranked(int64_t tet, int code, bool quad, int64_t rec,
                                                          const int32_t* __restrict__ tet_edge_rank,
                                                          unsigned* __restrict__ edge_bits,
                                                          unsigned* __restrict__ eblock_cnt,
                                                          unsigned* __restrict__ corner_rank) {
  const int4 r03 = __ldg(reinterpret_cast<const int4*>(tet_edge_rank) + 2 * tet);
  const int4 r45 = __ldg(reinterpret_cast<const int4*>(tet_edge_rank) + 2 * tet + 1);
  const int rr[6] = {r03.x, r03.y, r03.z, r03.w, r45.x, r45.y};
  const int n = quad ? 4 : 3;
  unsigned rk[4], old[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = c_loop_edge[code][k < n ? k : 0];
    unsigned r = 0;
#pragma unroll
    for (int q = 0; q < 6; ++q)
      if (q == e) r = (unsigned)rr[q];      // select without dynamic register indexing
    rk[k] = r;
    old[k] = 0xffffffffu;
    if (k < n) old[k] = atomicOr(edge_bits + (r >> 5), 1u << (r & 31u));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < n && !(old[k] & (1u << (rk[k] & 31u)))) atomicAdd(eblock_cnt + rk[k] / kEdgeBlock, 1u);
  reinterpret_cast<uint4*>(corner_rank)[rec] = make_uint4(rk[0], rk[1], rk[2], n == 4 ? rk[3] : 0xffffffffu);
}

__device__ __forceinline__ void mark_polygon_edges(const int4 v4, int code, bool quad, int64_t rec,
                                                   const int32_t* __restrict__ edge_off,
                                                   const int2* __restrict__ edge_ab, unsigned* __restrict__ edge_bits,
                                                   unsigned* __restrict__ eblock_cnt,
                                                   unsigned* __restrict__ corner_rank) {
  const int vv[4] = {v4.x, v4.y, v4.z, v4.w};
  const int n = quad ? 4 : 3;
  // the four corners advance in lockstep so that their loads are in flight together (the chains are independent)
  int bb[4], lo[4], hi[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = c_loop_edge[code][k < n ? k : 0];
    const int p = vv[c_edge_p[e]], q = vv[c_edge_q[e]];
    const int a = min(p, q);
    bb[k] = max(p, q);
    lo[k] = __ldg(edge_off + a);
    hi[k] = __ldg(edge_off + a + 1) - 1;
  }
  bool more = true;
  while (more) {  // first entry with .y >= b (the edge exists: it is an edge of this very tet)
    int mid[4], y[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mid[k] = (lo[k] + hi[k]) >> 1;
      y[k] = __ldg(&edge_ab[mid[k]].y);
    }
    more = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (lo[k] < hi[k]) {
        if (y[k] < bb[k]) lo[k] = mid[k] + 1; else hi[k] = mid[k];
      }
      more = more || (lo[k] < hi[k]);
    }
  }
  unsigned old[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    old[k] = 0xffffffffu;
    if (k < n) old[k] = atomicOr(edge_bits + ((unsigned)lo[k] >> 5), 1u << ((unsigned)lo[k] & 31u));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < n && !(old[k] & (1u << ((unsigned)lo[k] & 31u)))) atomicAdd(eblock_cnt + (unsigned)lo[k] / kEdgeBlock, 1u);
  reinterpret_cast<uint4*>(corner_rank)[rec] =
      make_uint4((unsigned)lo[0], (unsigned)lo[1], (unsigned)lo[2], n == 4 ? (unsigned)lo[3] : 0xffffffffu);
}

// EMIT: 0 records only (tet-range shards), 1 sort keys + MSD histogram (general path), 2 static edge table marks,
// 3 the same with the per-tet rank table (experimental)
template <int EMIT>
__global__ void __launch_bounds__(kCompactThreads)
compact_kernel(const FwdBlock* __restrict__ blk, const unsigned* __restrict__ m1_words,
               const unsigned* __restrict__ m2_words, int64_t nwords, int64_t tet_begin,
               const unsigned* __restrict__ occ_bits, const unsigned* __restrict__ tile_cnt, int64_t ntiles,
               DevCounters* __restrict__ ctr, d3h_tet_record* __restrict__ records, int64_t cap_records, int key_bits,
               int msd_shift, unsigned long long* __restrict__ keys, unsigned* __restrict__ vals,
               unsigned* __restrict__ msd_hist, unsigned* __restrict__ edge_bits, unsigned* __restrict__ eblock_cnt,
               unsigned* __restrict__ corner_rank) {
  constexpr int WARPS = kCompactThreads / 32;
  const int4* __restrict__ tets = reinterpret_cast<const int4*>(blk->a.tets);
  __shared__ unsigned s_pre[kCompactThreads];  // exclusive per-thread prefix in the tile: T1 | T2 << 16
  __shared__ unsigned s_m1[kCompactThreads], s_m2[kCompactThreads];
  __shared__ unsigned s_w[WARPS];
  __shared__ unsigned long long s_sum[WARPS];

  const unsigned tile = blockIdx.x;
  const unsigned tc = __ldcg(tile_cnt + tile);
  const bool is_last = (int64_t)tile == ntiles - 1;
  unsigned long long* tr = trace_begin(ctr->trace, ctr->trace_frame, K_COMPACT);
  if (tc == 0u && !is_last) return;  // most tiles hold no surface
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;

  // (T1,T2) valid tets in all earlier tiles: every non-empty tile sums the tile counters itself (a few KB from L2,
  // all loads independent) -- no scan kernel, no look-back chain.  The last tile also publishes the grid totals.
  unsigned long long sum = 0ull;  // T1 in the low half, T2 in the high half
  for (unsigned i = threadIdx.x; i < tile; i += kCompactThreads) {
    const unsigned c = __ldcg(tile_cnt + i);
    sum += (unsigned long long)(c & 0xffffu) | ((unsigned long long)(c >> 16) << 32);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_sum[warp] = sum;

  // each thread owns one word of both bitmaps (= 32 consecutive tets)
  const int64_t w0 = (int64_t)tile * kCompactThreads + threadIdx.x;
  const unsigned a1 = (w0 < nwords) ? __ldcg(m1_words + w0) : 0u;
  const unsigned a2 = (w0 < nwords) ? __ldcg(m2_words + w0) : 0u;
  s_m1[threadIdx.x] = a1;
  s_m2[threadIdx.x] = a2;
  const unsigned c = __popc(a1) | (__popc(a2) << 16);
  unsigned incl = c;  // both halves stay below 2^16 (a tile holds 8192 tets)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (unsigned)o) incl += n;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  unsigned wpre = 0;
  unsigned long long excl = 0ull;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) {
    if (w < (int)warp) wpre += s_w[w];
    excl += s_sum[w];
  }
  s_pre[threadIdx.x] = wpre + incl - c;
  const unsigned e1 = (unsigned)(excl & 0xffffffffull), e2 = (unsigned)(excl >> 32);
  if (is_last && threadIdx.x == 0) {
    const unsigned t1 = e1 + (tc & 0xffffu), t2 = e2 + (tc >> 16);
    ctr->n_tri = t1;
    ctr->n_quad = t2;
    ctr->n_valid = t1 + t2;
    const bool fits = (int64_t)t1 + t2 <= cap_records;
    ctr->work_tri = fits ? t1 : 0u;
    ctr->work_quad = fits ? t2 : 0u;
  }
  if (tc == 0u) return;
  __syncthreads();

  // ---- visit the tile's valid tets in tet order: up to kRounds per thread at a time, so that the dependent
  // chain (index load -> bitmap look-ups -> stores) of several tets is in flight together ----
  const unsigned nvalid_tile = (tc & 0xffffu) + (tc >> 16);
  constexpr int kRounds = 4;
  for (unsigned i0 = threadIdx.x; i0 < nvalid_tile; i0 += kRounds * kCompactThreads) {
    int4 v4[kRounds];
    unsigned g1[kRounds], g2[kRounds];
    int64_t tet[kRounds];
    bool quad[kRounds], live[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      const unsigned i = i0 + r * kCompactThreads;
      live[r] = i < nvalid_tile;
      v4[r] = make_int4(0, 0, 0, 0);
      g1[r] = g2[r] = 0;
      tet[r] = 0;
      quad[r] = false;
      if (live[r]) {
        int lo = 0, hi = kCompactThreads - 1;  // owner thread: last th with pre1[th] + pre2[th] <= i
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          const unsigned pm = s_pre[mid];
          if ((pm & 0xffffu) + (pm >> 16) <= i) lo = mid; else hi = mid - 1;
        }
        const int th = lo;
        const unsigned pt = s_pre[th];
        unsigned r1 = pt & 0xffffu, r2 = pt >> 16;
        const unsigned rem = i - (r1 + r2);
        const unsigned b1 = s_m1[th], b2 = s_m2[th];
        const int bit = (int)__fns(b1 | b2, 0, (int)rem + 1);  // position of the (rem+1)-th set bit
        const unsigned below = (1u << bit) - 1u;
        r1 += __popc(b1 & below);
        r2 += __popc(b2 & below);
        quad[r] = (b2 >> bit) & 1u;
        g1[r] = e1 + r1;  // tri / quad valid tets before this one, grid-wide
        g2[r] = e2 + r2;
        tet[r] = tet_begin + ((int64_t)tile * kCompactThreads + th) * 32 + bit;
        live[r] = (int64_t)g1[r] + g2[r] < cap_records;
        if (live[r]) v4[r] = __ldg(tets + tet[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      if (!live[r]) continue;
      const int code = (int)(occ_of(occ_bits, v4[r].x) | (occ_of(occ_bits, v4[r].y) << 1) |
                             (occ_of(occ_bits, v4[r].z) << 2) | (occ_of(occ_bits, v4[r].w) << 3));
      const unsigned cr = quad[r] ? g2[r] : g1[r], ob = quad[r] ? g1[r] : g2[r];
      int4* out = reinterpret_cast<int4*>(records + ((int64_t)g1[r] + g2[r]));
      out[0] = v4[r];
      out[1] = make_int4(code, (int)cr, (int)tet[r], (int)ob);
      if (EMIT == 1) emit_polygon_keys(v4[r], code, quad[r], cr, ob, key_bits, msd_shift, keys, vals, msd_hist);
      if (EMIT == 2)
        mark_polygon_edges(v4[r], code, quad[r], (int64_t)g1[r] + g2[r], blk->a.edge_off,
                           reinterpret_cast<const int2*>(blk->a.edge_ab), edge_bits, eblock_cnt, corner_rank);
      if (EMIT == 3)
        mark_polygon_edges_ranked(tet[r], code, quad[r], (int64_t)g1[r] + g2[r], blk->a.tet_edge_rank, edge_bits,
                                  eblock_cnt, corner_rank);
    }
  }
  trace_end(tr);
}

void launch_classify(const d3h_forward_args& a, const Workspace& ws, d3h_tet_record* records, int64_t cap_records,
                     bool emit_keys, cudaStream_t stream, int parts) {
  const int64_t n = a.tet_end - a.tet_begin;
  if (n <= 0) return;
  const int64_t nchunks = (n + kChunkTets - 1) / kChunkTets;
  const int64_t ntiles = (n + kTileTets - 1) / kTileTets;
  if (parts & kPartHead) {
    // One-shot grid (a warp per 256-tet chunk, CTAs retire every few microseconds) at the lowest priority: measured as
    // fast as a persistent grid (profiles/bench_stream.cu) and, unlike it, it leaves CTA slots to the latency-bound
    // kernels of the frames running on the other lanes.
    const int mocc = a.watertight_template ? 0 : 1;
    const int64_t nblocks = (nchunks * 32 + kClassifyThreads - 1) / kClassifyThreads;
    ProfScope ps(K_CLASSIFY, stream);
    if (mocc)
      launch_k(classify_kernel<true>, (unsigned)nblocks, kClassifyThreads, stream, kLaunchStream, ws.blk, a.tet_begin,
               a.tet_end, ws.occ_bits, ws.mocc_bits, ws.m1_words, ws.m2_words, ws.tile_cnt, nchunks);
    else
      launch_k(classify_kernel<false>, (unsigned)nblocks, kClassifyThreads, stream, kLaunchStream, ws.blk, a.tet_begin,
               a.tet_end, ws.occ_bits, (const unsigned*)nullptr, ws.m1_words, ws.m2_words, ws.tile_cnt, nchunks);
  }
  if (!(parts & kPartTail)) return;
  const int64_t nwords = nchunks * kClassifyItems;  // every word of a visited chunk is written
  const int key_bits = key_bits_for(a.n_grid);
  const int msd_shift = msd_shift_for(a.n_grid);
  ProfScope ps(K_COMPACT, stream);
  unsigned long long* const nokeys = nullptr;
  unsigned* const nou = nullptr;
  if (emit_keys && a.edge_off != nullptr && a.tet_edge_rank != nullptr)
    launch_k(compact_kernel<3>, (unsigned)ntiles, kCompactThreads, stream, kLaunchLatency, ws.blk, ws.m1_words,
             ws.m2_words, nwords, a.tet_begin, ws.occ_bits, ws.tile_cnt, ntiles, ws.ctr, records, cap_records, key_bits,
             msd_shift, nokeys, nou, nou, ws.edge_bits, ws.eblock_cnt, ws.corner_rank);
  else if (emit_keys && a.edge_off != nullptr)
    launch_k(compact_kernel<2>, (unsigned)ntiles, kCompactThreads, stream, kLaunchLatency, ws.blk, ws.m1_words,
             ws.m2_words, nwords, a.tet_begin, ws.occ_bits, ws.tile_cnt, ntiles, ws.ctr, records, cap_records, key_bits,
             msd_shift, nokeys, nou, nou, ws.edge_bits, ws.eblock_cnt, ws.corner_rank);
  else if (emit_keys)
    launch_k(compact_kernel<1>, (unsigned)ntiles, kCompactThreads, stream, kLaunchLatency, ws.blk, ws.m1_words,
             ws.m2_words, nwords, a.tet_begin, ws.occ_bits, ws.tile_cnt, ntiles, ws.ctr, records, cap_records, key_bits,
             msd_shift, ws.keys, ws.vals, ws.msd_hist, nou, nou, nou);
  else
    launch_k(compact_kernel<0>, (unsigned)ntiles, kCompactThreads, stream, kLaunchLatency, ws.blk, ws.m1_words,
             ws.m2_words, nwords, a.tet_begin, ws.occ_bits, ws.tile_cnt, ntiles, ws.ctr, records, cap_records, key_bits,
             msd_shift, nokeys, nou, nou, nou, nou, nou);
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
keys_from_records_kernel(const d3h_tet_record* __restrict__ records, const DevCounters* __restrict__ ctr, int key_bits,
                         int msd_shift, unsigned long long* __restrict__ keys, unsigned* __restrict__ vals,
                         unsigned* __restrict__ msd_hist) {
  const int64_t n = (int64_t)ctr->work_tri + ctr->work_quad;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int4 v4 = reinterpret_cast<const int4*>(records + i)[0];
    const int4 meta = reinterpret_cast<const int4*>(records + i)[1];
    emit_polygon_keys(v4, meta.x, __popc((unsigned)meta.x) == 2, (unsigned)meta.y, (unsigned)meta.w, key_bits, msd_shift,
                      keys, vals, msd_hist);
  }
}

void launch_rank_records(const d3h_forward_args& a, const Workspace& ws, d3h_tet_record* records, int64_t n_records,
                         cudaStream_t stream) {
  {
    ProfScope ps(K_RANK_RECORDS, stream);
    launch_k(rank_records_kernel, 1u, 1024u, stream, kLaunchLatency, records, n_records, ws.ctr);
  }
  int64_t blocks = (n_records + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 2) blocks = 148 * 2;
  ProfScope ps(K_COMPACT, stream);
  launch_k(keys_from_records_kernel, (unsigned)blocks, 256u, stream, kLaunchLatency, records, ws.ctr,
           key_bits_for(a.n_grid), msd_shift_for(a.n_grid), ws.keys, ws.vals, ws.msd_hist);
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
#ifndef D3H_CPU_EMU
  pack_tets_i64_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const longlong2*>(tets), n_tets, n_grid, reinterpret_cast<int4*>(out_tets),
      reinterpret_cast<unsigned long long*>(bad_count_dev));
#else   // tests/emu: g++ has no <<< >>>
  launch_k(pack_tets_i64_kernel, (unsigned)blocks, 256u, (cudaStream_t)stream, kLaunchStream,
           reinterpret_cast<const longlong2*>(tets), n_tets, n_grid, reinterpret_cast<int4*>(out_tets),
           reinterpret_cast<unsigned long long*>(bad_count_dev));
#endif
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
#ifndef D3H_CPU_EMU
  check_tets_i32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int4*>(tets), n_tets, n_grid, reinterpret_cast<unsigned long long*>(bad_count_dev));
#else   // tests/emu: g++ has no <<< >>>
  launch_k(check_tets_i32_kernel, (unsigned)blocks, 256u, (cudaStream_t)stream, kLaunchStream,
           reinterpret_cast<const int4*>(tets), n_tets, n_grid, reinterpret_cast<unsigned long long*>(bad_count_dev));
#endif
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("d3h_check_tets_i32: %s", cudaGetErrorString(e)); return D3H_E_CUDA; }
  return D3H_OK;
}

// Stage 1 of the extraction: occupancy bitmap + the only O(F) kernel (classify + ordered compaction).
//
// Replaces gshell_tets.py:260-275 (occ_n, occ_fx4, occ_sum, valid_tets), :307-309 (tetindex, num_triangles)
// and the boolean-mask compactions `tet_fx4[valid_tets]`, `idx_map[num_triangles == k]` (:277, :323-324).
//
// HBM traffic: 16 B per tet (one int32x4 load, streamed once, no L1 allocation) + 4 B per grid vertex for the
// bitmap build.  The four per-tet sign lookups hit a N/8-byte bitmap (268 KB at 128^3) that stays in L1/L2.
// Ordered compaction uses warp ballots, one block scan of per-warp counts and a decoupled look-back over
// tile aggregates (single pass over the tet stream; status words are self-contained so no fences are needed).
#include "d3h_internal.cuh"

namespace d3h {

// ------------------------------------------------------------------------------------------------
// K0: occupancy bitmaps + reset of all per-call scan state
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prepare_kernel(const float* __restrict__ sdf, const float* __restrict__ msdf,
                                                      int64_t n_grid, int msdf_negate, int want_mocc,
                                                      unsigned* __restrict__ occ_bits,
                                                      unsigned* __restrict__ mocc_bits, Workspace ws) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  // reset (tiny): counters, classify / rle / poly status words, radix histograms
  if (tid < (int64_t)(sizeof(DevCounters) / 4)) reinterpret_cast<unsigned*>(ws.ctr)[tid] = 0u;
  for (int64_t i = tid; i < ws.ntiles_classify; i += nthreads) ws.st_classify[i] = 0ull;
  for (int64_t i = tid; i < ws.ntiles_rle; i += nthreads) ws.st_rle[i] = 0ull;
  for (int64_t i = tid; i < ws.ntiles_poly * 8; i += nthreads) ws.st_poly[i] = 0u;
  for (int64_t i = tid; i < kMaxPasses * kRadix; i += nthreads) ws.radix_hist[i] = 0u;

  // bitmaps: one 32-bit word per warp iteration, coalesced 128 B reads
  const int64_t nwords = (n_grid + 31) / 32;
  const int64_t warp = tid >> 5, nwarps = nthreads >> 5;
  const unsigned lane = lane_id();
  for (int64_t w = warp; w < nwords; w += nwarps) {
    const int64_t v = w * 32 + lane;
    float s = (v < n_grid) ? __ldg(sdf + v) : 0.f;
    unsigned word = __ballot_sync(0xffffffffu, s > 0.f);
    if (lane == 0) occ_bits[w] = word;
    if (want_mocc) {
      float m = (v < n_grid) ? __ldg(msdf + v) : 0.f;
      if (msdf_negate) m = -m;
      unsigned mw = __ballot_sync(0xffffffffu, m > 0.f);
      if (lane == 0) mocc_bits[w] = mw;
    }
  }
}

void launch_prepare(const d3h_forward_args& a, const Workspace& ws, cudaStream_t stream) {
  const int64_t nwords = (a.n_grid + 31) / 32;
  int64_t blocks = (nwords * 32 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ProfScope ps(K_PREPARE, stream);
  prepare_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a.sdf, a.msdf, a.n_grid, a.msdf_negate,
                                                         a.watertight_template ? 0 : 1, ws.occ_bits, ws.mocc_bits, ws);
}

// ------------------------------------------------------------------------------------------------
// K1: classify + ordered compaction
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long kFlagAgg = 1ull << 62;  // tile aggregate published
constexpr unsigned long long kFlagInc = 2ull << 62;  // inclusive prefix published
constexpr unsigned long long kValMask = (1ull << 62) - 1;
// value = count(T1 class) in bits [0,31) | count(T2 class) in bits [31,62)

__device__ __forceinline__ unsigned occ_of(const unsigned* __restrict__ bits, int v) {
  return (__ldg(bits + (v >> 5)) >> (v & 31)) & 1u;
}

__global__ void __launch_bounds__(kClassifyThreads)
classify_kernel(const int4* __restrict__ tets, int64_t tet_begin, int64_t tet_end,
                const unsigned* __restrict__ occ_bits, const unsigned* __restrict__ mocc_bits,
                unsigned long long* __restrict__ status, DevCounters* __restrict__ ctr,
                d3h_tet_record* __restrict__ records, int64_t cap_records, int64_t ntiles) {
  constexpr int WARPS = kClassifyThreads / 32;
  __shared__ unsigned s_tile;
  __shared__ unsigned s_seg[kClassifyItems * WARPS];  // packed per-(item,warp) counts: T1 | T2 << 16
  __shared__ unsigned long long s_excl;

  if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->ticket_classify, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  if ((int64_t)tile >= ntiles) return;
  const int64_t base = tet_begin + (int64_t)tile * kClassifyTile;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;

  // 8 independent 16-byte streaming loads per thread
  int4 t[kClassifyItems];
#pragma unroll
  for (int j = 0; j < kClassifyItems; ++j) {
    const int64_t idx = base + j * kClassifyThreads + threadIdx.x;
    t[j] = (idx < tet_end) ? ld_stream_int4(tets + idx) : make_int4(0, 0, 0, 0);
  }
  unsigned code[kClassifyItems], m1[kClassifyItems], m2[kClassifyItems];
#pragma unroll
  for (int j = 0; j < kClassifyItems; ++j) {
    const int64_t idx = base + j * kClassifyThreads + threadIdx.x;
    unsigned c = occ_of(occ_bits, t[j].x) | (occ_of(occ_bits, t[j].y) << 1) | (occ_of(occ_bits, t[j].z) << 2) |
                 (occ_of(occ_bits, t[j].w) << 3);
    if (idx >= tet_end) c = 0u;
    int nocc = __popc(c);
    bool valid = (nocc != 0) && (nocc != 4);
    if (mocc_bits != nullptr && valid) {  // open-mesh prefilter, gshell_tets.py:275
      unsigned any_m = occ_of(mocc_bits, t[j].x) | occ_of(mocc_bits, t[j].y) | occ_of(mocc_bits, t[j].z) |
                       occ_of(mocc_bits, t[j].w);
      valid = any_m != 0u;
    }
    code[j] = valid ? c : 0u;
    m1[j] = __ballot_sync(0xffffffffu, valid && (nocc != 2));  // 1 or 3 occupied -> one triangle
    m2[j] = __ballot_sync(0xffffffffu, valid && (nocc == 2));  // 2 occupied      -> two triangles (quad)
    if (lane == 0) s_seg[j * WARPS + warp] = (unsigned)__popc(m1[j]) | ((unsigned)__popc(m2[j]) << 16);
  }
  __syncthreads();

  if (warp == 0) {
    // exclusive scan of the 64 segment counts (2 per lane), order = (item, warp) = tet order
    constexpr int NSEG = kClassifyItems * WARPS;
    static_assert(NSEG == 64, "two segments per lane");
    unsigned a0 = s_seg[2 * lane], a1 = s_seg[2 * lane + 1];
    unsigned sum = a0 + a1, incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (unsigned)o) incl += n;
    }
    unsigned excl = incl - sum;
    s_seg[2 * lane] = excl;
    s_seg[2 * lane + 1] = excl + a0;
    unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned long long agg = (unsigned long long)(total & 0xffffu) | ((unsigned long long)(total >> 16) << 31);

    // decoupled look-back, 32 predecessors per step
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

  const unsigned long long ex = s_excl;
  const unsigned e1 = (unsigned)(ex & 0x7fffffffull), e2 = (unsigned)(ex >> 31);
  const unsigned lt = lanemask_lt();
#pragma unroll
  for (int j = 0; j < kClassifyItems; ++j) {
    if (code[j] == 0u) continue;
    const unsigned seg = s_seg[j * WARPS + warp];
    const unsigned r1 = e1 + (seg & 0xffffu) + __popc(m1[j] & lt);
    const unsigned r2 = e2 + (seg >> 16) + __popc(m2[j] & lt);
    const bool quad = (m2[j] >> lane) & 1u;
    const int64_t slot = (int64_t)r1 + r2;  // rank among all valid tets = tet order
    if (slot < cap_records) {
      const int64_t idx = base + j * kClassifyThreads + threadIdx.x;
      int4* out = reinterpret_cast<int4*>(records + slot);
      out[0] = t[j];
      out[1] = make_int4((int)code[j], (int)(quad ? r2 : r1), (int)idx, 0);
    }
  }
}

void launch_classify(const d3h_forward_args& a, const Workspace& ws, d3h_tet_record* records, int64_t cap_records,
                     cudaStream_t stream) {
  const int64_t n = a.tet_end - a.tet_begin;
  const int64_t ntiles = (n + kClassifyTile - 1) / kClassifyTile;
  if (ntiles <= 0) return;
  ProfScope ps(K_CLASSIFY, stream);
  classify_kernel<<<(unsigned)ntiles, kClassifyThreads, 0, stream>>>(
      reinterpret_cast<const int4*>(a.tets), a.tet_begin, a.tet_end, ws.occ_bits,
      a.watertight_template ? nullptr : ws.mocc_bits, ws.st_classify, ws.ctr, records, cap_records, ntiles);
}

// ------------------------------------------------------------------------------------------------
// records gathered from several shards: recompute the class ranks over the concatenation
// (single block; the merged list is O(surface))
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
    if (cls == 1) records[i].class_rank = (int)(s_run1 + p1 + __popc(b1 & lt));
    if (cls == 2) records[i].class_rank = (int)(s_run2 + p2 + __popc(b2 & lt));
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

void launch_rank_records(const Workspace& ws, d3h_tet_record* records, int64_t n_records, cudaStream_t stream) {
  ProfScope ps(K_RANK_RECORDS, stream);
  rank_records_kernel<<<1, 1024, 0, stream>>>(records, n_records, ws.ctr);
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

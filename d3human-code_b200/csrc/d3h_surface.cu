// Stage 3: per-polygon work on the O(surface) records.
//
//   poly_faces_kernel   : watertight faces from triangle_table (gshell_tets.py:322-325), face normals / tangents
//                         splatted to the watertight vertices (auto_normals :9-34, compute_tangents :40-78 with the
//                         "UV indexed by vertex id" quirk of :327), mSDF cut case per polygon (:338-339, :401-404) and
//                         the ordered rank of every polygon inside its faces_aug bucket (6-way scan with look-back).
//   vertex_frame_kernel : normalise the splatted normals, average + Gram-Schmidt the tangents (:28-29, :69-73).
//   poly_cut_kernel     : boundary vertices on every polygon edge (:342-397), zeroing of unreferenced rows (:423-427),
//                         faces_aug emission in the reference's 6-bucket order (:406-420), final counts.
#include "d3h_internal.cuh"

namespace d3h {

struct UvParams {
  int nuv;      // ceil(sqrt(F)), gshell_tets.py:220 with max_idx = 2F
  float step;   // linspace step, end / (nuv - 1) in fp32
  float end;    // fp32(1 - 1/nuv)
  float pad;    // fp32(0.9 / nuv), :226
};

// torch.linspace(0, 1 - 1/nuv, nuv)[i] as ATen's CPU kernel evaluates it: lower half start + step*i, upper half
// end - step*(n-1-i) with a single rounding (vectorised fmadd); pinned against torch in the oracle tests.
__device__ __forceinline__ float uv_lin(const UvParams& p, int i) {
  return (i < p.nuv / 2) ? __fmul_rn(p.step, (float)i) : __fmaf_rn(-p.step, (float)(p.nuv - 1 - i), p.end);
}
// UV the reference reads for *vertex id* k (uvs_pre[faces], :327 -> :228-233)
__device__ __forceinline__ float2 vertex_uv(const UvParams& p, int k) {
  const int cell = k >> 2, c = k & 3;
  const int ix = cell % p.nuv, iy = cell / p.nuv;
  float u = uv_lin(p, ix), v = uv_lin(p, iy);
  if (c == 1 || c == 2) u = __fadd_rn(u, p.pad);
  if (c == 2 || c == 3) v = __fadd_rn(v, p.pad);
  return make_float2(u, v);
}

__device__ __forceinline__ int pos_in_loop(int code, int edge) {
  int r = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (c_loop_edge[code][k] == edge) r = k;
  return r;
}

// torch.cross on CPU: component = fma(a_i, b_j, -fl(a_j * b_i)) (pinned in the oracle tests)
__device__ __forceinline__ float cross_comp(float ai, float bj, float aj, float bi) {
  return __fmaf_rn(ai, bj, -__fmul_rn(aj, bi));
}

constexpr unsigned long long kPFlagAgg = 1ull << 62, kPFlagInc = 2ull << 62, kPValMask = (1ull << 62) - 1;

__global__ void __launch_bounds__(kPolyThreads)
poly_faces_kernel(const d3h_tet_record* __restrict__ records, DevCounters* __restrict__ ctr,
                  unsigned long long* __restrict__ status, const int32_t* __restrict__ corners,
                  const float4* __restrict__ w_vert, float* __restrict__ w_acc, unsigned* __restrict__ polyinfo,
                  int64_t* __restrict__ faces_wt, int64_t cap_faces_wt, UvParams uvp) {
  constexpr int WARPS = kPolyThreads / 32;
  __shared__ unsigned s_tile;
  __shared__ unsigned s_cnt[6][WARPS];
  __shared__ unsigned s_tile_excl[6];

  const unsigned t1 = ctr->work_tri, t2 = ctr->work_quad;
  const int64_t npoly = (int64_t)t1 + t2;
  const int64_t ntiles = (npoly + kPolyThreads - 1) / kPolyThreads;
  if (threadIdx.x == 0) s_tile = atomicAdd(&ctr->ticket_poly, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  if ((int64_t)tile >= ntiles) return;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)tile * kPolyThreads + threadIdx.x;
  const int64_t n_faces_total = (int64_t)t1 + 2ll * t2;
  const bool cross_quirk = (n_faces_total == 3);  // torch.cross without dim on a (3,3) tensor, gshell_tets.py:19

  int bucket = -1;
  unsigned mcase = 0;
  if (i < npoly) {
    const int4 meta = reinterpret_cast<const int4*>(records + i)[1];
    const int code = meta.x, rank = meta.y;
    const bool quad = __popc((unsigned)code) == 2;
    const int n = quad ? 4 : 3;
    const int64_t p0 = quad ? (3ll * t1 + 4ll * rank) : 3ll * rank;
    int L[4];
    float4 P[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      L[k] = (k < n) ? corners[p0 + k] : 0;
      P[k] = (k < n) ? w_vert[L[k]] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ---- watertight faces + splats ----
    const int ntri = quad ? 2 : 1;
    for (int t = 0; t < ntri; ++t) {
      const int64_t row = quad ? ((int64_t)t1 + 2ll * rank + t) : (int64_t)rank;
      int lp[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) lp[c] = pos_in_loop(code, c_tri_edge[code][3 * t + c]);
      // select without dynamic register indexing
      int vi[3];
      float4 pv[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        vi[c] = lp[c] == 0 ? L[0] : lp[c] == 1 ? L[1] : lp[c] == 2 ? L[2] : L[3];
        pv[c] = lp[c] == 0 ? P[0] : lp[c] == 1 ? P[1] : lp[c] == 2 ? P[2] : P[3];
      }
      if (row < cap_faces_wt) {
        faces_wt[3 * row + 0] = vi[0];
        faces_wt[3 * row + 1] = vi[1];
        faces_wt[3 * row + 2] = vi[2];
      }
      const float ax = __fsub_rn(pv[1].x, pv[0].x), ay = __fsub_rn(pv[1].y, pv[0].y), az = __fsub_rn(pv[1].z, pv[0].z);
      const float bx = __fsub_rn(pv[2].x, pv[0].x), by = __fsub_rn(pv[2].y, pv[0].y), bz = __fsub_rn(pv[2].z, pv[0].z);
      // accumulator row of a vertex: [nx ny nz count | tx ty tz -]; one 16-byte vector atomic per half (sm_90+)
      float nx = 0.f, ny = 0.f, nz = 0.f;
      if (!cross_quirk) {
        nx = cross_comp(ay, bz, az, by);
        ny = cross_comp(az, bx, ax, bz);
        nz = cross_comp(ax, by, ay, bx);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c)
        atomicAdd(reinterpret_cast<float4*>(w_acc + 8ll * vi[c]), make_float4(nx, ny, nz, 1.f));
      // tangent of this face
      const float2 uv0 = vertex_uv(uvp, vi[0]), uv1 = vertex_uv(uvp, vi[1]), uv2 = vertex_uv(uvp, vi[2]);
      const float u1x = __fsub_rn(uv1.x, uv0.x), u1y = __fsub_rn(uv1.y, uv0.y);
      const float u2x = __fsub_rn(uv2.x, uv0.x), u2y = __fsub_rn(uv2.y, uv0.y);
      float den = __fsub_rn(__fmul_rn(u1x, u2y), __fmul_rn(u1y, u2x));
      den = (den > 0.f) ? fmaxf(den, 1e-6f) : fminf(den, -1e-6f);
      const float tx = __fdiv_rn(__fsub_rn(__fmul_rn(ax, u2y), __fmul_rn(bx, u1y)), den);
      const float ty = __fdiv_rn(__fsub_rn(__fmul_rn(ay, u2y), __fmul_rn(by, u1y)), den);
      const float tz = __fdiv_rn(__fsub_rn(__fmul_rn(az, u2y), __fmul_rn(bz, u1y)), den);
#pragma unroll
      for (int c = 0; c < 3; ++c)
        atomicAdd(reinterpret_cast<float4*>(w_acc + 8ll * vi[c]) + 1, make_float4(tx, ty, tz, 0.f));
    }
    // ---- mSDF cut case (sign of the interpolated mSDF at the polygon corners) ----
    const unsigned mo0 = P[0].w > 0.f, mo1 = P[1].w > 0.f, mo2 = P[2].w > 0.f, mo3 = P[3].w > 0.f;
    int ncut;
    if (quad) {
      mcase = (mo0 << 3) | (mo1 << 2) | (mo2 << 1) | mo3;
      ncut = c_num_cut_quad[mcase];
      bucket = ncut ? (1 + ncut) : -1;  // buckets 2..5
    } else {
      mcase = (mo0 << 2) | (mo1 << 1) | mo2;
      ncut = c_num_cut_tri[mcase];
      bucket = ncut ? (ncut - 1) : -1;  // buckets 0..1
    }
  }

  // ---- ordered rank of the polygon inside its bucket ----
  unsigned my_ballot = 0;
#pragma unroll
  for (int b = 0; b < 6; ++b) {
    const unsigned m = __ballot_sync(0xffffffffu, bucket == b);
    if (bucket == b) my_ballot = m;
    if (lane == 0) s_cnt[b][warp] = __popc(m);
  }
  __syncthreads();
  if (warp == 0) {
    // exclusive offsets over the warps + tile totals, lanes 0..5 = buckets
    unsigned run = 0;
    if (lane < 6) {
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        const unsigned c = s_cnt[lane][w];
        s_cnt[lane][w] = run;
        run += c;
      }
    }
    // three status words per tile, each packing two 31-bit bucket counters; 32 predecessors per look-back step
#pragma unroll 1
    for (int wd = 0; wd < 3; ++wd) {
      const unsigned lo_cnt = __shfl_sync(0xffffffffu, run, 2 * wd), hi_cnt = __shfl_sync(0xffffffffu, run, 2 * wd + 1);
      const unsigned long long agg = (unsigned long long)lo_cnt | ((unsigned long long)hi_cnt << 31);
      unsigned long long excl = 0ull;
      unsigned long long* my = status + (int64_t)tile * 3 + wd;
      if (tile == 0) {
        if (lane == 0) st_relaxed_u64(my, kPFlagInc | agg);
      } else {
        if (lane == 0) st_relaxed_u64(my, kPFlagAgg | agg);
        int64_t look = (int64_t)tile - 1;
        while (true) {
          const int64_t idx = look - lane;
          unsigned long long w = kPFlagInc;
          if (idx >= 0) {
            do { w = ld_relaxed_u64(status + idx * 3 + wd); } while ((w >> 62) == 0ull);
          }
          const unsigned inc_mask = __ballot_sync(0xffffffffu, (w >> 62) == 2ull);
          const int first = inc_mask ? (__ffs(inc_mask) - 1) : 32;
          unsigned long long contrib = ((int)lane <= first) ? (w & kPValMask) : 0ull;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
          excl += contrib;
          if (inc_mask) break;
          look -= 32;
        }
        if (lane == 0) st_relaxed_u64(my, kPFlagInc | (excl + agg));
      }
      if (lane == 0) {
        const unsigned e_lo = (unsigned)(excl & 0x7fffffffull), e_hi = (unsigned)(excl >> 31);
        s_tile_excl[2 * wd] = e_lo;
        s_tile_excl[2 * wd + 1] = e_hi;
        if ((int64_t)tile == ntiles - 1) {
          ctr->bucket[2 * wd] = e_lo + lo_cnt;
          ctr->bucket[2 * wd + 1] = e_hi + hi_cnt;
        }
      }
    }
  }
  __syncthreads();
  if (i < npoly) {
    unsigned rank_in_bucket = 0;
    if (bucket >= 0) rank_in_bucket = s_tile_excl[bucket] + s_cnt[bucket][warp] + __popc(my_ballot & lanemask_lt());
    polyinfo[i] = (rank_in_bucket << 4) | mcase;
  }
  // the (3,3) torch.cross quirk: the three face normals are crossed along the *face* axis
  if (cross_quirk && i == 0) {
    float a[3][3], b[3][3];
    int fv[3][3];
    for (int64_t q = 0; q < npoly; ++q) {
      const int4 meta = reinterpret_cast<const int4*>(records + q)[1];
      const int code = meta.x, rank = meta.y;
      const bool quad = __popc((unsigned)code) == 2;
      const int64_t p0 = quad ? (3ll * t1 + 4ll * rank) : 3ll * rank;
      for (int t = 0; t < (quad ? 2 : 1); ++t) {
        const int64_t row = quad ? ((int64_t)t1 + 2ll * rank + t) : (int64_t)rank;
        float4 pv[3];
        for (int c = 0; c < 3; ++c) {
          fv[row][c] = corners[p0 + pos_in_loop(code, c_tri_edge[code][3 * t + c])];
          pv[c] = w_vert[fv[row][c]];
        }
        a[row][0] = __fsub_rn(pv[1].x, pv[0].x); a[row][1] = __fsub_rn(pv[1].y, pv[0].y); a[row][2] = __fsub_rn(pv[1].z, pv[0].z);
        b[row][0] = __fsub_rn(pv[2].x, pv[0].x); b[row][1] = __fsub_rn(pv[2].y, pv[0].y); b[row][2] = __fsub_rn(pv[2].z, pv[0].z);
      }
    }
    for (int c = 0; c < 3; ++c) {  // column c: vectors (a[0][c], a[1][c], a[2][c]) x (b[0][c], b[1][c], b[2][c])
      float fn[3];
      fn[0] = cross_comp(a[1][c], b[2][c], a[2][c], b[1][c]);
      fn[1] = cross_comp(a[2][c], b[0][c], a[0][c], b[2][c]);
      fn[2] = cross_comp(a[0][c], b[1][c], a[1][c], b[0][c]);
      for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 3; ++k) atomicAdd(w_acc + 8ll * fv[r][k] + c, fn[r]);  // counts were added above
    }
  }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 safe_normalize3(float x, float y, float z) {  // render/util.py:25-29
  const float d = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  const float len = __fsqrt_rn(fmaxf(d, 1e-20f));
  return make_float3(__fdiv_rn(x, len), __fdiv_rn(y, len), __fdiv_rn(z, len));
}

__global__ void __launch_bounds__(256)
vertex_frame_kernel(const DevCounters* __restrict__ ctr, const float* __restrict__ w_acc, float4* __restrict__ w_tng,
                    float* __restrict__ v_tng_wt, int64_t cap_verts, float* __restrict__ v_tng_aug,
                    int64_t cap_verts_aug) {
  const int64_t nv = ctr->n_verts;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += stride) {
    const float4 a0 = reinterpret_cast<const float4*>(w_acc + 8 * v)[0];
    const float4 a1 = reinterpret_cast<const float4*>(w_acc + 8 * v)[1];
    // auto_normals tail, gshell_tets.py:28-29
    float nx = a0.x, ny = a0.y, nz = a0.z;
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz));
    if (!(d > 1e-20f)) { nx = 0.f; ny = 0.f; nz = 1.f; }
    const float3 n = safe_normalize3(nx, ny, nz);
    // compute_tangents tail, gshell_tets.py:69-73
    const float cnt = a0.w;
    float3 t = safe_normalize3(__fdiv_rn(a1.x, cnt), __fdiv_rn(a1.y, cnt), __fdiv_rn(a1.z, cnt));
    const float dp = __fadd_rn(__fadd_rn(__fmul_rn(t.x, n.x), __fmul_rn(t.y, n.y)), __fmul_rn(t.z, n.z));
    t = safe_normalize3(__fsub_rn(t.x, __fmul_rn(dp, n.x)), __fsub_rn(t.y, __fmul_rn(dp, n.y)),
                        __fsub_rn(t.z, __fmul_rn(dp, n.z)));
    w_tng[v] = make_float4(t.x, t.y, t.z, 0.f);
    if (v < cap_verts) { v_tng_wt[3 * v] = t.x; v_tng_wt[3 * v + 1] = t.y; v_tng_wt[3 * v + 2] = t.z; }
    if (v < cap_verts_aug) { v_tng_aug[3 * v] = t.x; v_tng_aug[3 * v + 1] = t.y; v_tng_aug[3 * v + 2] = t.z; }
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPolyThreads)
poly_cut_kernel(const d3h_tet_record* __restrict__ records, const DevCounters* __restrict__ ctr,
                const int32_t* __restrict__ corners, const float4* __restrict__ w_vert,
                const float4* __restrict__ w_tng, const unsigned* __restrict__ polyinfo,
                float* __restrict__ verts_aug, float* __restrict__ v_tng_aug, float* __restrict__ msdf_aug,
                int64_t cap_verts_aug, int64_t* __restrict__ faces_aug, int64_t cap_faces_aug,
                d3h_counts* __restrict__ counts) {
  const unsigned t1 = ctr->work_tri, t2 = ctr->work_quad;
  const int64_t npoly = (int64_t)t1 + t2;
  const int64_t nv = ctr->n_verts;
  // face-row base of each bucket: buckets hold polygons cut into (1,2 | 1,2,3,4) triangles
  int64_t fbase[7];
  {
    const int ncut_of[6] = {1, 2, 1, 2, 3, 4};
    int64_t run = 0;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      fbase[b] = run;
      run += (int64_t)ctr->bucket[b] * ncut_of[b];
    }
    fbase[6] = run;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    counts->n_valid_tets = ctr->n_valid;
    counts->n_tri_tets = ctr->n_tri;
    counts->n_quad_tets = ctr->n_quad;
    counts->n_corners = 3ll * ctr->n_tri + 4ll * ctr->n_quad;
    counts->n_verts = nv;
    counts->n_faces_aug = fbase[6];
    for (int b = 0; b < 6; ++b) counts->bucket_polys[b] = ctr->bucket[b];
    counts->bad_index = 0;
    counts->reserved[0] = (ctr->n_valid != t1 + t2) ? 1 : 0;  // record buffer overflowed: surface stages skipped
    counts->reserved[1] = 0;
    counts->reserved[2] = 0;
  }
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npoly) return;
  const int4 meta = reinterpret_cast<const int4*>(records + i)[1];
  const int code = meta.x, rank = meta.y;
  const bool quad = __popc((unsigned)code) == 2;
  const int n = quad ? 4 : 3;
  const int64_t p0 = quad ? (3ll * t1 + 4ll * rank) : 3ll * rank;
  int L[4];
  float4 P[4], T[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    L[k] = (k < n) ? corners[p0 + k] : 0;
    P[k] = (k < n) ? w_vert[L[k]] : make_float4(0.f, 0.f, 0.f, 0.f);
    T[k] = (k < n) ? w_tng[L[k]] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const unsigned info = polyinfo[i];
  const unsigned mcase = info & 15u;
  const int64_t brank = info >> 4;
  const int ncut = quad ? c_num_cut_quad[mcase] : c_num_cut_tri[mcase];
  // locals referenced by this polygon's cut triangles
  unsigned used_mask = 0;
  for (int e = 0; e < 3 * ncut; ++e) used_mask |= 1u << (quad ? c_cut_quad[mcase][e] : c_cut_tri[mcase][e]);

  // ---- boundary vertex on every polygon edge k -> k+1 ----
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k >= n) break;
    const int kn = (k + 1 == n) ? 0 : k + 1;
    const float4 pi = P[k], pj = (kn == 0) ? P[0] : (kn == 1) ? P[1] : (kn == 2) ? P[2] : P[3];
    const float4 ti = T[k], tj = (kn == 0) ? T[0] : (kn == 1) ? T[1] : (kn == 2) ? T[2] : T[3];
    float u0, u1, D;
    boundary_weights(pi.w, pj.w, u0, u1, D);
    const int64_t row = nv + p0 + k;
    if (row < cap_verts_aug) {
      const bool used = (used_mask >> (n + k)) & 1u;
      verts_aug[3 * row + 0] = used ? lerp2(pi.x, u0, pj.x, u1) : 0.f;
      verts_aug[3 * row + 1] = used ? lerp2(pi.y, u0, pj.y, u1) : 0.f;
      verts_aug[3 * row + 2] = used ? lerp2(pi.z, u0, pj.z, u1) : 0.f;
      v_tng_aug[3 * row + 0] = lerp2(ti.x, u0, tj.x, u1);
      v_tng_aug[3 * row + 1] = lerp2(ti.y, u0, tj.y, u1);
      v_tng_aug[3 * row + 2] = lerp2(ti.z, u0, tj.z, u1);
      msdf_aug[row] = lerp2(pi.w, u0, pj.w, u1);
    }
  }
  // ---- cut triangles ----
  if (ncut > 0) {
    const int bucket = quad ? (1 + ncut) : (ncut - 1);
    const int64_t row0 = fbase[bucket] + brank * ncut;
    for (int e = 0; e < 3 * ncut; ++e) {
      const int loc = quad ? c_cut_quad[mcase][e] : c_cut_tri[mcase][e];
      int64_t g;
      if (loc < n) g = (loc == 0) ? L[0] : (loc == 1) ? L[1] : (loc == 2) ? L[2] : L[3];
      else g = nv + p0 + (loc - n);
      const int64_t frow = row0 + e / 3;
      if (frow < cap_faces_aug) faces_aug[3 * frow + (e % 3)] = g;
    }
  }
}

// ------------------------------------------------------------------------------------------------
void launch_surface(const d3h_forward_args& a, const Workspace& ws, const d3h_tet_record* records,
                    cudaStream_t stream) {
  UvParams uvp;
  {
    // map_uv(face_gidx_pre, num_tets*2): N = int(ceil(sqrt((max_idx+1)//2))), gshell_tets.py:220,319
    const int64_t half = (2 * a.n_tets + 1) / 2;
    uvp.nuv = (int)ceil(sqrt((double)half));
    if (uvp.nuv < 1) uvp.nuv = 1;
    uvp.end = (float)(1.0 - (1.0 / (double)uvp.nuv));
    uvp.step = (uvp.nuv > 1) ? uvp.end / (float)(uvp.nuv - 1) : 0.f;
    uvp.pad = (float)(0.9 / (double)uvp.nuv);
  }
  const unsigned nblk = (unsigned)(ws.ntiles_poly > 0 ? ws.ntiles_poly : 1);
  if (ws.cap_tets > 0) {
    {
      ProfScope ps(K_POLY_FACES, stream);
      poly_faces_kernel<<<nblk, kPolyThreads, 0, stream>>>(records, ws.ctr, ws.st_poly, a.tape_corners, ws.vert, ws.acc,
                                                           ws.polyinfo, a.faces_wt, a.cap_faces_wt, uvp);
    }
    int64_t vb = (ws.cap_corners + 255) / 256;
    if (vb > 148 * 8) vb = 148 * 8;
    ProfScope ps(K_VERTEX_FRAME, stream);
    vertex_frame_kernel<<<(unsigned)vb, 256, 0, stream>>>(ws.ctr, ws.acc, ws.tng, a.v_tng_wt, a.cap_verts,
                                                          a.v_tng_aug, a.cap_verts_aug);
  }
  ProfScope ps(K_POLY_CUT, stream);
  poly_cut_kernel<<<nblk, kPolyThreads, 0, stream>>>(records, ws.ctr, a.tape_corners, ws.vert, ws.tng, ws.polyinfo,
                                                     a.verts_aug, a.v_tng_aug, a.msdf_aug, a.cap_verts_aug, a.faces_aug,
                                                     a.cap_faces_aug, ws.counts);
}

}  // namespace d3h

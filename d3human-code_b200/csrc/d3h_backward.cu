// Backward of the float pipeline (the reference relies on autograd through gshell_tets.py:291-303, 342-397, 427).
//
//   zero_kernel      : dense (N,3)+(N)+(N) gradients and the per-vertex accumulators, one launch, 16-byte stores
//   boundary_adjoint : one thread per polygon corner p (= boundary vertex V+p): pulls g_verts_aug / g_msdf_aug of that
//                      row back onto the two watertight vertices of its polygon edge and onto their mSDF values
//   crossing_adjoint : one thread per watertight vertex (= crossing edge (a,b), sorted by a): adds its own upstream
//                      rows, then scatters into pos / sdf / msdf of the two grid vertices.  Edges are sorted by `a`,
//                      so lanes that share `a` are adjacent: their a-side contributions are combined with a segmented
//                      warp reduction before one atomic per run; b-side contributions use plain float atomics.
//
// Gradient formulas: SURVEY.md appendix A.5 (verified against the reference's autograd in tests/).
#include "d3h_internal.cuh"

namespace d3h {

__global__ void __launch_bounds__(256) zero_kernel(float4* __restrict__ p0, int64_t n0, float4* __restrict__ p1,
                                                   int64_t n1, float4* __restrict__ p2, int64_t n2,
                                                   float4* __restrict__ p3, int64_t n3, float* t0, int64_t tn0,
                                                   float* t1, int64_t tn1, float* t2, int64_t tn2) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = tid; i < n0; i += stride) p0[i] = z;
  for (int64_t i = tid; i < n1; i += stride) p1[i] = z;
  for (int64_t i = tid; i < n2; i += stride) p2[i] = z;
  for (int64_t i = tid; i < n3; i += stride) p3[i] = z;
  // scalar tails (buffers are only guaranteed 4-byte granular in length)
  if (tid < tn0) t0[tid] = 0.f;
  if (tid < tn1) t1[tid] = 0.f;
  if (tid < tn2) t2[tid] = 0.f;
}

// accumulators per watertight vertex: [0..2] g_vert, [3] g_sg (stop-grad mSDF attribute), [4] g_mv (mSDF through
// the boundary coefficients); stride 8 floats
__global__ void __launch_bounds__(256)
boundary_adjoint_kernel(const int32_t* __restrict__ corners, const float* __restrict__ verts_wt,
                        const float* __restrict__ msdf_wt, int64_t nv, int64_t t1, int64_t t2,
                        const float* __restrict__ g_verts_aug, const float* __restrict__ g_msdf_aug,
                        float* __restrict__ acc) {
  const int64_t ncorn = 3 * t1 + 4 * t2;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ncorn) return;
  int64_t pn;  // next corner of the same polygon
  if (p < 3 * t1) {
    const int64_t b = p - p % 3;
    pn = b + (p % 3 + 1) % 3;
  } else {
    const int64_t q = p - 3 * t1;
    const int64_t b = 3 * t1 + (q - q % 4);
    pn = b + (q % 4 + 1) % 4;
  }
  const int i = corners[p], j = corners[pn];
  const float mi = __ldg(msdf_wt + i), mj = __ldg(msdf_wt + j);
  float u0, u1, D;
  const bool nz = boundary_weights(mi, mj, u0, u1, D);
  const int64_t row = nv + p;
  // the row is zeroed in forward unless its polygon's cut references it (gshell_tets.py:423-427):
  // that is exactly when the mSDF sign changes across the edge
  const bool used = (mi > 0.f) != (mj > 0.f);
  float gx = 0.f, gy = 0.f, gz = 0.f, gm = 0.f;
  if (g_verts_aug != nullptr && used) {
    gx = __ldg(g_verts_aug + 3 * row);
    gy = __ldg(g_verts_aug + 3 * row + 1);
    gz = __ldg(g_verts_aug + 3 * row + 2);
  }
  if (g_msdf_aug != nullptr) gm = __ldg(g_msdf_aug + row);
  if (gx == 0.f && gy == 0.f && gz == 0.f && gm == 0.f) return;
  float* ai = acc + 8ll * i;
  float* aj = acc + 8ll * j;
  if (u0 != 0.f) {
    atomicAdd(ai + 0, gx * u0); atomicAdd(ai + 1, gy * u0); atomicAdd(ai + 2, gz * u0);
    atomicAdd(ai + 3, gm * u0);
  }
  if (u1 != 0.f) {
    atomicAdd(aj + 0, gx * u1); atomicAdd(aj + 1, gy * u1); atomicAdd(aj + 2, gz * u1);
    atomicAdd(aj + 3, gm * u1);
  }
  if (nz) {
    const float gu0 = gx * __ldg(verts_wt + 3ll * i) + gy * __ldg(verts_wt + 3ll * i + 1) + gz * __ldg(verts_wt + 3ll * i + 2);
    const float gu1 = gx * __ldg(verts_wt + 3ll * j) + gy * __ldg(verts_wt + 3ll * j + 1) + gz * __ldg(verts_wt + 3ll * j + 2);
    const float inv = 1.f / D;
    const float gD = -(gu0 * u0 + gu1 * u1) * inv;
    atomicAdd(ai + 4, gu1 * inv + gD);
    atomicAdd(aj + 4, -(gu0 * inv + gD));
  }
}

// segmented (by key) inclusive suffix-sum inside a warp; lanes with equal key must be contiguous.
// returns the run total in the first lane of every run.
__device__ __forceinline__ float seg_reduce_to_head(float v, int key, unsigned active) {
  const unsigned lane = lane_id();
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float nv = __shfl_down_sync(active, v, o);
    const int nk = __shfl_down_sync(active, key, o);
    if (lane + o < 32 && ((active >> (lane + o)) & 1u) && nk == key) v += nv;
  }
  return v;
}

__global__ void __launch_bounds__(256)
crossing_adjoint_kernel(const int32_t* __restrict__ edges, const float* __restrict__ pos,
                        const float* __restrict__ sdf, const float* __restrict__ msdf, int msdf_negate, int64_t nv,
                        const float* __restrict__ msdf_wt, const float* __restrict__ g_verts_aug,
                        const float* __restrict__ g_msdf_aug, const float* __restrict__ g_verts_wt,
                        const float* __restrict__ g_msdf_wt, const float* __restrict__ acc, float* __restrict__ g_pos,
                        float* __restrict__ g_sdf, float* __restrict__ g_msdf) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = v < nv;
  const unsigned active = __ballot_sync(0xffffffffu, live);
  if (!live) return;
  const int a = edges[2 * v], b = edges[2 * v + 1];
  const float4 c0 = reinterpret_cast<const float4*>(acc + 8 * v)[0];
  const float c1 = acc[8 * v + 4];
  float gx = c0.x, gy = c0.y, gz = c0.z, gsg = c0.w, gmv = c1;
  const bool used = __ldg(msdf_wt + v) > 0.f;  // verts_aug[v] was zeroed in forward otherwise
  if (g_verts_aug != nullptr && used) {
    gx += __ldg(g_verts_aug + 3 * v); gy += __ldg(g_verts_aug + 3 * v + 1); gz += __ldg(g_verts_aug + 3 * v + 2);
  }
  if (g_verts_wt != nullptr) {
    gx += __ldg(g_verts_wt + 3 * v); gy += __ldg(g_verts_wt + 3 * v + 1); gz += __ldg(g_verts_wt + 3 * v + 2);
  }
  if (g_msdf_aug != nullptr) gsg += __ldg(g_msdf_aug + v);
  if (g_msdf_wt != nullptr) gsg += __ldg(g_msdf_wt + v);

  float w0, w1, dd;
  crossing_weights(__ldg(sdf + a), __ldg(sdf + b), w0, w1, dd);
  float ma = __ldg(msdf + a), mb = __ldg(msdf + b);
  if (msdf_negate) { ma = -ma; mb = -mb; }
  const float pax = __ldg(pos + 3ll * a), pay = __ldg(pos + 3ll * a + 1), paz = __ldg(pos + 3ll * a + 2);
  const float pbx = __ldg(pos + 3ll * b), pby = __ldg(pos + 3ll * b + 1), pbz = __ldg(pos + 3ll * b + 2);
  const float gw0 = gx * pax + gy * pay + gz * paz + gmv * ma;
  const float gw1 = gx * pbx + gy * pby + gz * pbz + gmv * mb;
  const float inv = 1.f / dd;
  const float gdd = -(gw0 * w0 + gw1 * w1) * inv;
  const float gm_in = gmv + gsg;

  // a side: runs of equal `a` are contiguous (edges sorted lexicographically) -> one atomic per run and warp
  float sax = seg_reduce_to_head(gx * w0, a, active);
  float say = seg_reduce_to_head(gy * w0, a, active);
  float saz = seg_reduce_to_head(gz * w0, a, active);
  float sas = seg_reduce_to_head(gw1 * inv + gdd, a, active);
  float sam = seg_reduce_to_head(gm_in * w0, a, active);
  const unsigned lane = lane_id();
  const int a_prev = __shfl_up_sync(active, a, 1);
  const bool head = (lane == 0) || !((active >> (lane - 1)) & 1u) || (a_prev != a);
  if (head) {
    atomicAdd(g_pos + 3ll * a, sax); atomicAdd(g_pos + 3ll * a + 1, say); atomicAdd(g_pos + 3ll * a + 2, saz);
    atomicAdd(g_sdf + a, sas);
    if (g_msdf != nullptr) atomicAdd(g_msdf + a, sam);
  }
  // b side
  atomicAdd(g_pos + 3ll * b, gx * w1); atomicAdd(g_pos + 3ll * b + 1, gy * w1); atomicAdd(g_pos + 3ll * b + 2, gz * w1);
  atomicAdd(g_sdf + b, -(gw0 * inv + gdd));
  if (g_msdf != nullptr) atomicAdd(g_msdf + b, gm_in * w1);
}

void launch_backward(const d3h_backward_args& a, cudaStream_t stream) {
  const int64_t n = a.n_grid, nv = a.n_verts;
  float* acc = reinterpret_cast<float*>(a.workspace);
  // split every buffer in a 16-byte-aligned body and a scalar tail
  auto body = [](float* p, int64_t len, float4*& b4, int64_t& n4, float*& tail, int64_t& ntail) {
    const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
    n4 = aligned ? len / 4 : 0;
    b4 = reinterpret_cast<float4*>(p);
    tail = p + 4 * n4;
    ntail = len - 4 * n4;
  };
  float4 *b0, *b1, *b2;
  int64_t n0, n1, n2, tn0, tn1, tn2;
  float *t0, *t1, *t2;
  body(a.g_pos, 3 * n, b0, n0, t0, tn0);
  body(a.g_sdf, n, b1, n1, t1, tn1);
  body(a.g_msdf, a.g_msdf ? n : 0, b2, n2, t2, tn2);
  int64_t work = n0 + n1 + n2 + 2 * nv;
  int64_t blocks = (work / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  // unaligned buffers (never produced by torch) fall back to the scalar tail path in a loop-free kernel: reject instead
  {
    ProfScope ps(K_ZERO, stream);
    zero_kernel<<<(unsigned)blocks, 256, 0, stream>>>(b0, n0, b1, n1, b2, n2, reinterpret_cast<float4*>(acc), 2 * nv,
                                                      t0, tn0, t1, tn1, t2, tn2);
  }
  if (nv <= 0) return;
  const int64_t ncorn = 3 * a.n_tri_tets + 4 * a.n_quad_tets;
  if (ncorn > 0 && (a.g_verts_aug != nullptr || a.g_msdf_aug != nullptr)) {
    ProfScope ps(K_BOUNDARY_ADJ, stream);
    boundary_adjoint_kernel<<<(unsigned)((ncorn + 255) / 256), 256, 0, stream>>>(
        a.tape_corners, a.verts_wt, a.msdf_wt, nv, a.n_tri_tets, a.n_quad_tets, a.g_verts_aug, a.g_msdf_aug, acc);
  }
  ProfScope ps(K_CROSSING_ADJ, stream);
  crossing_adjoint_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, stream>>>(
      a.tape_edges, a.pos, a.sdf, a.msdf, a.msdf_negate, nv, a.msdf_wt, a.g_verts_aug, a.g_msdf_aug, a.g_verts_wt,
      a.g_msdf_wt, acc, a.g_pos, a.g_sdf, a.g_msdf);
}

}  // namespace d3h

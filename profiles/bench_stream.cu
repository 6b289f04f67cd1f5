// Micro-benchmark of the O(F) classification stream (profiles/ only; not part of the library).
// Builds a res^3 Kuhn grid + sphere occupancy bitmap on the device and times several ways of streaming the 16 B/tet
// index array through the sign look-up, to find what limits classify_kernel.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo profiles/bench_stream.cu -o profiles/build/bench_stream
//   profiles/build/bench_stream [res=128] [reps=20]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void make_grid(int res, int4* tets, unsigned* bits, int64_t n_grid) {
  const int n = res + 1;
  const int64_t ncubes = (int64_t)res * res * res;
  const int perm[6][3] = {{1, 2, 4}, {1, 4, 2}, {2, 1, 4}, {2, 4, 1}, {4, 1, 2}, {4, 2, 1}};
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncubes; c += (int64_t)gridDim.x * blockDim.x) {
    const int cx = c % res, cy = (c / res) % res, cz = c / ((int64_t)res * res);
    const int64_t base = (int64_t)cz * n * n + (int64_t)cy * n + cx;
    auto corner = [&](int b) { return (int)(base + ((b >> 2) & 1) * n * n + ((b >> 1) & 1) * n + (b & 1)); };
    for (int p = 0; p < 6; ++p)
      tets[c * 6 + p] = make_int4(corner(0), corner(perm[p][0]), corner(perm[p][0] | perm[p][1]), corner(7));
  }
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < (n_grid + 31) / 32; w += (int64_t)gridDim.x * blockDim.x) {
    unsigned word = 0;
    for (int b = 0; b < 32; ++b) {
      const int64_t v = w * 32 + b;
      if (v >= n_grid) break;
      const float x = -1.f + 2.f * (v % n) / res, y = -1.f + 2.f * ((v / n) % n) / res, z = -1.f + 2.f * (v / ((int64_t)n * n)) / res;
      if (0.6f - sqrtf(x * x + y * y + z * z) > 0.f) word |= 1u << b;
    }
    bits[w] = word;
  }
}

__device__ __forceinline__ int4 ld_hint(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ld_noalloc(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ld_plain(const int4* p) { return __ldg(p); }
__device__ __forceinline__ unsigned occ_of(const unsigned* __restrict__ bits, int v) { return (__ldg(bits + (v >> 5)) >> (v & 31)) & 1u; }

// ---- variant 0: read-only (upper bound of the stream) ----
__global__ void __launch_bounds__(256) k_readonly(const int4* __restrict__ tets, int64_t n, unsigned* out) {
  unsigned acc = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
  for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * 8 + (threadIdx.x & 31); i0 < n; i0 += stride) {
    int4 t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = (i0 + j * 32 < n) ? ld_noalloc(tets + i0 + j * 32) : make_int4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc ^= t[j].x ^ t[j].y ^ t[j].z ^ t[j].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}

// ---- variants 1-4: the classify loop, parameterised by loader / persistence / items ----
template <int ITEMS, int LOADER, bool PERSISTENT>
__global__ void __launch_bounds__(256) k_classify(const int4* __restrict__ tets, int64_t n, const unsigned* __restrict__ bits,
                                                  unsigned* __restrict__ m1, unsigned* __restrict__ m2, unsigned* __restrict__ tile_cnt) {
  const unsigned lane = threadIdx.x & 31;
  const int64_t nchunks = (n + 32 * ITEMS - 1) / (32 * ITEMS);
  const int64_t warps_total = PERSISTENT ? (int64_t)gridDim.x * 8 : nchunks;
  for (int64_t chunk = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5; chunk < nchunks; chunk += warps_total) {
    const int64_t base = chunk * 32 * ITEMS;
    int4 t[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int64_t idx = base + j * 32 + lane;
      t[j] = (idx < n) ? (LOADER == 0 ? ld_hint(tets + idx) : LOADER == 1 ? ld_noalloc(tets + idx) : ld_plain(tets + idx)) : make_int4(0, 0, 0, 0);
    }
    unsigned w1 = 0, w2 = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int64_t idx = base + j * 32 + lane;
      const unsigned c = occ_of(bits, t[j].x) + occ_of(bits, t[j].y) + occ_of(bits, t[j].z) + occ_of(bits, t[j].w);
      const bool valid = (c != 0u) && (c != 4u) && (idx < n);
      const unsigned b1 = __ballot_sync(0xffffffffu, valid && (c != 2u));
      const unsigned b2 = __ballot_sync(0xffffffffu, valid && (c == 2u));
      if (lane == (unsigned)(j & 31)) { w1 = b1; w2 = b2; }
    }
    if (lane < (unsigned)ITEMS) { m1[chunk * ITEMS + lane] = w1; m2[chunk * ITEMS + lane] = w2; }
    unsigned cnt = __popc(w1) | (__popc(w2) << 16);
#pragma unroll
    for (int o = 1; o < ITEMS; o <<= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0 && cnt != 0u) atomicAdd(tile_cnt + (base >> 13), cnt);
    if (!PERSISTENT) break;
  }
}

// ---- variant 5: TMA bulk copies into a shared-memory ring, mbarrier producer/consumer ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int STAGES, int STAGE_TETS>
__global__ void __launch_bounds__(256 + 32) k_classify_tma(const int4* __restrict__ tets, int64_t n, const unsigned* __restrict__ bits,
                                                           unsigned* __restrict__ m1, unsigned* __restrict__ m2, unsigned* __restrict__ tile_cnt) {
  extern __shared__ __align__(128) unsigned char smem[];
  int4* ring = reinterpret_cast<int4*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_TETS * 16);
  uint64_t* empty = full + STAGES;
  const int64_t nstages_total = (n + STAGE_TETS - 1) / STAGE_TETS;  // global stage index space
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (warp == 8) {  // producer warp
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t g = blockIdx.x; g < nstages_total; g += gridDim.x) {
        mbar_wait(empty + s, ph ^ 1);
        const int64_t t0 = g * STAGE_TETS;
        const int64_t cnt = (n - t0 < STAGE_TETS) ? (n - t0) : STAGE_TETS;
        mbar_expect_tx(full + s, (uint32_t)(cnt * 16));
        bulk_g2s(ring + (size_t)s * STAGE_TETS, tets + t0, (uint32_t)(cnt * 16), full + s);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
    return;
  }
  // consumers: 8 warps, each takes STAGE_TETS/8 tets of every stage, 32 at a time
  constexpr int PER_WARP = STAGE_TETS / 8;  // tets per warp per stage
  constexpr int ITEMS = PER_WARP / 32;
  int s = 0;
  uint32_t ph = 0;
  for (int64_t g = blockIdx.x; g < nstages_total; g += gridDim.x) {
    mbar_wait(full + s, ph);
    const int4* src = ring + (size_t)s * STAGE_TETS + warp * PER_WARP;
    const int64_t base = g * STAGE_TETS + warp * PER_WARP;
    int4 t[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) t[j] = src[j * 32 + lane];
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);  // data is in registers: hand the slot back before the look-ups
    unsigned w1 = 0, w2 = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int64_t idx = base + j * 32 + lane;
      const bool in = idx < n;
      const unsigned c = in ? (occ_of(bits, t[j].x) + occ_of(bits, t[j].y) + occ_of(bits, t[j].z) + occ_of(bits, t[j].w)) : 0u;
      const bool valid = (c != 0u) && (c != 4u);
      const unsigned b1 = __ballot_sync(0xffffffffu, valid && (c != 2u));
      const unsigned b2 = __ballot_sync(0xffffffffu, valid && (c == 2u));
      if (lane == (unsigned)j) { w1 = b1; w2 = b2; }
    }
    if (lane < (unsigned)ITEMS && base + lane * 32 < n) { m1[(base >> 5) + lane] = w1; m2[(base >> 5) + lane] = w2; }
    unsigned cnt = __popc(w1) | (__popc(w2) << 16);
#pragma unroll
    for (int o = 1; o < ITEMS; o <<= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0 && cnt != 0u) atomicAdd(tile_cnt + (base >> 13), cnt);
    if (++s == STAGES) { s = 0; ph ^= 1; }
  }
}

struct Result { const char* name; float us; unsigned long long valid; };

int main(int argc, char** argv) {
  const int res = argc > 1 ? atoi(argv[1]) : 128;
  const int reps = argc > 2 ? atoi(argv[2]) : 20;
  const int64_t n = 6ll * res * res * res, n_grid = (int64_t)(res + 1) * (res + 1) * (res + 1);
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int4* tets; unsigned *bits, *m1, *m2, *tile_cnt, *out;
  CK(cudaMalloc(&tets, n * 16)); CK(cudaMalloc(&bits, (n_grid / 32 + 2) * 4));
  CK(cudaMalloc(&m1, (n / 32 + 1024) * 4)); CK(cudaMalloc(&m2, (n / 32 + 1024) * 4));
  const int64_t ntiles = n / 8192 + 2;
  CK(cudaMalloc(&tile_cnt, ntiles * 4)); CK(cudaMalloc(&out, 64));
  make_grid<<<sms * 8, 256>>>(res, tets, bits, n_grid);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<Result> results;
  auto run = [&](const char* name, auto launch) {
    std::vector<float> ts;
    unsigned long long valid = 0;
    for (int r = 0; r < reps + 3; ++r) {
      CK(cudaMemsetAsync(tile_cnt, 0, ntiles * 4));
      CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (r >= 3) ts.push_back(ms * 1e3f);
    }
    std::vector<unsigned> h(ntiles);
    CK(cudaMemcpy(h.data(), tile_cnt, ntiles * 4, cudaMemcpyDeviceToHost));
    for (auto c : h) valid += (c & 0xffff) + (c >> 16);
    std::sort(ts.begin(), ts.end());
    results.push_back({name, ts[ts.size() / 2], valid});
    printf("%-44s median %7.2f us  min %7.2f us  %7.1f GB/s  valid=%llu\n", name, ts[ts.size() / 2], ts[0], n * 16.0 / ts[ts.size() / 2] / 1e3, valid);
  };
  const int64_t nchunks8 = (n + 255) / 256;
  const unsigned nonpers8 = (unsigned)((nchunks8 + 7) / 8);
  run("0 read-only, persistent x8", [&] { k_readonly<<<sms * 8, 256>>>(tets, n, out); });
  run("1 classify hint persistent items8", [&] { k_classify<8, 0, true><<<sms * 4, 256>>>(tets, n, bits, m1, m2, tile_cnt); });
  run("1b classify hint persistent items8 x8ctas", [&] { k_classify<8, 0, true><<<sms * 8, 256>>>(tets, n, bits, m1, m2, tile_cnt); });
  run("2 classify hint one-shot items8", [&] { k_classify<8, 0, false><<<nonpers8, 256>>>(tets, n, bits, m1, m2, tile_cnt); });
  run("3 classify plain ldg persistent items8", [&] { k_classify<8, 2, true><<<sms * 4, 256>>>(tets, n, bits, m1, m2, tile_cnt); });
  run("3b classify plain ldg one-shot items8", [&] { k_classify<8, 2, false><<<nonpers8, 256>>>(tets, n, bits, m1, m2, tile_cnt); });
  run("4 classify no_allocate persistent items8", [&] { k_classify<8, 1, true><<<sms * 4, 256>>>(tets, n, bits, m1, m2, tile_cnt); });
  run("4b classify no_allocate one-shot items8", [&] { k_classify<8, 1, false><<<nonpers8, 256>>>(tets, n, bits, m1, m2, tile_cnt); });
  run("4c classify no_allocate one-shot items4", [&] { k_classify<4, 1, false><<<(unsigned)(((n + 127) / 128 + 7) / 8), 256>>>(tets, n, bits, m1, m2, tile_cnt); });
  run("4d classify no_allocate persistent items16", [&] { k_classify<16, 1, true><<<sms * 2, 256>>>(tets, n, bits, m1, m2, tile_cnt); });
  {
    constexpr int ST = 4, STT = 2048;  // 4 stages x 32 KB
    const size_t smem = (size_t)ST * STT * 16 + 2 * ST * 8;
    CK(cudaFuncSetAttribute(k_classify_tma<ST, STT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    run("5 TMA ring 4x32KB, 1 CTA/SM", [&] { k_classify_tma<ST, STT><<<sms, 288, smem>>>(tets, n, bits, m1, m2, tile_cnt); });
  }
  {
    constexpr int ST = 3, STT = 2048;  // 3 stages x 32 KB, 2 CTAs/SM
    const size_t smem = (size_t)ST * STT * 16 + 2 * ST * 8;
    CK(cudaFuncSetAttribute(k_classify_tma<ST, STT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    run("5b TMA ring 3x32KB, 2 CTA/SM", [&] { k_classify_tma<ST, STT><<<sms * 2, 288, smem>>>(tets, n, bits, m1, m2, tile_cnt); });
  }
  {
    constexpr int ST = 6, STT = 1024;  // 6 stages x 16 KB, 2 CTAs/SM
    const size_t smem = (size_t)ST * STT * 16 + 2 * ST * 8;
    CK(cudaFuncSetAttribute(k_classify_tma<ST, STT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    run("5c TMA ring 6x16KB, 2 CTA/SM", [&] { k_classify_tma<ST, STT><<<sms * 2, 288, smem>>>(tets, n, bits, m1, m2, tile_cnt); });
  }
  {
    constexpr int ST = 4, STT = 1024;  // 4 stages x 16 KB, 3 CTAs/SM
    const size_t smem = (size_t)ST * STT * 16 + 2 * ST * 8;
    CK(cudaFuncSetAttribute(k_classify_tma<ST, STT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    run("5d TMA ring 4x16KB, 3 CTA/SM", [&] { k_classify_tma<ST, STT><<<sms * 3, 288, smem>>>(tets, n, bits, m1, m2, tile_cnt); });
  }
  return 0;
}

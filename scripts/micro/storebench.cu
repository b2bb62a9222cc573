// Ceiling probes for write-dominated streaming on B200: how fast can warps push 12.8 KB tiles from shared
// memory to HBM with 1-D bulk stores, with and without a small bulk load per tile (the ad2 forward shape)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o storebench storebench.cu ; run: ./storebench
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "../../vael_b200/csrc/common.cuh"
using namespace vael;

template <int SO, bool LOADS>
__global__ void __launch_bounds__(256) k_store(float* out, const float* in, long long n_sub, int tile_bytes, int in_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int per_warp = SO * tile_bytes + 2 * 4096 + 128;
  uint8_t* wb = smem + (size_t)warp * per_warp;
  uint64_t* full = reinterpret_cast<uint64_t*>(wb + SO * tile_bytes + 2 * 4096);
  if (lane == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); fence_mbar_init(); }
  __syncwarp();
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  // fill the stages once (values do not matter)
  for (int i = lane; i < SO * tile_bytes / 4; i += 32) reinterpret_cast<float*>(wb)[i] = (float)i;
  fence_proxy_async();
  __syncwarp();
  uint32_t phase = 0;
  long long it = 0;
  if (LOADS && lane == 0 && gw < n_sub) { mbar_arrive_expect_tx(&full[0], in_bytes); bulk_g2s(wb + SO * tile_bytes, in + gw * (in_bytes / 4), in_bytes, &full[0]); }
  for (long long sub = gw; sub < n_sub; sub += GW, ++it) {
    const int so = (int)(it % SO);
    if (LOADS) {
      const int s = (int)(it & 1);
      const long long nxt = sub + GW;
      if (lane == 0 && nxt < n_sub) { mbar_arrive_expect_tx(&full[s ^ 1], in_bytes); bulk_g2s(wb + SO * tile_bytes + (s ^ 1) * 4096, in + nxt * (in_bytes / 4), in_bytes, &full[s ^ 1]); }
      mbar_wait(&full[s], (phase >> s) & 1u); phase ^= 1u << s;
    }
    if (lane == 0) {
      bulk_wait_read<SO - 1>();
      bulk_s2g(out + sub * (tile_bytes / 4), wb + so * tile_bytes, tile_bytes);
      bulk_commit();
    }
    __syncwarp();
  }
  if (lane == 0) bulk_wait_all<0>();
}

// the same tiles pushed with plain coalesced 128-bit stores (LDS.128 from the staged tile, STG.128 to global)
template <bool LOADS>
__global__ void __launch_bounds__(256) k_store_stg(float* out, const float* in, long long n_sub, int tile_bytes, int in_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int per_warp = tile_bytes + 2 * 4096 + 128;
  uint8_t* wb = smem + (size_t)warp * per_warp;
  uint64_t* full = reinterpret_cast<uint64_t*>(wb + tile_bytes + 2 * 4096);
  if (lane == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); fence_mbar_init(); }
  __syncwarp();
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  for (int i = lane; i < tile_bytes / 4; i += 32) reinterpret_cast<float*>(wb)[i] = (float)i;
  __syncwarp();
  uint32_t phase = 0;
  long long it = 0;
  if (LOADS && lane == 0 && gw < n_sub) { mbar_arrive_expect_tx(&full[0], in_bytes); bulk_g2s(wb + tile_bytes, in + gw * (in_bytes / 4), in_bytes, &full[0]); }
  const float4* st4 = reinterpret_cast<const float4*>(wb);
  for (long long sub = gw; sub < n_sub; sub += GW, ++it) {
    if (LOADS) {
      const int s = (int)(it & 1);
      const long long nxt = sub + GW;
      if (lane == 0 && nxt < n_sub) { mbar_arrive_expect_tx(&full[s ^ 1], in_bytes); bulk_g2s(wb + tile_bytes + (s ^ 1) * 4096, in + nxt * (in_bytes / 4), in_bytes, &full[s ^ 1]); }
      mbar_wait(&full[s], (phase >> s) & 1u); phase ^= 1u << s;
    }
    float4* o4 = reinterpret_cast<float4*>(out + sub * (tile_bytes / 4));
#pragma unroll 5
    for (int i = lane; i < tile_bytes / 16; i += 32) o4[i] = st4[i];
  }
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
// strided boxes: [32 rows x C cols] of a [B, W] tensor, chunk after chunk along the row (the adw forward shape)
template <int SO>
__global__ void __launch_bounds__(256) k_store2d(const __grid_constant__ CUtensorMap map, long long n_sub, int C, int n_chunks) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int tile_bytes = 32 * C * 4;
  uint8_t* wb = smem + (size_t)warp * (SO * tile_bytes);
  for (int i = lane; i < SO * tile_bytes / 4; i += 32) reinterpret_cast<float*>(wb)[i] = (float)i;
  fence_proxy_async();
  __syncwarp();
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  long long n = 0;
  for (long long sub = gw; sub < n_sub; sub += GW) {
    for (int i = 0; i < n_chunks; ++i, ++n) {
      if (lane == 0) {
        bulk_wait_read<SO - 1>();
        tma_store_2d(&map, wb + (n % SO) * tile_bytes, i * C, (int)(sub * 32));
        bulk_commit();
      }
      __syncwarp();
    }
  }
  if (lane == 0) bulk_wait_all<0>();
}
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make_map(float* base, long long B, int W, int C, int l2promo) {
  void* p = nullptr; cudaDriverEntryPointQueryResult st;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st);
  alignas(64) CUtensorMap m; memset(&m, 0, sizeof(m));
  const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)B}; const cuuint64_t strides[1] = {(cuuint64_t)W * 4};
  const cuuint32_t box[2] = {(cuuint32_t)C, 32}; const cuuint32_t es[2] = {1, 1};
  CUresult r = ((enc_fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
  return m;
}

__global__ void k_fill(float4* out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) out[i] = make_float4(1, 2, 3, 4);
}

// NON-persistent tile writers: grid = n_sub / warps-per-CTA, every warp owns ONE 12.8 KB tile (consecutive warps,
// consecutive tiles), reads the tile's 2.5 KB input, and writes the tile either with 128-bit stores or with one bulk
// store from shared memory
template <bool BULK, bool LOADS>
__global__ void __launch_bounds__(256) k_tile_once(float* out, const float* in, long long n_sub, int tile_bytes, int in_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const long long sub = (long long)blockIdx.x * nwarps + warp;
  if (sub >= n_sub) return;
  float acc = 0.0f;
  if (LOADS) {
    const float4* i4 = reinterpret_cast<const float4*>(in + sub * (in_bytes / 4));
    for (int i = lane; i < in_bytes / 16; i += 32) { const float4 t = __ldg(i4 + i); acc += t.x + t.y + t.z + t.w; }
  }
  if (BULK) {
    float4* st4 = reinterpret_cast<float4*>(smem + (size_t)warp * tile_bytes);
    for (int i = lane; i < tile_bytes / 16; i += 32) st4[i] = make_float4(acc, 1, 2, 3);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) { bulk_s2g(out + sub * (tile_bytes / 4), st4, tile_bytes); bulk_commit(); bulk_wait_all<0>(); }
  } else {
    float4* o4 = reinterpret_cast<float4*>(out + sub * (tile_bytes / 4));
#pragma unroll 5
    for (int i = lane; i < tile_bytes / 16; i += 32) o4[i] = make_float4(acc, 1, 2, 3);
  }
}
// persistent, but every warp owns a CONTIGUOUS range of tiles
template <int SO>
__global__ void __launch_bounds__(256) k_store_ranges(float* out, long long n_sub, int tile_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  uint8_t* wb = smem + (size_t)warp * (SO * tile_bytes);
  for (int i = lane; i < SO * tile_bytes / 4; i += 32) reinterpret_cast<float*>(wb)[i] = (float)i;
  fence_proxy_async();
  __syncwarp();
  const long long gw = (long long)blockIdx.x * nwarps + warp, GW = (long long)gridDim.x * nwarps;
  const long long per = (n_sub + GW - 1) / GW, lo = gw * per, hi = min(n_sub, lo + per);
  long long it = 0;
  for (long long sub = lo; sub < hi; ++sub, ++it) {
    if (lane == 0) {
      bulk_wait_read<SO - 1>();
      bulk_s2g(out + sub * (tile_bytes / 4), wb + (it % SO) * tile_bytes, tile_bytes);
      bulk_commit();
    }
    __syncwarp();
  }
  if (lane == 0) bulk_wait_all<0>();
}

// one block per contiguous 16 KB chunk (how torch's elementwise fill is shaped), no grid-stride loop
__global__ void k_fill_chunks(float4* out, long long n4) {
  const long long base = (long long)blockIdx.x * 1024;   // 1024 float4 = 16 KB per block of 256 threads
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long i = base + j * 256 + threadIdx.x;
    if (i < n4) out[i] = make_float4(1, 2, 3, 4);
  }
}

template <typename F> float timeit(F f) {
  for (int i = 0; i < 3; ++i) f();
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); for (int i = 0; i < 10; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / 10;
}

int main() {
  const int tile = 12800, inb = 2560;
  const long long n_sub = 262144;  // 3.36 GB of output
  float *out, *in;
  cudaMalloc(&out, n_sub * tile); cudaMalloc(&in, n_sub * inb); cudaMemset(in, 0, n_sub * inb);
  { float ms = timeit([&] { k_fill<<<148 * 16, 256>>>((float4*)out, n_sub * tile / 16); }); printf("fill(STG.128)            %7.0f GB/s\n", n_sub * (double)tile / ms / 1e6); }
  { const long long n4 = n_sub * tile / 16; float ms = timeit([&] { k_fill_chunks<<<(unsigned)((n4 + 1023) / 1024), 256>>>((float4*)out, n4); }); printf("fill, one block per 16 KB %7.0f GB/s\n", n_sub * (double)tile / ms / 1e6); }
  for (int g : {1, 2, 4, 8, 32}) { float ms = timeit([&] { k_fill<<<148 * g, 256>>>((float4*)out, n_sub * tile / 16); }); printf("fill grid-stride, %2d CTAs/SM %7.0f GB/s\n", g, n_sub * (double)tile / ms / 1e6); }
  { float ms = timeit([&] { cudaMemsetAsync(out, 0, n_sub * tile); }); printf("cudaMemset                %7.0f GB/s\n", n_sub * (double)tile / ms / 1e6); }
  for (int nw : {1, 2, 4, 8}) {
    const unsigned grid = (unsigned)((n_sub + nw - 1) / nw);
    const int smem = nw * tile;
    cudaFuncSetAttribute(k_tile_once<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_tile_once<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float a = timeit([&] { k_tile_once<false, false><<<grid, nw * 32>>>(out, in, n_sub, tile, inb); });
    float b = timeit([&] { k_tile_once<false, true><<<grid, nw * 32>>>(out, in, n_sub, tile, inb); });
    float c = timeit([&] { k_tile_once<true, false><<<grid, nw * 32, smem>>>(out, in, n_sub, tile, inb); });
    float d = timeit([&] { k_tile_once<true, true><<<grid, nw * 32, smem>>>(out, in, n_sub, tile, inb); });
    printf("one tile per warp, %d warps/CTA: STG %7.0f  STG+loads %7.0f  bulk %7.0f  bulk+loads %7.0f GB/s\n", nw,
           n_sub * (double)tile / a / 1e6, n_sub * (double)(tile + inb) / b / 1e6, n_sub * (double)tile / c / 1e6,
           n_sub * (double)(tile + inb) / d / 1e6);
  }
  for (int nw : {2, 4, 6}) {
    const int smem = nw * 2 * tile;
    cudaFuncSetAttribute(k_store_ranges<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float a = timeit([&] { k_store_ranges<2><<<148, nw * 32, smem>>>(out, n_sub, tile); });
    printf("persistent, contiguous range per warp, %d warps/SM: bulk %7.0f GB/s\n", nw, n_sub * (double)tile / a / 1e6);
  }
  for (int nw = 1; nw <= 6; ++nw) {
    const int smem = nw * (2 * tile + 2 * 4096 + 128);
    cudaFuncSetAttribute(k_store<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_store<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float a = timeit([&] { k_store<2, false><<<148, nw * 32, smem>>>(out, in, n_sub, tile, inb); });
    float b = timeit([&] { k_store<2, true><<<148, nw * 32, smem>>>(out, in, n_sub, tile, inb); });
    printf("warps/SM %d  stores only %7.0f GB/s   stores+loads %7.0f GB/s (out+in bytes)\n", nw, n_sub * (double)tile / a / 1e6, n_sub * (double)(tile + inb) / b / 1e6);
  }
  for (int nw : {2, 3, 4, 5, 6, 8})
    for (int cps : {1, 2}) {
      const int smem = nw * (tile + 2 * 4096 + 128);
      if (smem * cps > 220 * 1024) continue;
      cudaFuncSetAttribute(k_store_stg<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaFuncSetAttribute(k_store_stg<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      float a = timeit([&] { k_store_stg<false><<<148 * cps, nw * 32, smem>>>(out, in, n_sub, tile, inb); });
      float b = timeit([&] { k_store_stg<true><<<148 * cps, nw * 32, smem>>>(out, in, n_sub, tile, inb); });
      printf("STG.128 from staged tiles: CTAs/SM %d x %d warps  stores only %7.0f GB/s   stores+loads %7.0f GB/s\n", cps, nw,
             n_sub * (double)tile / a / 1e6, n_sub * (double)(tile + inb) / b / 1e6);
    }
  for (int cps = 2; cps <= 4; ++cps) {  // several smaller CTAs per SM
    const int nw = 2, smem = nw * (2 * tile + 2 * 4096 + 128);
    float a = timeit([&] { k_store<2, false><<<148 * cps, nw * 32, smem>>>(out, in, n_sub, tile, inb); });
    printf("CTAs/SM %d x 2 warps  stores only %7.0f GB/s\n", cps, n_sub * (double)tile / a / 1e6);
  }
  {  // strided 2-D boxes: W = 1000, box 100 (10 chunks), 200 (5), 250 (4) columns; rows 32
    const long long B = 26843 * 32;  // 3.4 GB
    float* big; cudaMalloc(&big, B * 4000);
    for (int C : {100, 200, 250}) for (int promo : {0, 2}) for (int nw : {2, 3, 4}) {
      const int smem = nw * 2 * 32 * C * 4;
      if (smem > 227 * 1024) continue;
      CUtensorMap m = make_map(big, B, 1000, C, promo);
      cudaFuncSetAttribute(k_store2d<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      float a = timeit([&] { k_store2d<2><<<148, nw * 32, smem>>>(m, B / 32, C, 1000 / C); });
      printf("2-D box %3d cols x 32 rows, l2promo %d, warps/SM %d: %7.0f GB/s\n", C, promo, nw, B * 4000.0 / a / 1e6);
    }
    cudaFree(big);
  }
  cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
  return 0;
}

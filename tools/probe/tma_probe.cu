// tma_probe.cu -- stand-alone check of the TMA usage pattern of k_inter (u8 2-D tensor map, small box).
// usage: tma_probe W H PITCH BOXW BOXH MODE X Y     MODE 0: raw PTX, map as __grid_constant__ parameter
//        1: libcu++ wrappers (cuda::device::experimental), 2: raw PTX, map in global memory + acquire fence
//        3: raw PTX, map in global memory, no fence
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
namespace cde = cuda::device::experimental;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int MAXT = 64 * 32;

__device__ void fetch_raw(const void* tmap, int x, int y, uint8_t* out, int bytes, int fence) {
  __shared__ __align__(128) uint8_t tile[MAXT];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (fence) asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(tmap) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(tile)),
                 "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(&bar))
                 : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(smem_u32(&bar)), "r"(0)
                 : "memory");
  } while (!done);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}
__global__ void k_param(const __grid_constant__ CUtensorMap m, int x, int y, uint8_t* out, int bytes) { fetch_raw(&m, x, y, out, bytes, 0); }
__global__ void k_global(const void* m, int x, int y, uint8_t* out, int bytes, int fence) { fetch_raw(m, x, y, out, bytes, fence); }
__global__ void k_cde(const __grid_constant__ CUtensorMap m, int x, int y, uint8_t* out, int bytes) {
  __shared__ __align__(128) uint8_t tile[MAXT];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ cuda::barrier<cuda::thread_scope_block> bar;
  if (threadIdx.x == 0) {
    init(&bar, blockDim.x);
    cde::fence_proxy_async_shared_cta();
  }
  __syncthreads();
  cuda::barrier<cuda::thread_scope_block>::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(tile, &m, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv) {
  if (argc < 9) return 2;
  const int W = atoi(argv[1]), H = atoi(argv[2]), P = atoi(argv[3]), BW = atoi(argv[4]), BH = atoi(argv[5]), mode = atoi(argv[6]);
  const int x = atoi(argv[7]), y = atoi(argv[8]);
  uint8_t* d;
  cudaMalloc(&d, (size_t)P * H + 256);
  std::vector<uint8_t> h((size_t)P * H);
  for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i % P) * 3 + (i / P) * 7);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  alignas(64) CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, strides[1] = {(cuuint64_t)P};
  const cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH}, es[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int bytes = BW * BH;
  uint8_t* dout;
  cudaMalloc(&dout, bytes);
  std::vector<uint8_t> hout(bytes);
  void* dm;
  cudaMalloc(&dm, 128);
  cudaMemcpy(dm, &m, 128, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xEE, bytes);
  if (mode == 0) k_param<<<1, 32>>>(m, x, y, dout, bytes);
  else if (mode == 1) k_cde<<<1, 32>>>(m, x, y, dout, bytes);
  else k_global<<<1, 32>>>(dm, x, y, dout, bytes, mode == 2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("W %d H %d P %d box %dx%d mode %d at (%d,%d): encode rc=%d, %s", W, H, P, BW, BH, mode, x, y, (int)r, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    cudaMemcpy(hout.data(), dout, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < BH; r2++)
      for (int c = 0; c < BW; c++) {
        const int gx = x + c, gy = y + r2;
        const uint8_t want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[(size_t)gy * P + gx] : 0;
        bad += hout[r2 * BW + c] != want;
      }
    printf("  mismatches=%d", bad);
  }
  printf("\n");
  return e != cudaSuccess;
}

// TEST INFRASTRUCTURE -- not part of the product, never on a product path.
//
// A stand-in for <cuda_runtime.h> that lets the product's sources (csrc/*.cu, *.cc) compile with plain g++ into
// tests/simt/_build/libvp8gpu_simt.so, in which every kernel runs on the CPU under a SIMT emulator: the threads of
// a CTA are fibers, a warp collective (__shfl_sync, __syncwarp, __ballot_sync ...) is a rendez-vous of the 32
// fibers of the warp, CTAs run one after the other in launch order, streams are synchronous.  It exists so that the
// kernels' LOGIC (records in, pixels / tokens / decisions out) and the host orchestration around them can be
// checked in the container that has no GPU; it says nothing about memory ordering, residency or speed, and the
// product library (alfalfa_b200/libvp8gpu.so) neither contains nor loads any of this -- without a CUDA device
// vp8gpu_ctx_create fails.  Only tests/test_simt_*.py build and load the emulated library.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>

#ifndef VP8GPU_SIMT_EMUL
#error "tests/simt/include is only for -DVP8GPU_SIMT_EMUL builds"
#endif

// ---- language keywords ---------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static
// CTAs run one at a time, so a kernel's shared variables can be statics; they live in one section that the
// emulator fills with a poison pattern before every CTA (shared memory is not zero on the device either)
#define __shared__ static __attribute__((section("simt_shared")))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(8))) uint2 { uint32_t x, y; };
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
struct __attribute__((aligned(8))) int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(4))) ushort2 { uint16_t x, y; };
struct __attribute__((aligned(4))) uchar4 { uint8_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

// ---- the emulator (tests/simt/simt_runtime.cc) ---------------------------------------------------------
namespace simt {
struct Warp {
  uint64_t slot[32];
  unsigned arrived, gen, live_mask;
};
struct Fiber {
  void* sp;
  uint3 tid;
  int lane;
  Warp* warp;
  bool done;
};
extern Fiber* cur;
extern uint3 g_block;
extern dim3 g_bdim, g_gdim;
void yield();                      // let the other threads of the CTA run (every spin-wait must call it)
void warp_barrier(unsigned mask);  // rendez-vous of the live lanes named by mask
void cta_barrier();
uint8_t* dyn_smem();
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& thread_body);

template <class F>
struct Launch {
  dim3 grid, block;
  size_t smem;
  F f;
  template <class... A>
  void operator()(A... args) {
    auto tup = std::make_tuple(args...);
    run_grid(grid, block, smem, [&] { std::apply(f, tup); });
  }
};
template <class F>
Launch<F> make_launch(dim3 grid, dim3 block, size_t smem, F f) {
  return Launch<F>{grid, block, smem, f};
}
}  // namespace simt

#define threadIdx (::simt::cur->tid)
#define blockIdx (::simt::g_block)
#define blockDim (::simt::g_bdim)
#define gridDim (::simt::g_gdim)

// ---- warp / CTA collectives ----------------------------------------------------------------------------
static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::warp_barrier(mask); }
static inline void __syncthreads() { simt::cta_barrier(); }
static inline void __nanosleep(unsigned) { simt::yield(); }
static inline void __threadfence() {}

namespace simt {
template <class T>
inline T exchange(unsigned mask, T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  Fiber* f = cur;
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  f->warp->slot[f->lane] = bits;
  warp_barrier(mask);
  const uint64_t r = f->warp->slot[src_lane & 31];
  warp_barrier(mask);
  T out;
  memcpy(&out, &r, sizeof(T));
  return out;
}
}  // namespace simt
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  const int lane = simt::cur->lane;
  return simt::exchange(mask, v, (lane & ~(width - 1)) + (src & (width - 1)));
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
  (void)width;
  return simt::exchange(mask, v, simt::cur->lane ^ x);
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  (void)width;
  const int lane = simt::cur->lane;
  return simt::exchange(mask, v, lane >= (int)delta ? lane - (int)delta : lane);
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  (void)width;
  const int lane = simt::cur->lane;
  return simt::exchange(mask, v, lane + (int)delta < 32 ? lane + (int)delta : lane);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  simt::Fiber* f = simt::cur;
  f->warp->slot[f->lane] = pred ? 1 : 0;
  simt::warp_barrier(mask);
  unsigned r = 0;
  const unsigned members = mask & f->warp->live_mask;
  for (int l = 0; l < 32; l++)
    if (((members >> l) & 1) && f->warp->slot[l]) r |= 1u << l;
  simt::warp_barrier(mask);
  return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
inline int __reduce_add_sync(unsigned mask, int v) {
  simt::Fiber* f = simt::cur;
  f->warp->slot[f->lane] = (uint64_t)(uint32_t)v;
  simt::warp_barrier(mask);
  int r = 0;
  const unsigned members = mask & f->warp->live_mask;
  for (int l = 0; l < 32; l++)
    if ((members >> l) & 1) r += (int)(uint32_t)f->warp->slot[l];
  simt::warp_barrier(mask);
  return r;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return (unsigned)__reduce_add_sync(mask, (int)v); }

// ---- loads, atomics, bit tricks ----------------------------------------------------------------------
template <class T>
inline T __ldg(const T* p) { return *p; }
template <class T>
inline T __ldcg(const T* p) {
  asm volatile("" ::: "memory");
  return *p;
}
template <class T>
inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T>
inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T>
inline T atomicMax(T* p, T v) {
  T old = *p;
  if (v > old) *p = v;
  return old;
}
template <class T>
inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
  s &= 31;
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
  const uint64_t v = (uint64_t)a | ((uint64_t)b << 32);
  uint32_t r = 0;
  for (int k = 0; k < 4; k++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * k)) & 7))) & 0xFF) << (8 * k);
  return r;
}
inline long long clock64() { return 0; }
// IEEE single-precision operations that the compiler must not contract
__attribute__((__noinline__)) inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
__attribute__((__noinline__)) inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
__attribute__((__noinline__)) inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }

// CUDA's min / max overloads in the global namespace
template <class A, class B, class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
inline typename std::common_type<A, B>::type min(A a, B b) {
  typedef typename std::common_type<A, B>::type C;
  return (C)a < (C)b ? (C)a : (C)b;
}
template <class A, class B, class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
inline typename std::common_type<A, B>::type max(A a, B b) {
  typedef typename std::common_type<A, B>::type C;
  return (C)a > (C)b ? (C)a : (C)b;
}

// ---- the runtime API the engine uses: one device, synchronous streams ---------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600 };
struct simt_stream;
struct simt_event;
typedef simt_stream* cudaStream_t;
typedef simt_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaHostAllocDefault = 0, cudaHostAllocMapped = 2, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventBlockingSync = 1,
       cudaEnableDefault = 0 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
typedef void (*cudaHostFn_t)(void*);

extern "C" {
cudaError_t cudaSetDevice(int);
cudaError_t cudaGetDeviceCount(int*);
cudaError_t cudaDeviceGetAttribute(int*, cudaDeviceAttr, int);
cudaError_t cudaDeviceGetStreamPriorityRange(int*, int*);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaGetLastError();
const char* cudaGetErrorString(cudaError_t);
cudaError_t cudaMalloc(void**, size_t);
cudaError_t cudaFree(void*);
cudaError_t cudaHostAlloc(void**, size_t, unsigned);
cudaError_t cudaFreeHost(void*);
cudaError_t cudaHostGetDevicePointer(void**, void*, unsigned);
cudaError_t cudaMemcpy(void*, const void*, size_t, cudaMemcpyKind);
cudaError_t cudaMemcpyAsync(void*, const void*, size_t, cudaMemcpyKind, cudaStream_t = nullptr);
cudaError_t cudaMemset(void*, int, size_t);
cudaError_t cudaMemsetAsync(void*, int, size_t, cudaStream_t = nullptr);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t*, unsigned);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t*, unsigned, int);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0);
cudaError_t cudaLaunchHostFunc(cudaStream_t, cudaHostFn_t, void*);
cudaError_t cudaEventCreate(cudaEvent_t*);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t*, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t);
cudaError_t cudaEventQuery(cudaEvent_t);
cudaError_t cudaEventElapsedTime(float*, cudaEvent_t, cudaEvent_t);
cudaError_t cudaGetDriverEntryPoint(const char*, void**, unsigned long long, cudaDriverEntryPointQueryResult* = nullptr);
}
cudaError_t cudaMemcpy2D(void*, size_t, const void*, size_t, size_t, size_t, cudaMemcpyKind);
cudaError_t cudaMemcpy2DAsync(void*, size_t, const void*, size_t, size_t, size_t, cudaMemcpyKind, cudaStream_t = nullptr);
cudaError_t cudaMemset2D(void*, size_t, int, size_t, size_t);
#define CUDART_CB
template <class T>
inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
template <class T>
inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned f) { return cudaHostAlloc(reinterpret_cast<void**>(p), n, f); }
template <class T>
inline cudaError_t cudaHostGetDevicePointer(T** d, void* h, unsigned f) { return cudaHostGetDevicePointer(reinterpret_cast<void**>(d), h, f); }
template <class K>
inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <class T>
inline cudaError_t cudaMemcpyToSymbol(T& symbol, const void* src, size_t n, size_t offset = 0, cudaMemcpyKind = cudaMemcpyHostToDevice) {
  memcpy(reinterpret_cast<char*>(&symbol) + offset, src, n);
  return cudaSuccess;
}
template <class T>
inline cudaError_t cudaMemcpyFromSymbol(void* dst, const T& symbol, size_t n, size_t offset = 0, cudaMemcpyKind = cudaMemcpyDeviceToHost) {
  memcpy(dst, reinterpret_cast<const char*>(&symbol) + offset, n);
  return cudaSuccess;
}

// ---- what the TMA helpers of kernels.cu become (see the VP8GPU_SIMT_EMUL branches there) -----------
namespace simt {
struct TensorMap2D {  // what the cuTensorMapEncodeTiled stand-in writes into the 128 opaque bytes
  uint64_t magic;
  const uint8_t* base;
  uint64_t dim0, dim1, stride1;
  uint32_t box0, box1;
};
// copies the box at (x, y) (zero fill outside the tensor) and returns the bytes written
uint32_t tma_copy_2d(void* dst, const void* tmap, int x, int y);
}  // namespace simt

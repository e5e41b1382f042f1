// TEST INFRASTRUCTURE -- the SIMT emulator behind tests/simt/include/cuda_runtime.h (read the header first).
//
// Threads of a CTA are fibers on private stacks, switched by hand (x86-64: six callee-saved registers and the
// stack pointer).  Scheduling is round-robin and only happens inside yield() / a barrier, so a run is
// deterministic.  CTAs of a grid run one after the other in launch order (x fastest): every cross-CTA
// dependency of the product's kernels points to a CTA with a smaller ticket, which has finished by then.
// One launch at a time per process (a mutex), streams are synchronous, "device memory" is the heap.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <sys/mman.h>

#include <chrono>
#include <mutex>
#include <vector>

#if !defined(__x86_64__)
#error "the fiber switch is written for x86-64"
#endif

extern "C" void simt_switch(void** save_sp, void* load_sp);
asm(".text\n"
    ".globl simt_switch\n"
    ".type simt_switch,@function\n"
    "simt_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size simt_switch,.-simt_switch\n");

extern "C" char __start_simt_shared[] __attribute__((weak));
extern "C" char __stop_simt_shared[] __attribute__((weak));

namespace simt {

Fiber* cur = nullptr;
uint3 g_block = {0, 0, 0};
dim3 g_bdim, g_gdim;

namespace {
constexpr size_t kStack = 256 * 1024;
std::mutex g_launch_mu;
std::vector<Fiber> g_fibers;
std::vector<Warp> g_warps;
std::vector<char*> g_stacks;
std::vector<uint8_t> g_dyn;
const std::function<void()>* g_body = nullptr;
void* g_main_sp = nullptr;
int g_live = 0;                      // threads of the CTA that have not returned
unsigned g_cta_arrived = 0, g_cta_gen = 0;
unsigned long long g_idle_switches = 0;  // switches since anything made visible progress

[[noreturn]] void stuck(const char* what) {
  fprintf(stderr, "simt: %s (block %u,%u thread %u)\n", what, g_block.x, g_block.y, cur ? cur->tid.x : 0u);
  abort();
}

void switch_to(Fiber* next) {
  Fiber* prev = cur;
  cur = next;
  simt_switch(&prev->sp, next->sp);
}

// SIMT_ORDER=reverse schedules the threads of a CTA from the last to the first.  Between two barriers a thread runs
// undisturbed, so with the default order a value written by lane a and read by lane b > a without a barrier in
// between happens to be there, and with the reverse order the same holds for b < a: a missing __syncwarp fails in
// at least one of the two orders (tests/test_simt_emulation.py runs both).
const bool g_reverse = [] {
  const char* v = getenv("SIMT_ORDER");
  return v && strcmp(v, "reverse") == 0;
}();

Fiber* next_runnable(Fiber* from) {
  const size_t n = g_fibers.size();
  const size_t start = from - g_fibers.data();
  for (size_t k = 1; k <= n; k++) {
    Fiber* f = &g_fibers[g_reverse ? (start + n - k) % n : (start + k) % n];
    if (!f->done) return f;
  }
  return nullptr;
}

void release_warp_if_complete(Warp* w) {
  const unsigned live = (unsigned)__builtin_popcount(w->live_mask);
  if (live && w->arrived >= live) {  // barriers name the whole (live) warp in this code base
    w->arrived = 0;
    w->gen++;
    g_idle_switches = 0;
  }
}

void fiber_main() {
  (*g_body)();
  Fiber* me = cur;
  me->done = true;
  g_live--;
  g_idle_switches = 0;
  me->warp->live_mask &= ~(1u << me->lane);
  if (me->warp->arrived) release_warp_if_complete(me->warp);
  if (g_live && g_cta_arrived >= (unsigned)g_live) {
    g_cta_arrived = 0;
    g_cta_gen++;
  }
  Fiber* next = next_runnable(me);
  void* dummy;
  if (next) {
    cur = next;
    simt_switch(&dummy, next->sp);
  } else {
    simt_switch(&dummy, g_main_sp);
  }
  __builtin_unreachable();
}
}  // namespace

void yield() {
  Fiber* next = next_runnable(cur);
  if (!next || next == cur) {
    if (++g_idle_switches > 50000000ull) stuck("a thread spins and nobody else can run");
    return;
  }
  if (++g_idle_switches > 4000000000ull) stuck("no progress: dead-locked wait");
  switch_to(next);
}

void warp_barrier(unsigned mask) {
  Warp* w = cur->warp;
  const unsigned members = mask & w->live_mask;
  const unsigned need = (unsigned)__builtin_popcount(members);
  if (need <= 1) return;
  const unsigned gen = w->gen;
  if (++w->arrived >= need) {
    w->arrived = 0;
    w->gen++;
    g_idle_switches = 0;
    return;
  }
  while (w->gen == gen) yield();
}

void cta_barrier() {
  const unsigned gen = g_cta_gen;
  if (++g_cta_arrived >= (unsigned)g_live) {
    g_cta_arrived = 0;
    g_cta_gen++;
    g_idle_switches = 0;
    return;
  }
  while (g_cta_gen == gen) yield();
}

uint8_t* dyn_smem() { return g_dyn.data(); }

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  std::lock_guard<std::mutex> lk(g_launch_mu);
  const size_t nthreads = (size_t)block.x * block.y * block.z;
  if (nthreads == 0 || grid.x == 0 || grid.y == 0 || grid.z == 0) return;
  while (g_stacks.size() < nthreads) {
    void* p = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
    if (p == MAP_FAILED) stuck("cannot allocate a fiber stack");
    g_stacks.push_back(static_cast<char*>(p));
  }
  g_dyn.assign(smem + 64, 0);
  g_body = &body;
  g_bdim = block;
  g_gdim = grid;
  const size_t nwarps = (nthreads + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        g_block = uint3{bx, by, bz};
        if (__start_simt_shared && __stop_simt_shared > __start_simt_shared)
          memset(__start_simt_shared, 0xCD, (size_t)(__stop_simt_shared - __start_simt_shared));
        if (smem) memset(g_dyn.data(), 0xCD, smem);
        g_fibers.assign(nthreads, Fiber{});
        g_warps.assign(nwarps, Warp{});
        for (size_t i = 0; i < nthreads; i++) {
          Fiber& f = g_fibers[i];
          f.tid = uint3{(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / ((size_t)block.x * block.y))};
          f.lane = (int)(i & 31);
          f.warp = &g_warps[i >> 5];
          f.warp->live_mask |= 1u << f.lane;
          f.done = false;
          uint64_t* top = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(g_stacks[i]) + kStack) & ~(uintptr_t)15);
          top[-1] = 0;
          top[-2] = reinterpret_cast<uint64_t>(&fiber_main);
          for (int k = 3; k <= 8; k++) top[-k] = 0;
          f.sp = top - 8;
        }
        g_live = (int)nthreads;
        g_cta_arrived = 0;
        g_idle_switches = 0;
        cur = &g_fibers[g_reverse ? nthreads - 1 : 0];
        simt_switch(&g_main_sp, cur->sp);
        cur = nullptr;
      }
  g_body = nullptr;
}

uint32_t tma_copy_2d(void* dst, const void* tmap, int x, int y) {
  const TensorMap2D* t = static_cast<const TensorMap2D*>(tmap);
  if (t->magic != 0x54454e534f524d41ull) stuck("tma_copy_2d: not a tensor map");
  uint8_t* d = static_cast<uint8_t*>(dst);
  for (uint32_t r = 0; r < t->box1; r++)
    for (uint32_t c = 0; c < t->box0; c++) {
      const long long xx = (long long)x + c, yy = (long long)y + r;
      const bool in = xx >= 0 && yy >= 0 && (uint64_t)xx < t->dim0 && (uint64_t)yy < t->dim1;
      d[r * t->box0 + c] = in ? t->base[(size_t)yy * t->stride1 + (size_t)xx] : 0;
    }
  return t->box0 * t->box1;
}

}  // namespace simt

// ---- runtime API ---------------------------------------------------------------------------------------
struct simt_stream { int id; };
struct simt_event { std::chrono::steady_clock::time_point t; };

namespace {
CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType, cuuint32_t rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                      CUtensorMapFloatOOBfill) {
  if (rank != 2) return CUDA_ERROR_INVALID_VALUE;
  simt::TensorMap2D t{};
  t.magic = 0x54454e534f524d41ull;
  t.base = static_cast<const uint8_t*>(base);
  t.dim0 = dims[0], t.dim1 = dims[1], t.stride1 = strides[0];
  t.box0 = box[0], t.box1 = box[1];
  memset(map, 0, sizeof *map);
  memcpy(map, &t, sizeof t);
  return CUDA_SUCCESS;
}
}  // namespace

extern "C" {
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0, *hi = 0; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated runtime error"; }
cudaError_t cudaMalloc(void** p, size_t n) {
  *p = nullptr;
  if (posix_memalign(p, 256, n ? n : 1)) return cudaErrorMemoryAllocation;
  memset(*p, 0xA5, n);  // device memory is not zeroed
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return posix_memalign(p, 256, n ? n : 1) ? cudaErrorMemoryAllocation : cudaSuccess; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
}
cudaError_t cudaMemcpy2D(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind) {
  for (size_t r = 0; r < h; r++) memmove(static_cast<char*>(d) + r * dp, static_cast<const char*>(s) + r * sp, w);
  return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind k, cudaStream_t) {
  return cudaMemcpy2D(d, dp, s, sp, w, h, k);
}
cudaError_t cudaMemset2D(void* d, size_t dp, int v, size_t w, size_t h) {
  for (size_t r = 0; r < h; r++) memset(static_cast<char*>(d) + r * dp, v, w);
  return cudaSuccess;
}
extern "C" {
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new simt_stream{0}; return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = new simt_stream{0}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaLaunchHostFunc(cudaStream_t, cudaHostFn_t fn, void* arg) { fn(arg); return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new simt_event{std::chrono::steady_clock::now()}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q) {
  const bool ok = strcmp(name, "cuTensorMapEncodeTiled") == 0;
  *fn = ok ? reinterpret_cast<void*>(&encode_tiled) : nullptr;
  if (q) *q = ok ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
  return cudaSuccess;
}
}

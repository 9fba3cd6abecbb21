// tests/hostsim/hostsim_runtime.cpp -- TEST INFRASTRUCTURE: the execution engine behind cuda_host_shim.h and a host
// stand-in for the handful of CUDA runtime calls the library makes ("device" memory is host memory).
#include "cuda_host_shim.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <pthread.h>
#include <sched.h>
#include <semaphore.h>

thread_local HostDim threadIdx, blockIdx, blockDim, gridDim;

// dynamic shared memory of the kernels (one block at a time)
alignas(16) unsigned char smem_raw[232448];
double lp_smem[232448 / 8];

namespace hostsim {
namespace {
constexpr int MAXT = 1024, MAXW = MAXT / 32;

struct Bar {
    std::mutex m;
    unsigned count = 0, gen = 0;
};
struct Warp {
    Bar bar[2][32];                          // [phase][first lane of the mask]: disjoint lane groups never share a barrier
    unsigned long long slot[32];
    std::atomic<unsigned> exited{0};
};
Warp warps[MAXW];
Bar block_bar;
std::atomic<unsigned> block_exited{0};

struct Worker {
    pthread_t th;
    sem_t go;
    int tid;
};
Worker workers[MAXT];
int n_workers = 0;
sem_t done_sem;
std::atomic<int> remaining{0};
void (*cur_thunk)(void *) = nullptr;
void *cur_ctx = nullptr;
unsigned cur_block = 0, cur_bdim = 1, cur_gdim = 1;

// wait until `n` participants arrived, counting the ones that already left the kernel (`gone()` of them)
template <class Gone> void bar_wait(Bar &b, unsigned n, Gone gone)
{
    std::unique_lock<std::mutex> lk(b.m);
    const unsigned g = b.gen;
    b.count++;
    for (;;) {
        if (b.gen != g) return;
        if (b.count + gone() >= n) { b.count = 0; b.gen++; return; }
        lk.unlock();
        sched_yield();
        lk.lock();
    }
}

void *worker_main(void *arg)
{
    Worker *w = static_cast<Worker *>(arg);
    for (;;) {
        sem_wait(&w->go);
        threadIdx = HostDim{(unsigned)w->tid, 0, 0};
        blockIdx = HostDim{cur_block, 0, 0};
        blockDim = HostDim{cur_bdim, 1, 1};
        gridDim = HostDim{cur_gdim, 1, 1};
        cur_thunk(cur_ctx);
        warps[w->tid >> 5].exited.fetch_or(1u << (w->tid & 31));     // a thread that left the kernel never blocks the others
        block_exited.fetch_add(1);
        if (remaining.fetch_sub(1) == 1) sem_post(&done_sem);
    }
    return nullptr;
}

void ensure_workers(int n)
{
    static bool init = false;
    if (!init) { sem_init(&done_sem, 0, 0); init = true; }
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, 4u << 20);
    for (; n_workers < n; n_workers++) {
        Worker &w = workers[n_workers];
        w.tid = n_workers;
        sem_init(&w.go, 0, 0);
        if (pthread_create(&w.th, &at, worker_main, &w) != 0) { std::fprintf(stderr, "hostsim: pthread_create failed\n"); std::abort(); }
    }
}
}  // namespace

void block_barrier()
{
    bar_wait(block_bar, cur_bdim, [] { return block_exited.load(); });
}

int block_or(int pred)
{
    static std::atomic<int> acc{0};
    if (pred) acc.store(1);
    block_barrier();
    const int v = acc.load();
    block_barrier();
    if (threadIdx.x == 0) acc.store(0);
    block_barrier();
    return v;
}

void warp_exchange(unsigned mask, unsigned long long v, unsigned long long out[32], unsigned *live_mask)
{
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    Warp &w = warps[tid >> 5];
    if (cur_bdim < 32) mask &= (1u << cur_bdim) - 1u;
    const unsigned first = (unsigned)__builtin_ffs((int)mask) - 1u, n = (unsigned)__builtin_popcount(mask);
    auto gone = [&] { return (unsigned)__builtin_popcount(w.exited.load() & mask); };
    w.slot[lane] = v;
    bar_wait(w.bar[0][first], n, gone);
    const unsigned live = mask & ~w.exited.load();
    for (int l = 0; l < 32; l++) out[l] = ((mask >> l) & 1u) ? w.slot[l] : 0ull;
    *live_mask = live | (1u << lane);
    bar_wait(w.bar[1][first], n, gone);      // nobody overwrites a slot before everybody has read it
}

void launch_impl(Cfg c, void (*thunk)(void *), void *ctx)
{
    static std::mutex launch_mutex;          // one kernel at a time (streams are not modelled)
    std::lock_guard<std::mutex> lk(launch_mutex);
    if (c.grid == 0 || c.block == 0) return;
    if (c.block > (unsigned)MAXT || c.smem > sizeof(smem_raw)) { std::fprintf(stderr, "hostsim: launch configuration out of range\n"); std::abort(); }
    ensure_workers((int)c.block);
    cur_thunk = thunk; cur_ctx = ctx; cur_bdim = c.block; cur_gdim = c.grid;
    for (unsigned b = 0; b < c.grid; b++) {
        for (unsigned wi = 0; wi < (c.block + 31) / 32; wi++) {
            warps[wi].exited.store(0);
            for (int ph = 0; ph < 2; ph++) for (int l = 0; l < 32; l++) { warps[wi].bar[ph][l].count = 0; }
        }
        block_bar.count = 0;
        block_exited.store(0);
        cur_block = b;
        remaining.store((int)c.block);
        for (unsigned t = 0; t < c.block; t++) sem_post(&workers[t].go);
        sem_wait(&done_sem);
    }
}
}  // namespace hostsim

// ------------------------------------------------------------------ CUDA runtime stand-ins ("device" memory == host memory)
extern "C" {
cudaError_t cudaMalloc(void **p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int) { *v = (a == cudaDevAttrMultiProcessorCount) ? 2 : 0; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)1; return cudaSuccess; }     // streams are not modelled: launches run in call order
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)1; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "hostsim"; }
cudaError_t cudaFuncSetAttribute(const void *, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorNotSupported; }          // one process, host memory: no IPC
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int *n, const void *, int, size_t, unsigned) { *n = 2; return cudaSuccess; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, const void *, int, size_t) { *n = 2; return cudaSuccess; }
}

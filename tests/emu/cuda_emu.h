// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  A host-side SIMT emulation that lets the CPU test-suite execute the
// *source* of psmc_b200/csrc/psmc_estep.cu + psmc_kernels.cuh (kernels and host orchestration) without a GPU: every CUDA
// thread is a cooperative fiber (all threads of a block on one host thread, a hand-rolled x86-64 context switch at every
// warp shuffle / vote / __syncthreads), blocks of a grid run on a pool of host threads, the CUDA runtime calls the
// library makes are mapped onto malloc/memcpy.  It exists to catch logic errors in the kernels before GPU time is spent;
// it is thousands of times slower than a CPU implementation would be and is built into tests/emu/libpsmc_b200_emu.so,
// which nothing under psmc_b200/ or host/ ever loads (the product has no CPU path).
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <sched.h>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local /* a block lives on one host thread */

struct dim3 {
	unsigned x, y, z;
	dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 {
	double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
using std::max;
using std::min;

namespace simt_emu {

// CUDA threads are cooperative fibers: all threads of a block live on ONE host thread and hand the CPU to each other
// at every warp shuffle / vote / __syncthreads (a hand-rolled x86-64 context switch, ~20 ns); different blocks of a
// grid run on different host threads.
extern "C" void psmc_emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl psmc_emu_switch
.type psmc_emu_switch,@function
psmc_emu_switch:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
.size psmc_emu_switch,.-psmc_emu_switch
)");

struct Barrier {
	int expected = 0, waiting = 0;
	uint64_t gen = 0;
	void init(int n)
	{
		expected = n;
		waiting = 0;
	}
	inline void wait();
	void drop()
	{
		--expected; // a thread left the kernel: the others may already be waiting for it
		if (expected > 0 && waiting >= expected) {
			waiting = 0;
			++gen;
		}
	}
};

struct Warp {
	Barrier bar;
	uint64_t slot[2][32];
	int parity_of[32];
};
struct Block {
	Barrier bar;
	std::vector<Warp> warps;
};
struct Fiber {
	void *sp = nullptr;
	char *stack = nullptr;
	bool done = false;
	dim3 tid, bid, bdim, gdim;
	int lane = 0;
	Warp *warp = nullptr;
	Block *block = nullptr;
	const std::function<void()> *entry = nullptr;
};
inline thread_local Fiber *cur = nullptr;
inline thread_local unsigned char *dyn_smem_ptr = nullptr; // dynamic shared memory of the block this host thread is running
inline unsigned char *dyn_smem() { return dyn_smem_ptr; }
inline thread_local void *sched_sp = nullptr;
inline thread_local uint64_t progress = 0;

inline void yield() { psmc_emu_switch(&cur->sp, sched_sp); }
inline void Barrier::wait()
{
	const uint64_t g = gen;
	if (++waiting >= expected) {
		waiting = 0;
		++gen;
		++progress;
	} else {
		while (gen == g) yield();
	}
}
inline void trampoline()
{
	Fiber *f = cur;
	(*f->entry)();
	f->done = true;
	++progress;
	f->warp->bar.drop();
	f->block->bar.drop();
	psmc_emu_switch(&f->sp, sched_sp);
	abort(); // never resumed
}

inline std::mutex &launch_mutex()
{
	static std::mutex m;
	return m;
}

template <class T>
inline T exchange(T v, int src_lane)
{
	static_assert(sizeof(T) <= 8, "shuffle of up to 64 bits");
	Warp *w = cur->warp;
	const int l = cur->lane;
	const int p = w->parity_of[l];
	w->parity_of[l] = p ^ 1;
	uint64_t bits = 0;
	memcpy(&bits, &v, sizeof(T));
	w->slot[p][l] = bits;
	w->bar.wait();
	const uint64_t r = w->slot[p][src_lane];
	T out;
	memcpy(&out, &r, sizeof(T));
	return out;
}

// 20-bit reciprocal seed, like rcp.approx.ftz.f64
inline double rcp_seed(double s)
{
	int e;
	const double m = frexp(s, &e);
	const float r = 1.0f / (float)m;
	uint32_t rb;
	memcpy(&rb, &r, 4);
	rb &= 0xfffffff8u; // 20 mantissa bits
	float rt;
	memcpy(&rt, &rb, 4);
	return ldexp((double)rt, -e);
}

template <class F>
inline void launch(dim3 grid, dim3 block, F body, size_t smem_bytes = 0);

} // namespace simt_emu

#define threadIdx (simt_emu::cur->tid)
#define blockIdx (simt_emu::cur->bid)
#define blockDim (simt_emu::cur->bdim)
#define gridDim (simt_emu::cur->gdim)


template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32)
{
	const int l = simt_emu::cur->lane, base = l - l % width, src = l - (int)d;
	const T r = simt_emu::exchange(v, src >= base ? src : l);
	return src >= base ? r : v;
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32)
{
	const int l = simt_emu::cur->lane, base = l - l % width, src = l + (int)d;
	const T r = simt_emu::exchange(v, src < base + width ? src : l);
	return src < base + width ? r : v;
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32)
{
	const int l = simt_emu::cur->lane, base = l - l % width, src = l ^ m;
	const bool ok = src >= base && src < base + width;
	const T r = simt_emu::exchange(v, ok ? src : l);
	return ok ? r : v;
}
template <class T>
inline T __shfl_sync(unsigned, T v, int srcLane, int width = 32)
{
	const int l = simt_emu::cur->lane, base = l - l % width;
	return simt_emu::exchange(v, base + (srcLane % width + width) % width);
}
inline int __any_sync(unsigned, int pred)
{
	int any = 0;
	// 32 exchanges would be slow: gather through the slots directly
	simt_emu::Warp *w = simt_emu::cur->warp;
	const int l = simt_emu::cur->lane;
	const int p = w->parity_of[l];
	w->parity_of[l] = p ^ 1;
	w->slot[p][l] = pred ? 1 : 0;
	w->bar.wait();
	for (int i = 0; i < 32; ++i) any |= (int)w->slot[p][i];
	return any;
}
inline void __syncthreads() { simt_emu::cur->block->bar.wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { simt_emu::cur->warp->bar.wait(); }

template <class T>
inline T __ldg(const T *p)
{
	return *p;
}
inline int __double2hiint(double x)
{
	uint64_t b;
	memcpy(&b, &x, 8);
	return (int)(b >> 32);
}
inline int __double2loint(double x)
{
	uint64_t b;
	memcpy(&b, &x, 8);
	return (int)(b & 0xffffffffu);
}
inline double __hiloint2double(int hi, int lo)
{
	const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
	double x;
	memcpy(&x, &b, 8);
	return x;
}
inline long long __double_as_longlong(double x)
{
	long long b;
	memcpy(&b, &x, 8);
	return b;
}
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v)
{
	unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
	while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
	}
	return old;
}

namespace simt_emu {
template <class F>
inline void launch(dim3 grid, dim3 block, F body, size_t smem_bytes)
{
	std::lock_guard<std::mutex> lk(launch_mutex());
	const int nthreads = (int)(block.x * block.y * block.z);
	const int nwarps = (nthreads + 31) / 32;
	const long nblocks = (long)grid.x * grid.y;
	if (nblocks <= 0 || nthreads <= 0) return;
	const std::function<void()> entry = body;
	std::atomic<long> next{0};
	const size_t STACK = 256 << 10;
	auto runner = [&]() {
		std::vector<Fiber> fib(nthreads);
		std::unique_ptr<char[]> stacks(new char[(size_t)nthreads * STACK + 64]); // untouched pages cost nothing
		Block blk;
		blk.warps = std::vector<Warp>(nwarps);
		std::unique_ptr<unsigned char[]> dsm(new unsigned char[smem_bytes + 256]);
		dyn_smem_ptr = (unsigned char *)(((uintptr_t)dsm.get() + 127) & ~(uintptr_t)127);
		for (;;) {
			const long b = next.fetch_add(1);
			if (b >= nblocks) break;
			blk.bar.init(nthreads);
			for (int w = 0; w < nwarps; ++w) {
				blk.warps[w].bar.init(std::min(32, nthreads - 32 * w));
				for (int l = 0; l < 32; ++l) blk.warps[w].parity_of[l] = 0;
			}
			for (int t = 0; t < nthreads; ++t) {
				Fiber &f = fib[t];
				f.done = false;
				f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
				f.bid = dim3((unsigned)(b % grid.x), (unsigned)(b / grid.x), 0);
				f.bdim = block;
				f.gdim = grid;
				f.lane = t % 32;
				f.warp = &blk.warps[t / 32];
				f.block = &blk;
				f.entry = &entry;
				uintptr_t top = (uintptr_t)(stacks.get() + (size_t)(t + 1) * STACK);
				top &= ~(uintptr_t)15;
				void **sp = (void **)top;
				*--sp = nullptr;               // fake return address of the trampoline
				*--sp = (void *)&trampoline;   // popped by the first switch's ret
				for (int r = 0; r < 6; ++r) *--sp = nullptr;
				f.sp = (void *)sp;
			}
			for (;;) {
				bool alive = false;
				const uint64_t before = progress;
				for (int t = 0; t < nthreads; ++t) {
					if (fib[t].done) continue;
					alive = true;
					cur = &fib[t];
					psmc_emu_switch(&sched_sp, fib[t].sp);
				}
				cur = nullptr;
				if (!alive) break;
				if (progress == before) {
					fprintf(stderr, "simt_emu: deadlock in block %ld (a barrier is waiting for threads that never arrive)\n", b);
					abort();
				}
			}
		}
	};
	const int nrun = (int)std::min<long>(nblocks, std::max(1u, std::thread::hardware_concurrency()));
	if (nrun <= 1) {
		std::thread t(runner); // own thread: thread_local __shared__ storage and a fresh stack
		t.join();
	} else {
		std::vector<std::thread> th;
		for (int i = 0; i < nrun; ++i) th.emplace_back(runner);
		for (auto &x : th) x.join();
	}
}
} // namespace simt_emu

// ---------------------------------------------------------------------------------------------------------------
// the CUDA runtime calls psmc_estep.cu makes, on host memory; everything is synchronous
// ---------------------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef struct emu_stream_ *cudaStream_t;
typedef struct emu_event_ *cudaEvent_t;
struct emu_stream_ {
	int unused;
};
struct emu_event_ {
	int unused;
};
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
struct cudaDeviceProp {
	int multiProcessorCount;
	char name[64];
};
static inline const char *cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated allocation failure"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n)
{
	*n = 1;
	return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
	const char *env = getenv("PSMC_EMU_SMS");
	p->multiProcessorCount = env ? atoi(env) : 2;
	snprintf(p->name, sizeof(p->name), "SIMT emulation (host)");
	return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void **p, size_t n)
{
	*p = calloc(1, n ? n : 1);
	return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = (size_t)8 << 30; *t = (size_t)16 << 30; return cudaSuccess; }
static inline cudaError_t cudaFree(void *p)
{
	free(p);
	return cudaSuccess;
}
static inline cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t)
{
	memmove(d, s, n);
	return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind)
{
	memmove(d, s, n);
	return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t)
{
	memset(d, v, n);
	return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned)
{
	*s = new emu_stream_();
	return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t s)
{
	delete s;
	return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e)
{
	*e = new emu_event_();
	return cudaSuccess;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e)
{
	delete e;
	return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t)
{
	*ms = 0.f;
	return cudaSuccess;
}
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class K>
static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <class K>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t)
{
	*n = 2;
	return cudaSuccess;
}

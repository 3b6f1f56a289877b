// tools/ubench.cu -- micro-benchmarks that size the E-step kernels on B200 (FP64 pipe, shared-memory
// broadcast loads, warp shuffles, uniform constant loads).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

struct BigParams { double c[832]; }; // 6656 B like K1Params

template <int ILP> __global__ void k_dfma(double *out, int iters, double a, double b)
{
	double x[ILP];
#pragma unroll
	for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
	}
	double s = 0;
#pragma unroll
	for (int i = 0; i < ILP; ++i) s += x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// LDS patterns: mode 0 uniform address (all lanes same 16B), 1: lane%8 distinct contiguous 16B (128B), 2: lane-parity two addresses 16B,
// 3: uniform 8B (LDS.64), 4: lane%8 distinct contiguous 8B
template <int MODE> __global__ void k_lds(double *out, int iters)
{
	__shared__ __align__(16) double sm[2048];
	for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * 1e-3;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	double acc0 = 0, acc1 = 0;
	int base = 0;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int j = 0; j < 16; ++j) {
			if (MODE == 0) { double2 v = *reinterpret_cast<double2 *>(&sm[(base + j * 2) & 2046]); acc0 += v.x; acc1 += v.y; }
			if (MODE == 1) { double2 v = *reinterpret_cast<double2 *>(&sm[((base + j * 16) & 1023) + (lane & 7) * 2]); acc0 += v.x; acc1 += v.y; }
			if (MODE == 2) { double2 v = *reinterpret_cast<double2 *>(&sm[((base + j * 2) & 1022) + (lane & 1) * 1024]); acc0 += v.x; acc1 += v.y; }
			if (MODE == 3) { acc0 += sm[(base + j) & 2047]; }
			if (MODE == 4) { acc0 += sm[((base + j * 8) & 1023) + (lane & 7)]; }
		}
		base += 32;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1;
}
__global__ void k_shfl(double *out, int iters)
{
	double x = threadIdx.x, y = threadIdx.x * 0.5;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int j = 0; j < 8; ++j) { x += __shfl_up_sync(0xffffffffu, y, 1); y += __shfl_down_sync(0xffffffffu, x, 2); }
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
}
__global__ void k_shfl_lat(double *out, int iters)
{
	double x = threadIdx.x;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int j = 0; j < 16; ++j) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
// constant-bank streaming: every DFMA takes a different constant from a 6.6 KB kernel parameter
__global__ void k_ldcu(const __grid_constant__ BigParams P, double *out, int iters)
{
	double x0 = threadIdx.x, x1 = 1, x2 = 2, x3 = 3;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int j = 0; j < 832; j += 4) { x0 = fma(x0, P.c[j], 1.0); x1 = fma(x1, P.c[j + 1], 1.0); x2 = fma(x2, P.c[j + 2], 1.0); x3 = fma(x3, P.c[j + 3], 1.0); }
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
// same but only 64 distinct constants (512 B) reused
__global__ void k_ldcu_small(const __grid_constant__ BigParams P, double *out, int iters)
{
	double x0 = threadIdx.x, x1 = 1, x2 = 2, x3 = 3;
	for (int it = 0; it < iters * 13; ++it) {
#pragma unroll
		for (int j = 0; j < 64; j += 4) { x0 = fma(x0, P.c[j], 1.0); x1 = fma(x1, P.c[j + 1], 1.0); x2 = fma(x2, P.c[j + 2], 1.0); x3 = fma(x3, P.c[j + 3], 1.0); }
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
__global__ void k_bar(double *out, int iters)
{
	double x = threadIdx.x;
	const int id = 1 + (threadIdx.x >> 6);
	for (int it = 0; it < iters; ++it) {
		x = fma(x, 1.0000001, 0.5);
		asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <typename F> float timeit(F f)
{
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	f(); cudaDeviceSynchronize();
	cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
	float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main()
{
	cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
	int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
	printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, clk_khz);
	const int SM = p.multiProcessorCount;
	double *out; CK(cudaMalloc(&out, sizeof(double) * SM * 32 * 1024));
	BigParams P; for (int i = 0; i < 832; ++i) P.c[i] = 1.0 - 1e-9 * i;
	for (int wps = 1; wps <= 16; wps *= 2) { // warps per SMSP
		const int threads = 128 * wps > 1024 ? 1024 : 128 * wps, blocks = SM * ((128 * wps + threads - 1) / threads);
		const int iters = 20000;
		float ms;
		ms = timeit([&] { k_dfma<1><<<blocks, threads>>>(out, iters, 1.0000001, 0.5); });
		printf("wps %2d  DFMA ILP1: %.3f cyc/instr/warp-chain  (%.2f TFLOPS)\n", wps, ms * 1e-3 * clk_khz * 1e3 / iters, 2.0 * blocks * threads * iters / (ms * 1e-3) / 1e12);
		ms = timeit([&] { k_dfma<4><<<blocks, threads>>>(out, iters, 1.0000001, 0.5); });
		printf("wps %2d  DFMA ILP4: %.2f TFLOPS\n", wps, 2.0 * 4 * blocks * threads * iters / (ms * 1e-3) / 1e12);
		ms = timeit([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 0.5); });
		printf("wps %2d  DFMA ILP8: %.2f TFLOPS\n", wps, 2.0 * 8 * blocks * threads * iters / (ms * 1e-3) / 1e12);
	}
	{
		const int threads = 512, blocks = SM * 2, iters = 4000;
		const char *nm[5] = {"LDS.128 uniform", "LDS.128 8x16B contiguous", "LDS.128 2 addresses", "LDS.64 uniform", "LDS.64 8x8B contiguous"};
		float ms[5];
		ms[0] = timeit([&] { k_lds<0><<<blocks, threads>>>(out, iters); });
		ms[1] = timeit([&] { k_lds<1><<<blocks, threads>>>(out, iters); });
		ms[2] = timeit([&] { k_lds<2><<<blocks, threads>>>(out, iters); });
		ms[3] = timeit([&] { k_lds<3><<<blocks, threads>>>(out, iters); });
		ms[4] = timeit([&] { k_lds<4><<<blocks, threads>>>(out, iters); });
		for (int m = 0; m < 5; ++m) {
			double warp_instr_per_sm = (double)blocks / SM * (threads / 32) * iters * 16;
			printf("%-28s %.2f SM-cycles per warp-LDS (incl. the dependent DADDs)\n", nm[m], ms[m] * 1e-3 * clk_khz * 1e3 / warp_instr_per_sm);
		}
	}
	{
		const int threads = 512, blocks = SM * 2, iters = 4000;
		float ms = timeit([&] { k_shfl<<<blocks, threads>>>(out, iters); });
		double wi = (double)blocks / SM * (threads / 32) * iters * 16 * 2; // 16 double shuffles = 32 SHFL.32
		printf("SHFL.32: %.2f SM-cycles per warp-SHFL (throughput, with DADDs)\n", ms * 1e-3 * clk_khz * 1e3 / wi);
		ms = timeit([&] { k_shfl_lat<<<SM, 32>>>(out, iters); });
		printf("double shfl+DADD dependent: %.1f cycles per (2 SHFL + DADD)\n", ms * 1e-3 * clk_khz * 1e3 / (iters * 16.0));
		ms = timeit([&] { k_dfma<1><<<SM, 32>>>(out, 20000, 1.0000001, 0.5); });
		printf("DFMA dependent latency: %.1f cycles\n", ms * 1e-3 * clk_khz * 1e3 / 20000.0);
	}
	for (int wps = 1; wps <= 4; wps *= 2) {
		const int threads = 128 * wps, blocks = SM, iters = 200;
		float ms = timeit([&] { k_ldcu<<<blocks, threads>>>(P, out, iters); });
		printf("wps %d  DFMA with streaming 6.6KB constants: %.2f cycles per DFMA per SMSP-warp (%.2f TFLOPS)\n", wps, ms * 1e-3 * clk_khz * 1e3 / (iters * 832.0 * wps), 2.0 * blocks * threads * iters * 832 / (ms * 1e-3) / 1e12);
		ms = timeit([&] { k_ldcu_small<<<blocks, threads>>>(P, out, iters); });
		printf("wps %d  DFMA with 512B constants reused:     %.2f cycles per DFMA per SMSP-warp (%.2f TFLOPS)\n", wps, ms * 1e-3 * clk_khz * 1e3 / (iters * 832.0 * wps), 2.0 * blocks * threads * iters * 832 / (ms * 1e-3) / 1e12);
	}
	{
		const int threads = 128, blocks = SM * 3, iters = 20000;
		float ms = timeit([&] { k_bar<<<blocks, threads>>>(out, iters); });
		printf("bar.sync(64 threads)+DFMA loop: %.1f cycles per iteration (3 blocks of 2 pairs per SM)\n", ms * 1e-3 * clk_khz * 1e3 / iters);
	}
	return 0;
}

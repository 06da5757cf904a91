// tools/dmma_peak.cu -- FP64 micro-benchmarks for the roofline denominators of the 20- / 61-state kernels:
// register-resident mma.sync f64 throughput (m8n8k4, m16n8k4, m16n8k8, m16n8k16) and plain DFMA throughput.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak tools/dmma_peak.cu && ./dmma_peak
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 8

template <int SHAPE>
__global__ void __launch_bounds__(256) k_mma(double *out, double seed) {
	double acc[NACC][4];
	for (int i = 0; i < NACC; i++)
		for (int j = 0; j < 4; j++) acc[i][j] = seed * (i + j);
	double a[8], b[4];
	for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-9 + i;
	for (int i = 0; i < 4; i++) b[i] = seed - threadIdx.x * 1e-9 - i;
	for (int it = 0; it < ITERS; it++) {
#pragma unroll
		for (int i = 0; i < NACC; i++) {
			if (SHAPE == 0)
				asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a[0]), "d"(b[0]));
			else if (SHAPE == 1)
				asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
				             : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
				             : "d"(a[0]), "d"(a[1]), "d"(b[0]));
			else if (SHAPE == 2)
				asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
				             : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
				             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
			else
				asm volatile(
				    "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
				    : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
				    : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
		}
	}
	double s = 0;
	for (int i = 0; i < NACC; i++)
		for (int j = 0; j < 4; j++) s += acc[i][j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_fma(double *out, double seed) {
	double acc[16];
	for (int i = 0; i < 16; i++) acc[i] = seed * i;
	const double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
	for (int it = 0; it < ITERS; it++) {
#pragma unroll
		for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
	}
	double s = 0;
	for (int i = 0; i < 16; i++) s += acc[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static double time_ms(F launch) {
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	launch();
	cudaDeviceSynchronize();
	float best = 1e30f;
	for (int r = 0; r < 5; r++) {
		cudaEventRecord(e0);
		launch();
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		if (ms < best) best = ms;
	}
	return best;
}

int main() {
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	const int sms = prop.multiProcessorCount;
	double *out;
	cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
	const char *names[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
	const double flop[4] = {2.0 * 8 * 8 * 4, 2.0 * 16 * 8 * 4, 2.0 * 16 * 8 * 8, 2.0 * 16 * 8 * 16};
	for (int ctas = 1; ctas <= 4; ctas *= 2) {
		const int grid = sms * ctas;
		double ms[4];
		ms[0] = time_ms([&] { k_mma<0><<<grid, 256>>>(out, 1.0); });
		ms[1] = time_ms([&] { k_mma<1><<<grid, 256>>>(out, 1.0); });
		ms[2] = time_ms([&] { k_mma<2><<<grid, 256>>>(out, 1.0); });
		ms[3] = time_ms([&] { k_mma<3><<<grid, 256>>>(out, 1.0); });
		for (int s = 0; s < 4; s++) {
			const double total = flop[s] * NACC * ITERS * 8.0 * grid;  // 8 warps per CTA
			printf("{\"kernel\": \"dmma_%s\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", names[s], ctas, ms[s], total / ms[s] / 1e9);
		}
		const double fms = time_ms([&] { k_fma<<<grid, 256>>>(out, 1.0); });
		printf("{\"kernel\": \"dfma\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", ctas, fms, 2.0 * 16 * ITERS * 256.0 * grid / fms / 1e9);
	}
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) {
		printf("error: %s\n", cudaGetErrorString(e));
		return 1;
	}
	return 0;
}

// Roofline denominators measured on the device the library runs on: the FP64 FMA rate (the bound of the IMPLSCH kernels,
// SURVEY.md 8d) and a streaming copy (cross-check of MEASURED_PEAKS.json's HBM figure).  Not part of the hot path.
#include "internal.h"

namespace ew {

// 16 independent DFMA chains per thread: two accumulators per chain pair keep the FP64 pipe (2 issue cycles per warp
// instruction) full without any memory traffic.  The result is stored so that the chains are not dead code.
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (double)(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_copy_peak(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

}  // namespace ew

extern "C" int ecwam_b200_measure_peaks(double* fp64_tflops, double* copy_gbs) {
  using namespace ew;
  int dev = 0, nsm = 0;
  EW_CUDA_CHECK(cudaGetDevice(&dev));
  EW_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  cudaEvent_t e0, e1;
  EW_CUDA_CHECK(cudaEventCreate(&e0));
  EW_CUDA_CHECK(cudaEventCreate(&e1));
  if (fp64_tflops) {
    const int nb = nsm * 8, nt = 256, iters = 4096;
    double* out = nullptr;
    EW_CUDA_CHECK(cudaMalloc(&out, (size_t)nb * nt * sizeof(double)));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
      EW_CUDA_CHECK(cudaEventRecord(e0));
      k_dfma_peak<<<nb, nt>>>(out, iters, 0.999999, 1e-9);
      EW_CUDA_CHECK(cudaEventRecord(e1));
      EW_CUDA_CHECK(cudaEventSynchronize(e1));
      float ms = 0.f;
      EW_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
      const double tf = 2.0 * 16.0 * iters * (double)nb * nt / (ms * 1e-3) / 1e12;
      if (rep > 0 && tf > best) best = tf;
    }
    cudaFree(out);
    *fp64_tflops = best;
  }
  if (copy_gbs) {
    const size_t n = (size_t)1 << 27;   // 2 GiB read + 2 GiB written per pass
    double2 *a = nullptr, *b = nullptr;
    EW_CUDA_CHECK(cudaMalloc(&a, n * sizeof(double2)));
    EW_CUDA_CHECK(cudaMalloc(&b, n * sizeof(double2)));
    EW_CUDA_CHECK(cudaMemset(a, 0, n * sizeof(double2)));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
      EW_CUDA_CHECK(cudaEventRecord(e0));
      k_copy_peak<<<nsm * 16, 256>>>(a, b, n);
      EW_CUDA_CHECK(cudaEventRecord(e1));
      EW_CUDA_CHECK(cudaEventSynchronize(e1));
      float ms = 0.f;
      EW_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
      const double gbs = 2.0 * (double)n * sizeof(double2) / (ms * 1e-3) / 1e9;
      if (rep > 0 && gbs > best) best = gbs;
    }
    cudaFree(a);
    cudaFree(b);
    *copy_gbs = best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  EW_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// tools/kb_crossover.cu -- where would tensor cores pay for the KBRL dictionary products?  (BASELINE north_star: "the
// dictionary Gram product taken to tensor cores only past a measured crossover")
//
// The only matrix-shaped work of Projectron.update is d* = K^-1 k (algorithms/projectron.py:42): a D x D fp64 mat-VEC,
// once per mistake, on a matrix that changes at every insertion.  fp64 is required (delta = 1 - d*.k is compared with
// eta = 0.1 on an ill-conditioned K^-1), so the only tensor path is the fp64 DMMA (mma.sync m8n8k4; tcgen05 has no fp64
// kind).  This microbenchmark times, for L learners with their own K^-1 streaming from HBM (like the update kernel):
//   fma1   y = A x          one CTA per learner, thread per row, coalesced columns (A symmetric): what kbrl.cu does
//   dmma1  the same product through DMMA with x in column 0 of an 8-wide B operand (7 wasted columns)
//   fma8 / dmma8   Y = A X for EIGHT right-hand sides at once -- the GEMM-shaped variant (d* of 8 consecutive candidate
//          allocations speculated from one pass over K^-1; only valid while no candidate inserts a landmark)
// and prints one JSON line per D with time per learner, achieved GB/s (bytes = D^2 x 8 read once) and GFLOP/s.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/kb_crossover tools/kb_crossover.cu && /tmp/kb_crossover
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// y = A x, thread per row i reading column i (coalesced; A symmetric), x in shared memory; NR right-hand sides
template <int NR>
__global__ void __launch_bounds__(256) fma_kernel(const double *__restrict__ A, const double *__restrict__ X, double *__restrict__ Y, int D) {
    extern __shared__ double xs[];                           // [NR][D]
    const double *a = A + (size_t)blockIdx.x * D * D;
    for (int i = threadIdx.x; i < NR * D; i += blockDim.x) xs[i] = X[(size_t)blockIdx.x * NR * D + i];
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        double acc[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[r] = 0.0;
#pragma unroll 8
        for (int j = 0; j < D; ++j) {
            const double v = a[(size_t)j * D + i];
#pragma unroll
            for (int r = 0; r < NR; ++r) acc[r] += v * xs[r * D + j];
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) Y[((size_t)blockIdx.x * NR + r) * D + i] = acc[r];
    }
}

// Y = A X through mma.sync.m8n8k4.f64: a warp owns 8 output rows; per k-step of 4 it feeds A[8 x 4] (row-major fragment:
// lane -> row lane / 4, col lane % 4) and X[4 x 8] (lane -> k lane % 4, column lane / 4).  NR = 1 leaves columns 1..7 zero.
// A rows are staged through shared memory in 8 x 32 slabs so that the global loads are coalesced 256-byte rows.
template <int NR>
__global__ void __launch_bounds__(256) dmma_kernel(const double *__restrict__ A, const double *__restrict__ X, double *__restrict__ Y, int D) {
    extern __shared__ double sm[];
    double *xs = sm;                                         // [8][D] (columns NR.. are zero)
    double *slab = sm + 8 * D;                               // [warps][8][33]
    const double *a = A + (size_t)blockIdx.x * D * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 8 * D; i += blockDim.x) { const int r = i / D; xs[i] = r < NR ? X[(size_t)blockIdx.x * NR * D + i] : 0.0; }
    __syncthreads();
    double *my = slab + warp * 8 * 33;
    for (int row0 = warp * 8; row0 < D; row0 += nw * 8) {
        double c0 = 0.0, c1 = 0.0;
        for (int k0 = 0; k0 < D; k0 += 32) {
#pragma unroll
            for (int r = 0; r < 8; ++r) my[r * 33 + lane] = a[(size_t)(row0 + r) * D + k0 + lane];   // coalesced row segments
            __syncwarp();
#pragma unroll
            for (int kk = 0; kk < 32; kk += 4) {
                const double av = my[(lane >> 2) * 33 + kk + (lane & 3)];
                const double bv = xs[(lane >> 2) * D + k0 + kk + (lane & 3)];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv));
            }
            __syncwarp();
        }
        const int row = row0 + (lane >> 2), col = 2 * (lane & 3);
        if (col < NR) Y[((size_t)blockIdx.x * NR + col) * D + row] = c0;
        if (col + 1 < NR) Y[((size_t)blockIdx.x * NR + col + 1) * D + row] = c1;
    }
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main() {
    const int Ds[] = {64, 128, 256, 512, 1024};
    const size_t budget = (size_t)6 << 30;                   // bytes of K^-1 per configuration (> L2: every pass streams from HBM)
    for (int D : Ds) {
        const int L = (int)(budget / ((size_t)D * D * 8));
        double *A, *X, *Y;
        CK(cudaMalloc(&A, (size_t)L * D * D * 8)); CK(cudaMalloc(&X, (size_t)L * 8 * D * 8)); CK(cudaMalloc(&Y, (size_t)L * 8 * D * 8));
        {   // symmetric test matrices, identical in every learner (values do not matter for timing; used for the check below)
            std::vector<double> a((size_t)D * D), x((size_t)8 * D);
            for (int i = 0; i < D; ++i) for (int j = 0; j <= i; ++j) a[(size_t)i * D + j] = a[(size_t)j * D + i] = 1.0 / (1 + i + j) + (i == j);
            for (int i = 0; i < 8 * D; ++i) x[i] = ((i * 37) % 101) / 101.0 - 0.5;
            for (int l = 0; l < L; ++l) {
                CK(cudaMemcpy(A + (size_t)l * D * D, a.data(), a.size() * 8, cudaMemcpyHostToDevice));
                CK(cudaMemcpy(X + (size_t)l * 8 * D, x.data(), x.size() * 8, cudaMemcpyHostToDevice));
            }
        }
        const size_t sm_f1 = (size_t)D * 8, sm_f8 = (size_t)8 * D * 8, sm_d = (size_t)8 * D * 8 + 8 * 8 * 33 * 8;
        CK(cudaFuncSetAttribute(fma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_f8));
        CK(cudaFuncSetAttribute(dmma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_d));
        CK(cudaFuncSetAttribute(dmma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_d));
        const float t_f1 = time_ms([&] { fma_kernel<1><<<L, 256, sm_f1>>>(A, X, Y, D); }, 5);
        std::vector<double> y_f((size_t)D), y_d((size_t)D);
        CK(cudaMemcpy(y_f.data(), Y, D * 8, cudaMemcpyDeviceToHost));
        const float t_d1 = time_ms([&] { dmma_kernel<1><<<L, 256, sm_d>>>(A, X, Y, D); }, 5);
        CK(cudaMemcpy(y_d.data(), Y, D * 8, cudaMemcpyDeviceToHost));
        double err = 0.0;
        for (int i = 0; i < D; ++i) err = fmax(err, fabs(y_f[i] - y_d[i]));
        const float t_f8 = time_ms([&] { fma_kernel<8><<<L, 256, sm_f8>>>(A, X, Y, D); }, 5);
        const float t_d8 = time_ms([&] { dmma_kernel<8><<<L, 256, sm_d>>>(A, X, Y, D); }, 5);
        CK(cudaGetLastError());
        const double bytes = (double)L * D * D * 8, fl1 = 2.0 * L * D * D;
        printf("{\"D\": %d, \"learners\": %d, \"matrix_gb\": %.2f, \"fma1_ms\": %.3f, \"fma1_gbs\": %.0f, \"dmma1_ms\": %.3f, \"dmma1_gbs\": %.0f, "
               "\"fma8_ms\": %.3f, \"fma8_gflops\": %.0f, \"dmma8_ms\": %.3f, \"dmma8_gflops\": %.0f, \"max_abs_diff_fma_vs_dmma\": %.3g}\n",
               D, L, bytes / 1e9, t_f1, bytes / t_f1 / 1e6, t_d1, bytes / t_d1 / 1e6, t_f8, 8 * fl1 / t_f8 / 1e6, t_d8, 8 * fl1 / t_d8 / 1e6, err);
        fflush(stdout);
        CK(cudaFree(A)); CK(cudaFree(X)); CK(cudaFree(Y));
    }
    return 0;
}

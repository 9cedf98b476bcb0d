// How long does one warp take to pull 32 scattered rows of 12 x 16 B from L2 (one row per lane, all loads independent)?
// Variants: strong (ld.relaxed.gpu) / ld.global.cg / ld.volatile / cp.async.cg; 1 warp alone or 444 warps at once.
#include <cstdio>
#include <cuda_runtime.h>
typedef ulonglong2 W;
template <int MODE> __device__ __forceinline__ W ld(const W* p) {
    W v;
    if (MODE == 0) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    else if (MODE == 1) asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    else asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
template <int MODE, int U>
__global__ void gather(const W* buf, const int* rows, int reps, long long* out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    unsigned long long acc = 0;
    long long total = 0;
    for (int r = 0; r < reps; ++r) {
        const W* row = buf + (long long)rows[(warp * reps + r) * 32 + lane] * U;
        __syncwarp();
        long long t0 = clock64();
        W w[U];
#pragma unroll
        for (int j = 0; j < U; ++j) w[j] = ld<MODE>(row + j);
#pragma unroll
        for (int j = 0; j < U; ++j) acc += w[j].x ^ w[j].y;
        __syncwarp();
        total += clock64() - t0;
    }
    if (lane == 0) { out[warp * 2] = total / reps; out[warp * 2 + 1] = (long long)acc; }
}
template <int MODE, int U> void run(const char* name, const W* d, const int* drows, long long* dout, int warps) {
    long long* h = new long long[warps * 2];
    for (int rep = 0; rep < 2; ++rep) { gather<MODE, U><<<(warps + 2) / 3, 96>>>(d, drows, 200, dout); cudaDeviceSynchronize(); }
    cudaMemcpy(h, dout, warps * 16, cudaMemcpyDeviceToHost);
    long long s = 0, mx = 0;
    for (int i = 0; i < warps; ++i) { s += h[2 * i]; if (h[2 * i] > mx) mx = h[2 * i]; }
    printf("%-18s U=%2d warps=%3d : mean %lld cycles, max %lld cycles per 32-row gather\n", name, U, warps, s / warps, mx);
    delete[] h;
}
int main() {
    const int NROWS = 8192, U = 12, MAXW = 444, REPS = 200;
    W* d; cudaMalloc(&d, (size_t)NROWS * U * sizeof(W)); cudaMemset(d, 1, (size_t)NROWS * U * sizeof(W));
    int* hr = new int[MAXW * REPS * 32];
    unsigned long long x = 99;
    for (int i = 0; i < MAXW * REPS * 32; ++i) { x = x * 6364136223846793005ull + 1442695040888963407ull; hr[i] = (int)((x >> 33) % NROWS); }
    int* drows; cudaMalloc(&drows, sizeof(int) * MAXW * REPS * 32); cudaMemcpy(drows, hr, sizeof(int) * MAXW * REPS * 32, cudaMemcpyHostToDevice);
    long long* dout; cudaMalloc(&dout, MAXW * 16);
    for (int warps : {1, 444}) {
        run<0, 12>("ld.relaxed.gpu", d, drows, dout, warps);
        run<1, 12>("ld.global.cg", d, drows, dout, warps);
        run<2, 12>("ld.volatile", d, drows, dout, warps);
        run<0, 1>("ld.relaxed.gpu", d, drows, dout, warps);
        run<1, 1>("ld.global.cg", d, drows, dout, warps);
    }
    return 0;
}

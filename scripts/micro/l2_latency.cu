// Micro-benchmarks behind the resident kernel's design: dependent L2 loads, a cp.async round trip, and the one-way latency of a
// tagged-word hand-over between two SMs (ping-pong).   nvcc -arch=sm_100a -O3 -o l2lat l2_latency.cu && ./l2lat
#include <cstdio>
#include <cuda_runtime.h>
typedef ulonglong2 W;
__device__ __forceinline__ W ldv(const W* p) { W v; asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ W ldr(const W* p) { W v; asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ W ldcgw(const W* p) { W v; asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void cpa(void* s, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g) : "memory"); }

__global__ void chase(const W* buf, int n, int mode, long long* out) {
    // buf[i].x = index of the next element (pointer chase through n words spread over 4 MB)
    __shared__ W sm[64];
    unsigned long long idx = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        W v;
        if (mode == 0) v = ldv(buf + idx);
        else if (mode == 1) v = ldr(buf + idx);
        else if (mode == 2) v = ldcgw(buf + idx);
        else { cpa(sm + threadIdx.x, buf + idx); asm volatile("cp.async.wait_all;" ::: "memory"); v = sm[threadIdx.x]; }
        idx = v.x;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (t1 - t0) / n; out[1] = (long long)idx; }
}

// ping-pong: CTA 0 and CTA 1 (different SMs) bounce an epoch through two tagged words; mode selects the store / load flavour
__global__ void pingpong(W* a, W* b, int n, int mode, long long* out) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    long long t0 = clock64();
    for (int k = 1; k <= n; ++k) {
        if (me == 0) {
            W v = make_ulonglong2(k, k);
            if (mode == 0) __stcg(a, v); else asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(a), "l"(v.x), "l"(v.y) : "memory");
            for (;;) { W r = mode == 0 ? ldv(b) : ldr(b); if (r.x == (unsigned long long)k && r.y == (unsigned long long)k) break; }
        } else {
            for (;;) { W r = mode == 0 ? ldv(a) : ldr(a); if (r.x == (unsigned long long)k && r.y == (unsigned long long)k) break; }
            W v = make_ulonglong2(k, k);
            if (mode == 0) __stcg(b, v); else asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(b), "l"(v.x), "l"(v.y) : "memory");
        }
    }
    long long t1 = clock64();
    if (me == 0) out[0] = (t1 - t0) / n;      // cycles per round trip = 2 hand-overs
}

int main() {
    const int N = 1 << 18;      // 4 MB of 16 B words
    W* h = new W[N];
    unsigned long long x = 12345;
    // random cyclic permutation
    int* perm = new int[N];
    for (int i = 0; i < N; ++i) perm[i] = i;
    for (int i = N - 1; i > 0; --i) { x = x * 6364136223846793005ull + 1442695040888963407ull; int j = (int)((x >> 33) % (unsigned)(i + 1)); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
    for (int i = 0; i < N; ++i) { h[perm[i]].x = perm[(i + 1) % N]; h[perm[i]].y = 0; }
    W* d; long long* out; cudaMalloc(&d, N * sizeof(W)); cudaMalloc(&out, 64);
    cudaMemcpy(d, h, N * sizeof(W), cudaMemcpyHostToDevice);
    long long ho[2];
    const char* names[] = {"ld.volatile", "ld.relaxed.gpu", "ld.global.cg", "cp.async.cg + wait + lds"};
    for (int rep = 0; rep < 2; ++rep)
        for (int mode = 0; mode < 4; ++mode) {
            chase<<<1, 1>>>(d, 20000, mode, out);
            cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
            if (rep) printf("dependent %-26s : %lld cycles per load (L2-resident 4 MB chase)\n", names[mode], ho[0]);
        }
    W* pp; cudaMalloc(&pp, 4096); cudaMemset(pp, 0, 4096);
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(pp, 0, 4096);
        pingpong<<<2, 32>>>(pp, pp + 64, 20000, mode, out);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(ho, out, 8, cudaMemcpyDeviceToHost);
        printf("ping-pong (%s): %lld cycles per round trip = %lld per hand-over   [%s]\n", mode == 0 ? "st.cg / ld.volatile" : "st.relaxed.gpu / ld.relaxed.gpu", ho[0], ho[0] / 2, cudaGetErrorString(e));
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("SM clock %d kHz\n", clk);
    return 0;
}

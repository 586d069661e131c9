// pcie_probe.cu -- what a resident kernel pays to talk to pinned host memory (B200, PCIe).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie_probe pcie_probe.cu && ./pcie_probe
// Every figure is the median over 200 repetitions of a globaltimer interval taken inside the kernel.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

__device__ __forceinline__ unsigned long long gt() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint4 ldv4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ldv8(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// grid x block threads read `bytes` of host memory once (16 bytes per thread and round); out[rep] = ns
// between the first CTA's start and the last CTA's end (device counters).
__global__ void k_read(const uint4 *host, int n16, unsigned long long *t_first, unsigned long long *t_last,
                       unsigned int *sink) {
    const unsigned long long t0 = gt();
    unsigned int acc = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) {
        const uint4 v = ldv4(host + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    acc = __reduce_xor_sync(0xffffffffu, acc);
    __syncthreads();
    const unsigned long long t1 = gt();
    if (threadIdx.x == 0) {
        atomicMin(t_first, t0);
        atomicMax(t_last, t1);
        if (acc == 0x12345) *sink = acc;
    }
}

// every CTA reads ONE 8-byte word of host memory (its own)
__global__ void k_read8(const unsigned long long *host, unsigned long long *t_first, unsigned long long *t_last,
                        unsigned int *sink) {
    if (threadIdx.x != 0) return;
    const unsigned long long t0 = gt();
    const unsigned long long v = ldv8(host + blockIdx.x);
    const unsigned long long t1 = gt() + (v == 0x1234567 ? 1 : 0);
    atomicMin(t_first, t0);
    atomicMax(t_last, t1);
}

// one thread: posted 8-byte write to host memory, then (mode 1) __threadfence_system, (mode 2) __threadfence
__global__ void k_fence(unsigned long long *host, int mode, unsigned long long *out) {
    const unsigned long long t0 = gt();
    host[threadIdx.x] = t0;
    if (mode == 1) __threadfence_system();
    if (mode == 2) __threadfence();
    const unsigned long long t1 = gt();
    if (threadIdx.x == 0) *out = t1 - t0;
}

// ping-pong: the kernel spins on a host word, answers by writing another host word; the host measures the
// round trip with its own clock (host write -> device sees -> device writes -> host sees)
__global__ void k_pingpong(volatile unsigned long long *ping, volatile unsigned long long *pong, int rounds) {
    for (int r = 1; r <= rounds; r++) {
        while (ldv8((const unsigned long long *)ping) != (unsigned long long)r) {}
        *pong = (unsigned long long)r;
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

static double median(std::vector<double> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }

#include <time.h>
static double now_us() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; }

int main() {
    unsigned char *host;
    CK(cudaHostAlloc(&host, 1 << 20, cudaHostAllocMapped));
    for (int i = 0; i < (1 << 20); i++) host[i] = (unsigned char)i;
    unsigned long long *dt;
    unsigned int *sink;
    CK(cudaMalloc(&dt, 64));
    CK(cudaMalloc(&sink, 4));
    auto run_read = [&](int grid, int block, int bytes) {
        std::vector<double> v;
        for (int rep = 0; rep < 200; rep++) {
            unsigned long long init[2] = {~0ull, 0ull};
            CK(cudaMemcpy(dt, init, 16, cudaMemcpyHostToDevice));
            host[rep & 1023] ^= 1;                       // the host touches the buffer between reads
            k_read<<<grid, block>>>((const uint4 *)host, bytes / 16, dt, dt + 1, sink);
            CK(cudaDeviceSynchronize());
            unsigned long long r[2];
            CK(cudaMemcpy(r, dt, 16, cudaMemcpyDeviceToHost));
            v.push_back((double)(r[1] - r[0]));
        }
        printf("read %6d B of host memory, %4d CTAs x %4d threads: %8.0f ns\n", bytes, grid, block, median(v));
    };
    run_read(1, 32, 16);
    run_read(1, 32, 512);
    run_read(1, 256, 4096);
    run_read(1, 1024, 4096);
    run_read(2, 128, 4096);
    run_read(8, 32, 4096);
    run_read(32, 32, 4096);
    run_read(1, 256, 2048);
    run_read(1, 1024, 16384);
    run_read(8, 128, 16384);
    run_read(64, 32, 16384);
    for (int grid : {1, 64, 148, 592}) {
        std::vector<double> v;
        for (int rep = 0; rep < 200; rep++) {
            unsigned long long init[2] = {~0ull, 0ull};
            CK(cudaMemcpy(dt, init, 16, cudaMemcpyHostToDevice));
            k_read8<<<grid, 32>>>((const unsigned long long *)host, dt, dt + 1, sink);
            CK(cudaDeviceSynchronize());
            unsigned long long r[2];
            CK(cudaMemcpy(r, dt, 16, cudaMemcpyDeviceToHost));
            v.push_back((double)(r[1] - r[0]));
        }
        printf("one 8-byte host word per CTA, %4d CTAs at once:        %8.0f ns\n", grid, median(v));
    }
    for (int mode = 0; mode < 3; mode++) {
        for (int threads : {1, 32}) {
            std::vector<double> v;
            for (int rep = 0; rep < 200; rep++) {
                k_fence<<<1, threads>>>((unsigned long long *)host, mode, dt);
                CK(cudaDeviceSynchronize());
                unsigned long long r;
                CK(cudaMemcpy(&r, dt, 8, cudaMemcpyDeviceToHost));
                v.push_back((double)r);
            }
            printf("%2d posted 8-byte writes to host memory + %-24s %8.0f ns\n", threads,
                   mode == 0 ? "nothing:" : (mode == 1 ? "__threadfence_system():" : "__threadfence():"), median(v));
        }
    }
    {
        volatile unsigned long long *ping = (volatile unsigned long long *)host;
        volatile unsigned long long *pong = (volatile unsigned long long *)(host + 4096);
        *ping = 0;
        *pong = 0;
        const int rounds = 2000;
        k_pingpong<<<1, 1>>>(ping, pong, rounds);
        std::vector<double> v;
        for (int r = 1; r <= rounds; r++) {
            const double t0 = now_us();
            *ping = (unsigned long long)r;
            while (*pong != (unsigned long long)r) __builtin_ia32_pause();
            v.push_back((now_us() - t0) * 1e3);
        }
        CK(cudaDeviceSynchronize());
        printf("host write -> resident kernel sees it -> kernel writes -> host sees it: %8.0f ns (p10 %.0f, p90 %.0f)\n",
               median(v), [&] { auto w = v; std::sort(w.begin(), w.end()); return w[w.size() / 10]; }(),
               [&] { auto w = v; std::sort(w.begin(), w.end()); return w[w.size() * 9 / 10]; }());
    }
    return 0;
}

// Micro-benchmark: issue rate of legacy mma.sync shapes on sm_100a (cycles per instruction per warp, by warps per SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void k(int iters, long long *out, float *sink) {
    float c[4][4];
    for (int j = 0; j < 4; ++j) for (int e = 0; e < 4; ++e) c[j][e] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x * 5u, 11u};
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 4 independent accumulators
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else if (KIND == 2)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(b[0]));
            else if (KIND == 3)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else {  // FFMA baseline: 32 FMAs per "instruction slot group"
#pragma unroll
                for (int e = 0; e < 4; ++e) c[j][e] = fmaf(c[j][e], 1.0001f, 0.5f);
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    float s = 0; for (int j = 0; j < 4; ++j) for (int e = 0; e < 4; ++e) s += c[j][e];
    if (s == 12345.678f) sink[0] = s;
}

int main() {
    long long *out; float *sink;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&sink, 4);
    const char *names[5] = {"tf32 m16n8k8", "bf16 m16n8k16", "tf32 m16n8k4", "f16 m16n8k16", "ffma x4 (per 4 instr)"};
    const int iters = 2000;
    for (int kind = 0; kind < 5; ++kind)
        for (int warps : {1, 4, 8, 16, 32}) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (kind == 0) k<0><<<1, 32 * warps>>>(iters, out, sink);
                if (kind == 1) k<1><<<1, 32 * warps>>>(iters, out, sink);
                if (kind == 2) k<2><<<1, 32 * warps>>>(iters, out, sink);
                if (kind == 3) k<3><<<1, 32 * warps>>>(iters, out, sink);
                if (kind == 4) k<4><<<1, 32 * warps>>>(iters, out, sink);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            double per = (double)h / (iters * 4.0);
            printf("%-24s warps/SM %2d: %.2f cycles per instr per warp -> %.2f instr/clk/SM\n", names[kind], warps, per, warps / per);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

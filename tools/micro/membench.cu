// membench.cu -- what HBM bandwidth do the step kernel's access patterns allow, without compute?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o membench membench.cu && ./membench
// Patterns (N envs, float):
//   copy      : 1 stream in, 1 stream out, float4 per thread
//   planes    : read 8 planes + actions, write 6 planes + reward + flags (65 B/env), plane stride ld
//   tiled     : same bytes, but the 8 planes of a 1024-env tile are contiguous ([tile][plane][1024])
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void k_copy(const float4* __restrict__ in, float4* __restrict__ out, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i];
}

template <bool TILED>
__global__ void __launch_bounds__(256, 2) k_planes(float* __restrict__ st, int64_t ld, int64_t n, const int4* __restrict__ act,
                                                   float4* __restrict__ rew, uint32_t* __restrict__ fl) {
    const int64_t tile_envs = 1024;
    for (int64_t tile = blockIdx.x; tile * tile_envs < n; tile += gridDim.x) {
        const int64_t b = tile * tile_envs + threadIdx.x * 4;
        float4 v[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const float* src = TILED ? st + (tile * 8 + p) * tile_envs + threadIdx.x * 4 : st + p * ld + b;
            v[p] = *reinterpret_cast<const float4*>(src);
        }
        const int4 a = act[b / 4];
        float s = (float)(a.x + a.y + a.z + a.w);
#pragma unroll
        for (int p = 0; p < 8; ++p) { v[p].x += s; v[p].y += s; v[p].z += s; v[p].w += s; }
        const int order[6] = {0, 1, 2, 5, 6, 7};
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const int p = order[q];
            float* dst = TILED ? st + (tile * 8 + p) * tile_envs + threadIdx.x * 4 : st + p * ld + b;
            *reinterpret_cast<float4*>(dst) = v[p];
        }
        rew[b / 4] = v[3];
        fl[b / 4] = __float_as_uint(v[4].x) & 0x01010101u;
    }
}

int main() {
    const int64_t n = 1 << 24, ld = n;
    float *st, *in, *out; int4* act; float4* rew; uint32_t* fl;
    CK(cudaMalloc(&st, 8 * ld * 4)); CK(cudaMalloc(&in, 8 * ld * 4)); CK(cudaMalloc(&out, 8 * ld * 4));
    CK(cudaMalloc(&act, n * 4)); CK(cudaMalloc(&rew, n * 4)); CK(cudaMalloc(&fl, n));
    CK(cudaMemset(st, 0, 8 * ld * 4)); CK(cudaMemset(in, 0, 8 * ld * 4)); CK(cudaMemset(act, 0, n * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto time = [&](auto launch, double bytes, const char* name) {
        for (int i = 0; i < 5; ++i) launch();
        cudaEventRecord(e0);
        const int reps = 50;
        for (int i = 0; i < reps; ++i) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-28s %8.1f GB/s  (%.4f ms)\n", name, bytes * reps / (ms * 1e-3) / 1e9, ms / reps);
    };
    time([&] { k_copy<<<sms * 8, 256>>>((const float4*)in, (float4*)out, 8 * ld / 4); }, 2.0 * 8 * ld * 4, "copy 512MiB->512MiB");
    for (int grid : {sms * 2, sms * 4, sms * 8}) {
        char nm[64];
        snprintf(nm, sizeof nm, "planes grid=%d", grid);
        time([&] { k_planes<false><<<grid, 256>>>(st, ld, n, act, rew, fl); }, 65.0 * n, nm);
        snprintf(nm, sizeof nm, "tiled  grid=%d", grid);
        time([&] { k_planes<true><<<grid, 256>>>(st, ld, n, act, rew, fl); }, 65.0 * n, nm);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}

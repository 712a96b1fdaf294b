// membench.cu — HBM ceilings for the access mixes the field query produces (run on the B200 box):
//   write-only streams (the 4 KB/point descriptor output dominates cfg2a), read-only, and copy.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/membench tools/membench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE>  // 0: st.global (default), 1: st.global.cs, 2: st.global.wt? (use __stwt), 3: __stcg
__global__ void k_write(float4* __restrict__ dst, size_t n4, float v) {
    const float4 val = make_float4(v, v, v, v);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        if (MODE == 0) dst[i] = val;
        else if (MODE == 1) __stcs(dst + i, val);
        else if (MODE == 2) __stwt(dst + i, val);
        else __stcg(dst + i, val);
    }
}
// each warp writes 512 contiguous bytes per row of 4 KB, 8 warps per row, rows consecutive: the kernel's pattern
__global__ void k_write_rows(float4* __restrict__ dst, size_t rows, float v) {
    const float4 val = make_float4(v, v, v, v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t r0 = (size_t)blockIdx.x * 128;
    for (int p = 0; p < 128 && r0 + p < rows; ++p) __stcs(dst + (r0 + p) * 256 + warp * 32 + lane, val);
}
__global__ void k_read(const float4* __restrict__ src, size_t n4, float* out) {
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = __ldcs(src + i);
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 12345.678f) *out = acc;
}
__global__ void k_copy(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        __stcs(dst + i, __ldcs(src + i));
}

template <typename F> float time_ms(F f, int reps = 10) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); f(); CK(cudaDeviceSynchronize());
    std::vector<float> t;
    for (int i = 0; i < reps; ++i) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); t.push_back(ms); }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main() {
    const size_t bytes = 4096ull * 1000000ull;          // the cfg2a output: 1M rows x 4 KB
    const size_t n4 = bytes / 16;
    float4 *a, *b; float* o;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&o, 4));
    CK(cudaMemset(a, 0, bytes)); CK(cudaMemset(b, 0, bytes));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("{\"sms\": %d, \"bytes\": %zu", sms, bytes);
    for (int bpsm : {2, 4, 8, 16}) {
        int grid = sms * bpsm;
        float ms;
        ms = time_ms([&] { k_write<0><<<grid, 256>>>(a, n4, 1.f); });  printf(", \"write_default_g%d\": %.1f", bpsm, bytes / ms / 1e6);
        ms = time_ms([&] { k_write<1><<<grid, 256>>>(a, n4, 1.f); });  printf(", \"write_cs_g%d\": %.1f", bpsm, bytes / ms / 1e6);
        ms = time_ms([&] { k_write<3><<<grid, 256>>>(a, n4, 1.f); });  printf(", \"write_cg_g%d\": %.1f", bpsm, bytes / ms / 1e6);
        ms = time_ms([&] { k_read<<<grid, 256>>>(a, n4, o); });        printf(", \"read_g%d\": %.1f", bpsm, bytes / ms / 1e6);
        ms = time_ms([&] { k_copy<<<grid, 256>>>(a, b, n4); });        printf(", \"copy_rw_g%d\": %.1f", bpsm, 2.0 * bytes / ms / 1e6);
    }
    {
        float ms = time_ms([&] { k_write_rows<<<(unsigned)((1000000 + 127) / 128), 256>>>(a, 1000000, 1.f); });
        printf(", \"write_rows_tilepattern\": %.1f", bytes / ms / 1e6);
        ms = time_ms([&] { CK(cudaMemsetAsync(a, 0, bytes)); });
        printf(", \"cudaMemset\": %.1f", bytes / ms / 1e6);
        ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); });
        printf(", \"cudaMemcpyD2D_rw\": %.1f", 2.0 * bytes / ms / 1e6);
    }
    printf(", \"unit\": \"GB/s\"}\n");
    return 0;
}

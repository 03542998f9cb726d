// bulk_store_ubench.cu -- how fast does one SM's bulk-copy engine move shared memory to global memory?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulk_store_ubench bulk_store_ubench.cu && ./bulk_store_ubench
// One thread per CTA issues `reps` copies of `bytes` (cp.async.bulk.global.shared::cta), each committed and waited for
// (.read: the source may be reused) before the next; cycles per copy, for one CTA alone and for one CTA on every SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void bulk_store_kernel(char *dst, int bytes, int reps, int pieces, long long *cycles, long long *issue_cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) smem[i] = (unsigned char)i;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        char *mine = dst + (size_t)blockIdx.x * bytes * 2;
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(smem);
        long long t_issue = 0;
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const long long a = clock64();
            const int piece = bytes / pieces;
            for (int p = 0; p < pieces; ++p)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mine + (size_t)(r & 1) * bytes + (size_t)p * piece),
                             "r"(src + p * piece), "r"(piece) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            t_issue += clock64() - a;
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        const long long t1 = clock64();
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        cycles[blockIdx.x] = (t1 - t0) / reps;
        issue_cycles[blockIdx.x] = t_issue / reps;
    }
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    char *dst;
    long long *cyc, *iss;
    cudaMalloc(&dst, (size_t)sms * 2 * 65536 * 2);
    cudaMallocManaged(&cyc, sizeof(long long) * sms);
    cudaMallocManaged(&iss, sizeof(long long) * sms);
    cudaFuncSetAttribute(bulk_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int sizes[] = {2048, 4608, 18432, 36864, 65536};
    for (int grid : {1, sms})
        for (int bytes : sizes)
            for (int pieces : {1, 2, 8}) {
                bulk_store_kernel<<<grid, 128, bytes>>>(dst, bytes, 200, pieces, cyc, iss);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long worst = 0, worst_issue = 0;
                for (int i = 0; i < grid; ++i) { if (cyc[i] > worst) worst = cyc[i]; if (iss[i] > worst_issue) worst_issue = iss[i]; }
                printf("CTAs %3d  %6d B in %d piece(s): %6lld cycles per copy until read (%.1f B/cycle), issue %lld cycles\n", grid, bytes, pieces, worst,
                       (double)bytes / (double)worst, worst_issue);
            }
    return 0;
}

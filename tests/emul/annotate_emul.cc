// tests/emul/annotate_emul.cc — TEST INFRASTRUCTURE ONLY.  NOT part of the product, never linked into libregtools_jx.so.
//
// Host emulation harness for the `junctions annotate` device code: regtools_b200/csrc/annotate.cu is compiled by g++ with
// the CUDA qualifiers defined away and its kernel body run in a plain loop over "threads", and the handful of CUDA runtime
// calls annotate.cc makes are replaced by malloc / memcpy stand-ins defined HERE (the binary does not link libcudart).
// Purpose: check the logic of the kernel source and of the host code around it (GTF / BED readers, flat arrays, item
// buffer growth, TSV writer) against the oracle and the unmodified reference in the CPU test suite, where no GPU exists.
// It says nothing about the GPU build (race freedom, performance): tests/test_gpu_zzz_annotate.py does that on a B200.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>

#include <cuda_runtime.h>       // types and prototypes only

// ---- device built-ins the kernel source uses
struct Dim3Emul { unsigned x, y, z; };
static Dim3Emul blockIdx = {0, 1, 1}, blockDim = {128, 1, 1}, threadIdx = {0, 0, 0};
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
static inline uint32_t atomicMin(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
static inline uint32_t atomicExch(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = v; return o; }

#define RTJX_HOST_EMULATION 1
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#include "../../regtools_b200/csrc/annotate.cu"

namespace rtjx {
void launch_annotate(const AnnGtfView& g, const AnnJunctionView& j, int skip_single_exon, unsigned long long* items,
                     unsigned long long items_cap, AnnOut* out, uint32_t* counters, cudaStream_t) {
    const unsigned blocks = (j.n + 127u) / 128u;
    // descending block order on purpose: the item reservations then happen in another order than the junction order
    for (unsigned b = blocks; b-- > 0;)
        for (unsigned t = 0; t < 128; ++t) {
            blockIdx.x = b; threadIdx.x = t;
            annotate_kernel(g, j, skip_single_exon, items, items_cap, out, counters);
        }
}
}  // namespace rtjx

// ---- stand-ins for the CUDA runtime calls of annotate.cc (host memory)
extern "C" {
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
}

#include "../../include/rtjx.h"

int main(int argc, char** argv) {
    int include_single = 0; const char* out_path = nullptr;
    int c;
    while ((c = getopt(argc, argv, "So:")) != -1) {
        if (c == 'S') include_single = 1;
        else if (c == 'o') out_path = optarg;
        else return 2;
    }
    if (argc - optind != 3) { fprintf(stderr, "usage: annotate_emul [-S] [-o out] junctions.bed ref.fa ann.gtf\n"); return 2; }
    rtjx_annotate_params p;
    rtjx_annotate_params_default(&p);
    p.junctions_bed = argv[optind]; p.fasta = argv[optind + 1]; p.gtf = argv[optind + 2];
    p.include_single_exon = include_single; p.chatter_fd = 2;
    p.out_path = out_path;
    char err[512]; uint64_t n = 0;
    const int rc = rtjx_annotate(&p, out_path ? -1 : 1, &n, err, sizeof err);
    if (rc) fprintf(stderr, "%s\n", err);
    return rc ? 1 : 0;
}
